"""
exptool_b200 -- B200-native (sm_100a) implementation of exptool's basis-function-
expansion hot path: EOF / SL coefficient accumulation, force evaluation and leapfrog
orbit integration in the frozen expansion, behind exptool's own Python entry points.

    from exptool_b200.basis import eof, spheresl, potential
    from exptool_b200.utils import integrate, halo_methods

The CUDA library (libbfe.so, C ABI in include/bfe.h) is loaded on first use; there
is no CPU fallback.
"""
__version__ = '0.1.0'
