"""
compatibility.py -- host-side coordinate maps (exptool/basis/compatibility.py:16-99).

Only used for table geometry (set_table_params, init_table), which runs once on
the host; the per-particle maps are evaluated in the kernels
(csrc/bfe_device.cuh: bfe_r_to_xi, bfe_z_to_y, bfe_d_xi_to_r).
"""
import numpy as np


def r_to_xi(r, cmap, scale):
    '''compatibility.py:16-47; negatives map to 0.  The reference's in-place zeroing of the
    caller's array for cmap=0 (42-43) is not replicated.'''
    r = np.asarray(r, dtype=np.float64)
    scalar_input = r.ndim == 0
    if scalar_input:
        r = r[None]
    if cmap == 1:
        outval = (r / scale - 1.0) / (r / scale + 1.0)
    elif cmap == 2:
        with np.errstate(invalid='ignore', divide='ignore'):
            outval = np.log(r)
    else:
        outval = r.copy()
    outval[r < 0.] = 0.
    if scalar_input:
        return np.squeeze(outval)
    return outval


def xi_to_r(xi, cmap, scale):
    '''compatibility.py:54-62'''
    if cmap == 1:
        return (1.0 + xi) / (1.0 - xi) * scale
    elif cmap == 2:
        return np.exp(xi)
    return xi


def d_xi_to_r(xi, cmap, scale):
    '''compatibility.py:65-80'''
    if cmap == 1:
        return 0.5 * (1.0 - xi) * (1.0 - xi) / scale
    elif cmap == 2:
        return np.exp(-xi)
    return 1.0


def z_to_y(z, hscale):
    '''compatibility.py:83-91'''
    return (z / (np.abs(z) + 1.e-8)) * np.arcsinh(np.abs(z / hscale))


def y_to_z(y, hscale):
    '''compatibility.py:94-99'''
    return hscale * np.sinh(y)
