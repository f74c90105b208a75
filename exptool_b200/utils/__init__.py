"""Drop-in mirrors of exptool.utils.{halo_methods,integrate} for the BFE hot path."""
