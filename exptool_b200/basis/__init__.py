"""Drop-in mirrors of exptool.basis.{compatibility,eof,spheresl,potential} for the BFE hot path."""
