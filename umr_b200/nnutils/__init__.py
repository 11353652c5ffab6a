"""Drop-in replacements for the hot-path modules of the reference's `nnutils` package:
smr.py, geom_utils.py (hot functions), chamfer_python.py, loss_utils.py -- same names, argument
meaning and return values, backed by the sm_100a kernels."""
