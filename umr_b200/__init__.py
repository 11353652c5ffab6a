"""umr_b200 -- B200-native (sm_100a) differentiable soft rasteriser + geometric-loss kernels,
a drop-in for the hot path of NVlabs/UMR (include/umr_b200.h)."""
__version__ = "0.1.0"
