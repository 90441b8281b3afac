"""quilt_b200 — B200-native (sm_100a) implementation of QUILT2's per-sample imputation hot path.

Only what the path needs lives here: csrc/ (CUDA kernels + the C ABI of include/quilt_b200.h),
api.py (host-side mirror of the reference's rcpp_forwardBackwardGibbsNIPT interface over that ABI),
cabi.py (ctypes structs), synth.py (seeded synthetic inputs) and build.py (nvcc recipe).
"""
__version__ = "0.1.0"
