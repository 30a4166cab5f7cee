"""brats21_b200 — B200 (sm_100a) kernels and drop-in host modules for the BraTS21 segmentation hot path."""
__version__ = "0.1.0"
