"""B200-native mirror of the `ldmseg` package surface used by the sampling hot path.

Only `ldmseg.models`, `ldmseg.schedulers` and `ldmseg.utils.OutputDict` are provided;
everything is backed by the C-ABI library `lib/libldmseg_b200.so` (hand-written sm_100a kernels).
"""
__version__ = "0.1.0"
