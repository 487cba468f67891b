"""B200-native Zephyr pose-hypothesis scoring (OSSID hot path).

Host code is Python/PyTorch; all arithmetic runs in hand-written sm_100a CUDA
kernels behind the C-ABI declared in ``include/zs.h``.  There is no CPU fallback.
"""
__version__ = "0.1.0"
