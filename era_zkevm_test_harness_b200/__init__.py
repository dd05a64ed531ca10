"""B200-native prover hot path for the era-zkevm test harness (host side above the C ABI of libzkgpu.so)."""
from ._lib import ZkGpuError, load, LIB_PATH  # noqa: F401
from .context import GpuContext  # noqa: F401
