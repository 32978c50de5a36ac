"""verticut_b200 - B200-native MIH / linear-scan k-NN over binary codes (the hot path of tu-dresden/verticut).

The product is the CUDA library `verticut_b200/lib/libverticut_gpu.so` behind the C ABI of
`include/verticut_gpu.h`; `verticut_b200.capi` is its ctypes binding, `verticut_b200/host/` the C++ mirror
of the reference's own interfaces (BaseProxy, SearchWorker, CLIs), `verticut_b200.sharded` the one-process-per-GPU
id-sharded search with an NCCL all-gather top-k merge.
"""
from . import capi  # noqa: F401
from .capi import Index, VerticutError  # noqa: F401
