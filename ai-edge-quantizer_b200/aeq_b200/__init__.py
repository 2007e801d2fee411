"""aeq_b200 — B200-native numeric core for AI Edge Quantizer.

Host side is Python mirroring the reference's plugin surface
(`algorithm_manager`, `qtyping`, `recipe`, per-algorithm
`get_tensor_quant_params`); all arithmetic runs in hand-written sm_100a CUDA
kernels reached through the C ABI in include/aeqb200.h (libaeqb200.so).
There is no CPU fallback: importing `aeq_b200.device` / calling any algorithm
without the built library or without a GPU raises.
"""
__version__ = "0.1.0"
