"""teal_b200 -- B200-native activation-sparse batch-1 decode path behind TEAL's operator surface.

Python/PyTorch here is plumbing (device memory, streams, torch.distributed); the product is the
hand-written sm_100a CUDA in ``csrc/`` reached through the C ABI of ``include/teal_b200.h``.
"""
from .ops import (SparseGEMV, SparseQKVGEMV, qkv_gemv, splitk_sparse_gemv, sparse_gate_up, survivors,
                  to_column_major)
from .distribution import Distribution, layer_thresholds, get_layer_greedy_sparsities

__all__ = ["SparseGEMV", "SparseQKVGEMV", "qkv_gemv", "splitk_sparse_gemv", "sparse_gate_up", "survivors",
           "to_column_major", "Distribution", "layer_thresholds", "get_layer_greedy_sparsities"]
