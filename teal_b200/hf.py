"""HF-side operator surface: the thresholded Linear of teal/mlp.py:37-57 and
teal/self_attn.py:46-156, with the decode step routed through the fused sm_100a kernel.

* ``SparsifyFn`` keeps the reference's API and semantics (utils/utils.py:9-55): ``set_threshold(sparsity)``
  (``icdf(0.5 + s/2)``, exactly 0.0 at s == 0), ``forward`` with the prefill half-sequence rule, ``apply``.
* ``SparseLinear`` = ``proj(sparse_fn(x))`` for one projection: at decode (``q_len == 1``, batch 1, fp16, CUDA) it is
  ONE launch of the gather-GEMV (the masked lanes' weight rows are never read); otherwise it is literally the
  reference's mask-then-dense-matmul.  It keeps a one-time ``[Z][N]`` copy of the HF ``[N][Z]`` weight
  (generate.py:296-317's re-layout).
* ``patch_mlp`` / ``mlp_forward`` reproduce teal/mlp.py:14-57 on any module with gate_proj / up_proj / down_proj.
"""
from __future__ import annotations

import types
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class SparsifyFn(nn.Module):
    def __init__(self, distr=None, init_sparsity: Optional[float] = None, init_threshold: Optional[float] = None,
                 apply_prefill: bool = True):
        super().__init__()
        assert init_sparsity is None or init_threshold is None, "init_sparsity and init_threshold cannot both be specified"
        self.distr = distr
        self.apply_prefill = apply_prefill
        self.threshold = 0.0
        self.sparsity_level = 0.0
        if init_sparsity is not None:
            self.set_threshold(init_sparsity)
        elif init_threshold is not None:
            self.threshold = float(init_threshold)

    def set_threshold(self, sparsity: float):                      # utils/utils.py:28-30
        self.threshold = float(self.distr.icdf(0.5 + sparsity / 2)) if sparsity != 0.0 else 0.0
        self.sparsity_level = sparsity

    def get_threshold(self):
        return self.threshold

    def apply(self, x: torch.Tensor) -> torch.Tensor:             # utils/utils.py:51-52
        return x.abs().gt(self.threshold) * x

    def forward(self, x: torch.Tensor) -> torch.Tensor:           # utils/utils.py:32-49
        if x.size(1) > 1 and self.apply_prefill:
            half = x.size(1) // 2
            return torch.cat((x[:, :-half, :], self.apply(x[:, -half:, :])), dim=1)
        if x.size(1) > 1 and not self.apply_prefill:
            return x
        assert x.size(1) == 1, "supposedly x is decode only"
        return self.apply(x)


class SparseLinear(nn.Module):
    """``linear(sparse_fn(x))`` with the decode step on the fused kernel."""

    def __init__(self, linear: nn.Linear, sparse_fn: SparsifyFn):
        super().__init__()
        assert linear.bias is None, "TEAL's projections have no bias"
        self.linear = linear
        self.sparse_fn = sparse_fn
        self._wt = None  # [Z][N] storage, logical [N, Z]

    def _column_major(self) -> torch.Tensor:
        w = self.linear.weight
        if self._wt is None or self._wt.device != w.device or self._wt.data_ptr() == 0:
            self._wt = ops.to_column_major(w.detach().to(torch.float16))
        return self._wt

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        fast = x.is_cuda and x.dim() == 3 and x.shape[0] == 1 and x.shape[1] == 1 and x.dtype is torch.float16
        if fast:
            return ops.splitk_sparse_gemv(x, self._column_major(), self.sparse_fn.threshold, 0)
        return self.linear(self.sparse_fn(x))


def mlp_forward(self, x):
    """teal/mlp.py:37-57: grabbing mode records the two activation sites of a dense forward; otherwise the three
    projections see their thresholded inputs."""
    if getattr(self, "grabbing_mode", False):
        self.activation_module.grab_activations(x, "h1")
        inter = self.act_fn(self.gate_proj(x)) * self.up_proj(x)
        self.activation_module.grab_activations(inter, "h2")
        return self.down_proj(inter)
    inter = self.act_fn(self.sparse_linears["gate"](x)) * self.sparse_linears["up"](x)
    return self.sparse_linears["down"](inter)


def patch_mlp_for_calibration(mlp: nn.Module, file_path: str) -> nn.Module:
    """teal/mlp.py:14-35 with ``grabbing_mode=True``: no thresholds yet, an ``ActivationModule`` collects h1 / h2 of
    every prefill; ``mlp.activation_module.find_histogram(); .save_histogram()`` then writes the ``histograms.pt``
    that `patch_mlp` / `Distribution` read (teal/grab_acts.py:86-89)."""
    from .calibration import ActivationModule
    mlp.forward_old = mlp.forward
    mlp.grabbing_mode = True
    mlp.file_path = file_path
    mlp.activation_module = ActivationModule(file_path)
    mlp.forward = types.MethodType(mlp_forward, mlp)
    return mlp


def patch_mlp(mlp: nn.Module, distrs: Dict[str, object]) -> nn.Module:
    """teal/mlp.py:14-35: installs sparse_fns {'gate','up' <- h1; 'down' <- h2} and swaps the forward."""
    mlp.forward_old = mlp.forward
    mlp.distrs = distrs
    mlp.sparse_fns = nn.ModuleDict({"gate": SparsifyFn(distrs["h1"]), "up": SparsifyFn(distrs["h1"]),
                                    "down": SparsifyFn(distrs["h2"])})
    mlp.sparse_linears = nn.ModuleDict({"gate": SparseLinear(mlp.gate_proj, mlp.sparse_fns["gate"]),
                                        "up": SparseLinear(mlp.up_proj, mlp.sparse_fns["up"]),
                                        "down": SparseLinear(mlp.down_proj, mlp.sparse_fns["down"])})
    mlp.forward = types.MethodType(mlp_forward, mlp)
    return mlp


def patch_self_attn(attn: nn.Module, distrs: Dict[str, object]) -> nn.Module:
    """teal/self_attn.py:21-43: sparse_fns {'q','k','v' <- h1; 'o' <- h2}.  Instead of re-implementing the HF
    attention forward (the reference copies the flash-attn forward, :46-156), the four projection modules are
    swapped for ``SparseLinear`` wrappers, so whatever attention forward the installed transformers version uses
    computes ``proj(sparse_fns[p](x))`` -- and at decode the gather kernel."""
    attn.distrs = distrs
    attn.sparse_fns = nn.ModuleDict({"q": SparsifyFn(distrs["h1"]), "k": SparsifyFn(distrs["h1"]),
                                     "v": SparsifyFn(distrs["h1"]), "o": SparsifyFn(distrs["h2"])})
    for p in ("q", "k", "v", "o"):
        name = f"{p}_proj"
        setattr(attn, name, SparseLinear(getattr(attn, name), attn.sparse_fns[p]))
    return attn


def set_sparsities(layers, sparsities: Dict[str, list]) -> None:
    """teal/model.py:142-147 (set_sparsities): per-layer, per-projection sparsity -> thresholds; the dict is what
    ``teal_b200.distribution.get_layer_greedy_sparsities`` returns (utils/utils.py:243-258) or a uniform fill."""
    for i, layer in enumerate(layers):
        for p in ("q", "k", "v", "o"):
            layer.self_attn.sparse_fns[p].set_threshold(sparsities[p][i])
        for p in ("gate", "up", "down"):
            layer.mlp.sparse_fns[p].set_threshold(sparsities[p][i])
