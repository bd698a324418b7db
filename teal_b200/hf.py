"""HF-side operator surface: the thresholded Linear of teal/mlp.py:37-57 and
teal/self_attn.py:46-156, with the decode step routed through the fused sm_100a kernel.

* ``SparsifyFn`` keeps the reference's API and semantics (utils/utils.py:9-55): ``set_threshold(sparsity)``
  (``icdf(0.5 + s/2)``, exactly 0.0 at s == 0), ``forward`` with the prefill half-sequence rule, ``apply``.
* ``SparseLinear`` = ``proj(sparse_fn(x))`` for one projection: at decode (``q_len == 1``, batch 1, fp16, CUDA) it is
  ONE launch of the gather-GEMV (the masked lanes' weight rows are never read); otherwise it is literally the
  reference's mask-then-dense-matmul.  The HF ``[N][Z]`` weight is re-laid out in place to ``[Z][N]`` storage
  (generate.py:296-317's re-layout), no second copy.
* ``SparseModelSetters`` / ``sparse_model_class`` / ``get_sparse_model``: the model-level surface of teal/model.py:45-155
  and utils/utils.py:201-221 (``from_pretrained(..., histogram_path, uniform_sparsity | greedy_* | ...)``, the setters).
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
    """``linear(sparse_fn(x))`` with the decode step on the fused kernel.  The weight is re-laid out IN PLACE like the
    gpt-fast path does (generate.py:296-317: ``weight.data = weight.data.T.contiguous().T`` -- storage [Z][N], logical
    shape unchanged): one copy serves both branches (``F.linear`` accepts the strided weight), so a sparsified model
    takes no more memory than the dense one, and a later ``load_state_dict`` / in-place update cannot leave the decode
    path on stale weights (the layout is re-checked on every call)."""

    def __init__(self, linear: nn.Linear, sparse_fn: SparsifyFn):
        super().__init__()
        assert linear.bias is None, "TEAL's projections have no bias"
        self.linear = linear
        self.sparse_fn = sparse_fn

    def _column_major(self) -> torch.Tensor:
        w = self.linear.weight
        if w.dtype is not torch.float16:
            raise TypeError("the fused decode kernel is fp16 only (README.md:108)")
        if not (w.stride(0) == 1 and w.stride(1) > 1):  # first decode call, or the weight was re-assigned since
            w.data = ops.to_column_major(w.data)
        return w.data

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        fast = (x.is_cuda and x.dim() == 3 and x.shape[0] == 1 and x.shape[1] == 1 and x.dtype is torch.float16
                and self.linear.weight.dtype is torch.float16)
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


# ------------------------------------------------------------------------------------ model-level surface
class SparseModelSetters:
    """The model-level setters of the reference's ``SparseModelMixin`` (teal/model.py:95-155) for any HF-style causal LM
    whose decoder layers live in ``self.model.layers`` with ``.mlp`` / ``.self_attn``: same names, same semantics.
    Mixed into a transformers class by `sparse_model_class`; usable on its own around an already built model."""

    def build_sparse_layers(self, histogram_path: str, grab_acts: bool = False):
        """teal/model.py:108-120: freeze the parameters and patch every decoder layer; the thresholds come from
        ``<histogram_path>/layer-i/{mlp,self_attn}/histograms.pt`` (grab_acts: collect them instead)."""
        import os
        from .distribution import Distribution
        for _, prm in self.model.named_parameters():
            prm.requires_grad = False
        for i, layer in enumerate(self.model.layers):
            path = f"{histogram_path}/layer-{i}"
            layer.path, layer.grabbing_mode = path, grab_acts
            if grab_acts:
                patch_mlp_for_calibration(layer.mlp, f"{path}/mlp")
                continue
            d_mlp = {h: Distribution(os.path.join(path, "mlp"), h) for h in ("h1", "h2")}
            d_att = {h: Distribution(os.path.join(path, "self_attn"), h) for h in ("h1", "h2")}
            patch_mlp(layer.mlp, d_mlp)
            patch_self_attn(layer.self_attn, d_att)

    def set_grabbing_mode(self, mode: bool):
        for layer in self.model.layers:
            layer.mlp.grabbing_mode = mode
            layer.self_attn.grabbing_mode = mode

    def set_apply_prefill(self, apply_prefill: bool):
        for layer in self.model.layers:
            for p in ("q", "k", "v", "o"):
                layer.self_attn.sparse_fns[p].apply_prefill = apply_prefill
            for p in ("gate", "up", "down"):
                layer.mlp.sparse_fns[p].apply_prefill = apply_prefill

    def set_mlp_sparsity(self, sparsity: float):
        for layer in self.model.layers:
            for p in ("gate", "up", "down"):
                layer.mlp.sparse_fns[p].set_threshold(sparsity)

    def set_self_attn_sparsity(self, sparsity: float):
        for layer in self.model.layers:
            for p in ("q", "k", "v", "o"):
                layer.self_attn.sparse_fns[p].set_threshold(sparsity)

    def set_uniform_sparsity(self, sparsity: float):
        self.set_mlp_sparsity(sparsity)
        self.set_self_attn_sparsity(sparsity)

    def reset_sparsities(self):
        self.set_uniform_sparsity(0)

    def set_sparsities(self, sparsities: Dict[str, list]):
        set_sparsities(self.model.layers, sparsities)

    def load_greedy_sparsities(self, greedy_sparsity_path: str, greedy_sparsity_level: float):
        from .distribution import get_layer_greedy_sparsities
        levels = [greedy_sparsity_level] * len(self.model.layers)
        self.set_sparsities(get_layer_greedy_sparsities(levels, greedy_sparsity_path))

    def configure_sparsity(self, histogram_path: str, grab_acts: bool = False, greedy_sparsity_path: Optional[str] = None,
                           greedy_sparsity_level: Optional[float] = None, uniform_sparsity: Optional[float] = None,
                           mlp_sparsity: Optional[float] = None, self_attn_sparsity: Optional[float] = None,
                           apply_prefill: bool = True):
        """What ``SparseModelMixin.from_pretrained`` does after the model exists (teal/model.py:66-93)."""
        assert histogram_path is not None, "histogram_path must be provided"
        self.build_sparse_layers(histogram_path, grab_acts)
        if grab_acts:
            self.set_grabbing_mode(True)
            return self
        if greedy_sparsity_path is not None:
            assert greedy_sparsity_level is not None, "greedy_sparsity_level must be provided"
            self.load_greedy_sparsities(greedy_sparsity_path, greedy_sparsity_level)
        elif uniform_sparsity is not None:
            self.set_uniform_sparsity(uniform_sparsity)
        elif mlp_sparsity is not None or self_attn_sparsity is not None:
            if mlp_sparsity is not None:
                self.set_mlp_sparsity(mlp_sparsity)
            if self_attn_sparsity is not None:
                self.set_self_attn_sparsity(self_attn_sparsity)
        else:
            self.reset_sparsities()
        self.set_apply_prefill(apply_prefill)
        return self


_SPARSE_KW = ("histogram_path", "grab_acts", "greedy_sparsity_path", "greedy_sparsity_level", "uniform_sparsity", "mlp_sparsity",
              "self_attn_sparsity", "apply_prefill")


def sparse_model_class(base_cls):
    """``LlamaSparseForCausalLM`` / ``MistralSparseForCausalLM`` of teal/model.py:165-171 for any transformers causal-LM
    class: ``from_pretrained`` takes the reference's extra keyword arguments (teal/model.py:45-93)."""

    class _Sparse(SparseModelSetters, base_cls):
        @classmethod
        def from_pretrained(cls, pretrained_model_name_or_path, *model_args, **kwargs):
            extra = {k: kwargs.pop(k) for k in _SPARSE_KW if k in kwargs}
            model = super().from_pretrained(pretrained_model_name_or_path, *model_args, **kwargs)
            return model.configure_sparsity(**extra)

    _Sparse.__name__ = base_cls.__name__.replace("ForCausalLM", "SparseForCausalLM")
    return _Sparse


def get_sparse_model(model_name: str, device: str, histogram_path: str, **kwargs):
    """utils/utils.py:201-221: AutoConfig -> Llama / Mistral sparse class -> ``from_pretrained(..., torch_dtype=fp16)``."""
    import transformers
    cfg = transformers.AutoConfig.from_pretrained(model_name)
    base = {"llama": "LlamaForCausalLM", "mistral": "MistralForCausalLM"}.get(cfg.model_type)
    if base is None:
        raise ValueError(f"unsupported model_type {cfg.model_type!r} (the reference supports Llama and Mistral)")
    cls = sparse_model_class(getattr(transformers, base))
    return cls.from_pretrained(model_name, torch_dtype=torch.float16, device_map=device, histogram_path=histogram_path, **kwargs)
