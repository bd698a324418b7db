"""Host-side threshold calibration lookup (load-time only, never on the hot path).

Same semantics as the reference's ``Distribution`` (gpt-fast/distribution.py:16-26,48-66, duplicated
at utils/utils.py:72-123), the histogram -> projection mapping of ``monkeypatch_layer``
(gpt-fast/generate.py:266-287) and the greedy lookup (utils/utils.py:243-258), which the reference
imports into its gpt-fast loader but never wires (generate.py:260) -- here it is wired.
"""
from __future__ import annotations

import bisect
import csv
import os
from typing import Dict, List, Sequence

import numpy as np

PROJS = ["q", "k", "v", "o", "gate", "up", "down"]
PROJ_HIST = {"q": ("self_attn", "h1"), "k": ("self_attn", "h1"), "v": ("self_attn", "h1"),
             "o": ("self_attn", "h2"), "gate": ("mlp", "h1"), "up": ("mlp", "h1"), "down": ("mlp", "h2")}


class Distribution:
    """Cumulative histogram with linear-interpolated inverse CDF, evaluated in fp32 like the
    reference's torch code so thresholds agree to the last bit."""

    def __init__(self, file_path: str = None, hidden_type: str = None, *, centers=None, counts=None):
        if file_path is not None:
            import torch
            hist = torch.load(os.path.join(file_path, "histograms.pt"))
            centers, counts = hist[f"{hidden_type}_centers"].numpy(), hist[hidden_type].numpy()
        self.file_path, self.hidden_type = file_path, hidden_type
        self.bin_centers = np.asarray(centers, dtype=np.float32)
        self.counts = np.asarray(counts, dtype=np.float32)
        import torch
        c = torch.from_numpy(self.counts)
        self.total_count = np.float32(c.sum().item())
        # torch.cumsum in fp32 (sequential) -- keep torch here so rounding matches the reference
        self.cumulative_counts = torch.cumsum(c, dim=0).numpy()

    def icdf(self, q: float) -> float:
        f32 = np.float32
        # the reference multiplies a python float by a 0-dim fp32 tensor -> fp32 result
        target = f32(f32(q) * self.total_count)
        idx = int(np.searchsorted(self.cumulative_counts, target, side="left"))
        n = len(self.bin_centers)
        if idx == 0:
            return float(self.bin_centers[0])
        if idx == n:
            return float(self.bin_centers[-1])
        lo_c, hi_c = self.cumulative_counts[idx - 1], self.cumulative_counts[idx]
        lo_v, hi_v = self.bin_centers[idx - 1], self.bin_centers[idx]
        frac = f32(f32(target - lo_c) / f32(hi_c - lo_c))
        return float(f32(lo_v + f32(frac * f32(hi_v - lo_v))))


def threshold_for(distr: Distribution, sparsity: float, hf_zero_rule: bool = False) -> float:
    """gpt-fast: ``icdf(0.5 + 0.5*s)`` (generate.py:279); HF: exactly 0.0 at s == 0 (utils.py:29)."""
    if hf_zero_rule and sparsity == 0.0:
        return 0.0
    return distr.icdf(0.5 + 0.5 * sparsity)


def layer_thresholds(hist_path: str, layer_idx: int, sparsities: Dict[str, float]) -> Dict[str, float]:
    cache: Dict[tuple, Distribution] = {}
    out = {}
    for proj in PROJS:
        sub, h = PROJ_HIST[proj]
        if (sub, h) not in cache:
            cache[(sub, h)] = Distribution(os.path.join(hist_path, f"layer-{layer_idx}", sub), h)
        out[proj] = threshold_for(cache[(sub, h)], sparsities[proj])
    return out


def read_results_csv(path: str) -> List[Dict[str, float]]:
    with open(path, newline="") as f:
        return [{k: float(v) for k, v in row.items()} for row in csv.DictReader(f)]


def greedy_row(rows: Sequence[Dict[str, float]], target: float) -> Dict[str, float]:
    eff = np.array([r["Effective Sparsity"] for r in rows], dtype=np.float64)
    i = int(np.argsort(np.abs(eff - target), kind="stable")[0])
    return {p: float(rows[i][p]) for p in PROJS}


def get_layer_greedy_sparsities(layer_sparsities: Sequence[float], results_dir: str) -> Dict[str, List[float]]:
    out = {p: [0.0] * len(layer_sparsities) for p in PROJS}
    for layer, target in enumerate(layer_sparsities):
        sel = greedy_row(read_results_csv(os.path.join(results_dir, f"layer-{layer}", "results.csv")), target)
        for p in PROJS:
            out[p][layer] = sel[p]
    return out


def resolve_histogram_dir(model_dir: str) -> str:
    """Directory that holds ``layer-i/{mlp,self_attn}/histograms.pt`` for a model of the reference's ``models/`` tree.
    Fixture quirk: ``models/Llama-3-8B/histograms/layer-*`` are byte-identical copies of the
    Llama-2-7B files; the genuine Llama-3-8B histograms sit in ``histograms/Untitled/``."""
    base = os.path.join(model_dir, "histograms")
    untitled = os.path.join(base, "Untitled")
    if os.path.isdir(os.path.join(untitled, "layer-0")):
        return untitled
    return base
