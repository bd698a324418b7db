"""Calibration surface of the reference: collect the activations entering the projections,
turn them into the 10 000-bin histograms that `Distribution` / `SparsifyFn.set_threshold` consume, write
``histograms.pt`` in the reference's format.

Mirrors ``ActivationModule`` of utils/utils.py:125-185 (same method names and file layout, wired by
teal/mlp.py:33,44-47, teal/self_attn.py:41,71,147 and teal/grab_acts.py:86-89).  Load-time / offline work: plain torch
ops on whatever device the activations live on -- the reference hard-codes ``.to('cuda')`` for the sort and comes back
to the CPU for ``torch.histogram``; here the tensors stay where they are.
"""
from __future__ import annotations

import os
from collections import defaultdict
from typing import Dict

import torch


def find_histogram(acts: torch.Tensor, num_bins: int = 10000, outlier_threshold: float = 0.01):
    """(counts, bin_centers), both fp32, in the format `Distribution` reads -- the numbers of utils/utils.py:145-173:
    ``num_bins - 2`` uniform bins between the 1 % and 99 % order statistics plus one catch-all bin on either side out
    to the extreme values; centres are the edge mid-points.

    The values are sorted once (needed for the order statistics anyway), so the counts are differences of the edges'
    insertion points in the sorted array instead of a second pass (`torch.histogram` semantics: bins are
    [e_i, e_{i+1}), the last one closed)."""
    v = torch.sort(acts.detach().flatten().float())[0]
    n = v.numel()
    cut = int(outlier_threshold * n)
    lo, hi = v[cut], v[n - cut] if cut else v[0]   # v[-0] is v[0] in the reference's indexing, too
    edges = torch.cat([v[:1].cpu(), torch.linspace(lo.cpu(), hi.cpu(), num_bins - 1), v[-1:].cpu()])
    at = torch.searchsorted(v, edges.to(v.device), right=False)   # values < edge
    at[-1] = n                                                     # last bin is closed on the right
    at[0] = 0
    counts = (at[1:] - at[:-1]).clamp_(min=0).to(torch.float32).cpu()
    centers = ((edges[:-1] + edges[1:]) / 2).to(torch.float32)
    return counts, centers


class ActivationModule:
    """utils/utils.py:125-185."""

    def __init__(self, file_path: str):
        self.file_path = file_path
        self.activations = defaultdict(list)
        self.histograms = None
        self.store: Dict[str, object] = {}  # scratch the reference keeps here (position ids of the attention block)

    def grab_activations(self, x: torch.Tensor, key: str) -> None:
        if x.size(1) > 1:  # prefill only, like the reference (utils.py:135-136)
            self.activations[key].append(x.detach().squeeze(0).cpu().float())

    def combine_activations(self) -> Dict[str, torch.Tensor]:
        return {key: torch.cat(acts, dim=0) for key, acts in self.activations.items()}

    def save_activations(self) -> None:
        self.activations = self.combine_activations()
        torch.save(self.activations, f"{self.file_path}/activations.pt")

    def load_activations(self) -> None:
        self.activations = torch.load(f"{self.file_path}/activations.pt")

    def find_histogram(self, num_bins: int = 10000, outlier_threshold: float = 0.01) -> Dict[str, torch.Tensor]:
        if self.histograms is not None:
            return self.histograms
        self.activations = self.combine_activations()
        self.histograms = {}
        for key, acts in self.activations.items():
            counts, centers = find_histogram(acts, num_bins, outlier_threshold)
            self.histograms[key] = counts.cpu()
            self.histograms[f"{key}_centers"] = centers.cpu()
        return self.histograms

    def save_histogram(self) -> None:
        os.makedirs(self.file_path, exist_ok=True)
        torch.save(self.histograms, f"{self.file_path}/histograms.pt")
