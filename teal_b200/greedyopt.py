"""Block-wise greedy sparsity allocation (teal/greedyopt.py:26-159): for one decoder layer,
raise the sparsity of one projection at a time -- always the one whose increment hurts the layer's output least -- until
the cost-weighted mean sparsity reaches the target; every step is a row of ``lookup/layer-<i>/results.csv``, the table
``distribution.get_layer_greedy_sparsities`` / ``hf.SparseModelSetters.load_greedy_sparsities`` read back.

Offline calibration work, not the decode path: plain torch ops on whatever device the layer and its activations live on.
Where the reference synchronises after each of the seven candidate forwards of a step (a Python comparison of 0-dim
tensors, teal/greedyopt.py:140-142) and again for each reported number, the candidates' errors stay on the device and a
step makes two host reads (the arg-min, which the uniform-baseline forward depends on, then the two reported errors).
Same arithmetic, same tie rule (the first projection in ``PROJS`` order with the smallest error wins), same CSV.
"""
from __future__ import annotations

import csv
import os
from typing import Callable, Dict, Optional

import torch

from .distribution import PROJS

# Relative cost of a projection's weight matrix per model family (teal/greedyopt.py:26-52): q and o are dim x dim, k and v
# are shrunk by the GQA factor, the three MLP matrices are intermediate/dim times larger.  Constants of the reference.
WEIGHT_DICT: Dict[str, Dict[str, float]] = {
    "Llama-3-8B": dict(q=1, k=1 / 4, v=1 / 4, o=1, gate=3.5, up=3.5, down=3.5),
    "Llama-3-70B": dict(q=1, k=1 / 8, v=1 / 8, o=1, gate=3.5, up=3.5, down=3.5),
    "Llama-2-7B": dict(q=1, k=1 / 8, v=1 / 8, o=1, gate=2.6875, up=2.6875, down=2.6875),
    "Llama-2-13B": dict(q=1, k=1 / 8, v=1 / 8, o=1, gate=2.7, up=2.7, down=2.7),
    "Llama-2-70B": dict(q=1, k=1 / 8, v=1 / 8, o=1, gate=3.5, up=3.5, down=3.5),
    "Mistral-7B": dict(q=1, k=1 / 8, v=1 / 8, o=1, gate=3.5, up=3.5, down=3.5),
}


def effective_sparsity(sparsities: Dict[str, float], weights: Dict[str, float]) -> float:
    """Cost-weighted mean sparsity of a layer (teal/greedyopt.py:65-73; summed in the dict's order, like the reference)."""
    total = sum(weights.values())
    acc = 0
    for proj, value in sparsities.items():
        if proj in weights:
            acc += value * weights[proj]
    return acc / total


def activation_error(target: torch.Tensor, new: torch.Tensor, last_fraction: float = 0.25) -> torch.Tensor:
    """Mean over (batch, hidden) of the L2 norm -- taken along the SEQUENCE axis -- of the output difference over the last
    `last_fraction` of the positions (teal/greedyopt.py:90-94); a 0-dim tensor on the activations' device."""
    start = int(new.shape[1] * (1 - last_fraction))
    return torch.norm(target[:, start:] - new[:, start:], dim=1).mean()


def set_layer_sparsities(layer, sparsities: Dict[str, float]) -> None:
    """teal/greedyopt.py:53-63: the HF layer layout (`layer.mlp.sparse_fns`, `layer.self_attn.sparse_fns`)."""
    for proj in ("gate", "up", "down"):
        layer.mlp.sparse_fns[proj].set_threshold(sparsities[proj])
    for proj in ("q", "k", "v", "o"):
        layer.self_attn.sparse_fns[proj].set_threshold(sparsities[proj])


def layer_forward(layer, hidden_states: torch.Tensor) -> torch.Tensor:
    """One prefill-shaped pass through an HF decoder layer (teal/greedyopt.py:75-87)."""
    bsz, seq_len, _ = hidden_states.shape
    position_ids = torch.arange(seq_len, dtype=torch.long, device=hidden_states.device).unsqueeze(0).repeat(bsz, 1)
    return layer(hidden_states, None, position_ids, None, False, False, None)[0]


def greedy_sparsities(forward: Callable[[Dict[str, float]], torch.Tensor], weights: Dict[str, float], target_sparsity: float,
                      base_step_size: float = 0.05, last_fraction: float = 0.25, csv_path: Optional[str] = None,
                      log: Optional[Callable[[str], None]] = None) -> Dict[str, float]:
    """The greedy loop (teal/greedyopt.py:113-157) over any `forward(sparsities) -> layer output`.

    Step sizes are `base_step_size / weight`, so every increment costs the same effective sparsity; a projection at
    sparsity >= 1 is no longer a candidate; each accepted step appends `[effective sparsity, error, error of the uniform
    allocation with the same effective sparsity, the seven sparsities]` to the CSV."""
    with torch.no_grad():
        sparsities = {p: 0.0 for p in PROJS}
        target = forward(sparsities)
        steps = {p: base_step_size * (1 / weights[p]) for p in PROJS}
        out, writer = None, None
        if csv_path is not None:
            os.makedirs(os.path.dirname(csv_path), exist_ok=True)
            out = open(csv_path, "w", newline="")
            writer = csv.writer(out)
            writer.writerow(["Effective Sparsity", "Activation Error", "Baseline Error"] + list(PROJS))
        try:
            while effective_sparsity(sparsities, weights) < target_sparsity:
                cands = [p for p in PROJS if not sparsities[p] >= 1]
                if not cands:
                    raise RuntimeError("every projection is at sparsity >= 1 below the target effective sparsity")
                errs = []
                for p in cands:
                    trial = dict(sparsities)
                    trial[p] += steps[p]
                    errs.append(activation_error(target, forward(trial), last_fraction))
                errs = torch.stack(errs)
                # `error < best_error` never accepts a NaN (teal/greedyopt.py:140): rank it last
                pick = torch.argmin(torch.where(torch.isnan(errs), torch.full_like(errs, float("inf")), errs))
                best_err = errs[pick]
                # the uniform allocation at the effective sparsity the accepted step reaches: needs the pick on the host
                best = cands[int(pick)]
                sparsities[best] += steps[best]
                eff = effective_sparsity(sparsities, weights)
                base_err = activation_error(target, forward({p: eff for p in PROJS}), last_fraction)
                best_err, base_err = (float(v) for v in torch.stack([best_err, base_err]).tolist())
                if writer is not None:
                    writer.writerow([eff, best_err, base_err] + [sparsities[p] for p in PROJS])
                    out.flush()
                if log is not None:
                    log(f"Updated: Effective Sparsity: {eff:.4f}, Activation Error: {best_err:.4f}, Baseline Error: {base_err:.4f}")
        finally:
            if out is not None:
                out.close()
    return sparsities


def process_layer(layer, model_type: str, layer_idx: int, target_sparsity: float, base_step_size: float, last_fraction: float,
                  teal_path: str, device: Optional[str] = None) -> Dict[str, float]:
    """teal/greedyopt.py:101-159 with the reference's file layout: reads `<teal_path>/activations/act_<i>.pt` (the layer's
    prefill input, written by teal/grab_acts.py), writes `<teal_path>/lookup/layer-<i>/results.csv`, returns the final
    per-projection sparsities.  `device`: where to run (default: CUDA when there is one); the layer goes back to the CPU
    afterwards, like the reference's."""
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    x = torch.load(os.path.join(teal_path, "activations", f"act_{layer_idx}.pt"), map_location="cpu").to(device)
    layer = layer.to(device)

    def forward(sp):
        set_layer_sparsities(layer, sp)
        return layer_forward(layer, x)

    try:
        result = greedy_sparsities(forward, WEIGHT_DICT[model_type], target_sparsity, base_step_size, last_fraction,
                                   csv_path=os.path.join(teal_path, "lookup", f"layer-{layer_idx}", "results.csv"), log=print)
        set_layer_sparsities(layer, result)   # (the reference leaves the layer at the last uniform baseline it evaluated)
    finally:
        layer.to("cpu")
    return result


def main(argv=None):
    """`python -m teal_b200.greedyopt --model_name ... --model_type Llama-3-8B --teal_path ...` (teal/greedyopt.py:162-199):
    all layers of an HF checkpoint, one `results.csv` per layer."""
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_name", type=str, required=True)
    ap.add_argument("--model_type", type=str, required=True, choices=sorted(WEIGHT_DICT))
    ap.add_argument("--teal_path", type=str, required=True)
    ap.add_argument("--target_sparsity", type=float, default=0.9, help="Target effective sparsity")
    ap.add_argument("--base_step_size", type=float, default=0.05, help="Base step size for sparsity updates")
    ap.add_argument("--last_fraction", type=float, default=0.25, help="Fraction of sequence to use for error calculation")
    args = ap.parse_args(argv)
    from .hf import get_sparse_model
    model = get_sparse_model(args.model_name, device="cpu", histogram_path=os.path.join(args.teal_path, "histograms"))
    os.makedirs(os.path.join(args.teal_path, "lookup"), exist_ok=True)
    for i, layer in enumerate(model.model.layers):
        print(f"Processing layer {i}")
        process_layer(layer, args.model_type, i, args.target_sparsity, args.base_step_size, args.last_fraction, args.teal_path)


if __name__ == "__main__":
    main()
