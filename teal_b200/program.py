"""Step program: the whole decode step planned once and run as ONE persistent data-flow kernel per token.

Host side of ``teal_step_program_*`` (include/teal_b200.h, csrc/step_program.cu).  The plan is a plain
byte blob (phase table with device pointers embedded) produced by the native planner without touching
CUDA, so the schedule can be built and verified on a CPU-only box (tests/test_step_program.py); the
decoder uploads it verbatim and launches one kernel per token.

Replaces, for the decode step, what ``torch.compile(mode="reduce-overhead")`` builds around the
reference's ops (gpt-fast/generate.py:73-77,420; model.py:133-190,258-259).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterator, List, Optional, Sequence

from . import _lib
from ._lib import ItemInfo, LayerDesc, StepDesc

PROJ_ORDER = ("q", "k", "v", "o", "gate", "up", "down")
KIND_NAMES = {0: "embed", 1: "gemv", 2: "attn", 3: "lm_head", 4: "advance"}
DEP_NONE, DEP_ALL, DEP_LANES, DEP_QKV_HEAD = 0, 1, 2, 3
TUNING_KEYS = ("rows_small", "rows_large", "group_qkv", "group_gate_up", "head_rows", "wide_melems", "ctas_per_sm", "poll_ns",
               "opts", "debug_times", "warps", "stagger_ns", "rows_qkv", "rows_o", "rows_gate_up", "rows_down",
               "tiles_per_item")
ROWS_KIND = {"rows_qkv": 0, "rows_o": 1, "rows_gate_up": 2, "rows_down": 3}


def set_knob(desc: StepDesc, key: str, value: int) -> None:
    """One tuning knob of ``teal_step_desc`` (rows_qkv / rows_o / rows_gate_up / rows_down index rows_kind[])."""
    if key not in TUNING_KEYS:
        raise TypeError(f"unknown step-program knob {key!r}")
    if key in ROWS_KIND:
        desc.rows_kind[ROWS_KIND[key]] = int(value)
    else:
        setattr(desc, key, int(value))


def layer_desc(*, wqkv=0, wo=0, w1=0, w3=0, w2=0, attn_norm=0, ffn_norm=0, k_cache=0, v_cache=0, ld_qkv=0, ld_o=0, ld_1=0,
               ld_3=0, ld_2=0, thr: Optional[Dict[str, float]] = None) -> LayerDesc:
    """One ``teal_layer_desc``; pointers are integer device addresses, ``thr`` maps q,k,v,o,gate,up,down to the
    reference's Python-float thresholds (None / missing = -1 = keep every lane)."""
    d = LayerDesc()
    d.wqkv, d.wo, d.w1, d.w3, d.w2 = wqkv or None, wo or None, w1 or None, w3 or None, w2 or None
    d.attn_norm, d.ffn_norm, d.k_cache, d.v_cache = attn_norm or None, ffn_norm or None, k_cache or None, v_cache or None
    d.ld_qkv, d.ld_o, d.ld_1, d.ld_3, d.ld_2 = ld_qkv, ld_o, ld_1, ld_3, ld_2
    for i, p in enumerate(PROJ_ORDER):
        d.thr[i] = float(thr[p]) if thr is not None and thr.get(p) is not None else -1.0
    return d


class StepProgram:
    """A planned step: host blob + (optionally) its device copy."""

    def __init__(self, desc: StepDesc, layers: Sequence[LayerDesc], lib=None):
        self.lib = lib if lib is not None else _lib.load()  # `lib`: tests plan / run against the CPU emulation build
        self._layers = (LayerDesc * max(1, len(layers)))(*layers)  # keep-alive: desc points into it
        desc.layers = ctypes.cast(self._layers, ctypes.POINTER(LayerDesc))
        desc.n_layer = len(layers)
        self.desc = desc
        nb, ns = ctypes.c_size_t(0), ctypes.c_size_t(0)
        _lib.check(self.lib.teal_step_program_sizes(ctypes.byref(desc), ctypes.byref(nb), ctypes.byref(ns)), "step_program_sizes")
        self.program_bytes, self.scratch_bytes = int(nb.value), int(ns.value)
        self.host = None
        self.device = None  # torch uint8 tensor once uploaded

    @staticmethod
    def sizes(desc: StepDesc, layers: Sequence[LayerDesc]):
        return StepProgram(desc, layers)

    def plan(self) -> "StepProgram":
        """Fill the host blob (``desc.scratch`` must be set: its address is baked into the phase table)."""
        self.host = ctypes.create_string_buffer(self.program_bytes)
        _lib.check(self.lib.teal_step_program_plan(ctypes.byref(self.desc), self.host, self.program_bytes), "step_program_plan")
        return self

    # ---- introspection (CPU) ---------------------------------------------------------------------------
    @property
    def n_items(self) -> int:
        n = self.lib.teal_step_program_items(self.host)
        if n < 0:
            _lib.check(n, "step_program_items")
        return n

    def item(self, ticket: int) -> ItemInfo:
        info = ItemInfo()
        _lib.check(self.lib.teal_step_program_item(self.host, int(ticket), ctypes.byref(info)), "step_program_item")
        return info

    def items(self) -> Iterator[ItemInfo]:
        for t in range(self.n_items):
            yield self.item(t)

    # ---- device ----------------------------------------------------------------------------------------
    def upload(self, device):
        import torch
        blob = torch.frombuffer(bytearray(self.host.raw), dtype=torch.uint8)
        self.device = blob.to(device)
        assert self.device.data_ptr() % 16 == 0
        return self.device

    def run(self, stream_ptr: int) -> None:
        """One launch = one token (asynchronous on `stream_ptr`, CUDA-graph-capturable)."""
        _lib.check(self.lib.teal_step_program_run(self.host, self.device.data_ptr(), stream_ptr), "step_program_run")


def verify_schedule(prog: StepProgram) -> Dict[str, object]:
    """Replays the ticket order on the CPU and checks the invariants the kernel's spin-waits rely on:

    * every completion flag an item waits on is raised by an item with a SMALLER ticket, and a whole-phase wait
      (RMSNorm consumers, lm_head, bookkeeping) only names phases all of whose tickets are smaller
      => waits never point forward => no deadlock for any number of resident CTAs;
    * every publication receives exactly `arrivals` items and is raised exactly once; every flag is owned by
      exactly one publication;
    * the (column tile, K chunk) items of a projection cover each tile's K range exactly once.

    Returns statistics, among them how early the first consumer of a data-flow edge can start."""
    n = prog.n_items
    flag_owner: Dict[int, int] = {}       # flag -> ticket of the item that completes it
    cell_done: Dict[int, List[int]] = {}  # done cell -> completing tickets
    arrivals: Dict[tuple, int] = {}
    need: Dict[tuple, int] = {}
    cover: Dict[tuple, List[tuple]] = {}
    phase_range: Dict[int, List[int]] = {}
    phase_kind: Dict[int, int] = {}
    phase_dep: Dict[int, int] = {}
    slack: Dict[int, List[float]] = {}    # consumer phase -> (ticket - latest dep completion ticket)
    for t in range(n):
        it = prog.item(t)
        phase_range.setdefault(it.phase, [t, t])[1] = t
        phase_kind[it.phase] = it.kind
        phase_dep[it.phase] = it.dep_kind
        if it.dep_kind == DEP_ALL:
            got = cell_done.get(it.dep_cell, [])
            assert len(got) == it.dep_per_epoch, (t, it.phase, "whole-phase wait on an incomplete phase", len(got), it.dep_per_epoch)
            assert max(got) < t
        elif it.dep_kind in (DEP_LANES, DEP_QKV_HEAD):
            assert 0 < it.n_deps <= 32, (t, it.n_deps)
            latest = -1
            for f in list(it.deps)[: it.n_deps]:
                assert f in flag_owner, (t, it.phase, f, "waits on a flag no earlier ticket raises")
                assert flag_owner[f] < t
                latest = max(latest, flag_owner[f])
            slack.setdefault(it.phase, []).append(t - latest)
        else:
            assert it.n_deps == 0
        if it.kind == 1:
            for ct in list(it.ctiles)[: it.n_pubs]:
                cover.setdefault((it.phase, ct), []).append((it.k_begin, it.k_end))
        if it.pub_cell >= 0:
            for pub_index, pub_flag in zip(list(it.pub_indices)[: it.n_pubs], list(it.pub_flags)[: it.n_pubs]):
                key = (it.phase, pub_index)
                arrivals[key] = arrivals.get(key, 0) + 1
                need[key] = it.arrivals
                assert arrivals[key] <= it.arrivals, (t, key)
                if arrivals[key] == it.arrivals:
                    if pub_flag >= 0:
                        assert pub_flag not in flag_owner, (t, pub_flag, "flag raised twice")
                        flag_owner[pub_flag] = t
                    cell_done.setdefault(it.pub_cell, []).append(t)
    for key, got in arrivals.items():
        assert got == need[key], (key, got, need[key])
    for (phase, ct), spans in cover.items():
        spans.sort()
        assert spans[0][0] == 0, (phase, ct, spans[:2])
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0], (phase, ct, a, b)  # contiguous, no overlap
    return dict(items=n, phases=len(phase_range), flags=len(flag_owner), phase_range=phase_range, phase_kind=phase_kind,
                phase_dep=phase_dep, slack={p: (min(v), sum(v) / len(v)) for p, v in slack.items()})
