"""Step program (one persistent data-flow kernel per token): CPU checks of the native planner.

The planner (teal_step_program_plan) and the ticket -> item arithmetic (teal_step_program_item) are host code shared
with the kernel (csrc/step_program.cuh is __host__ __device__), so the schedule the GPU will run is replayed here
without a GPU: dependencies only point backwards in ticket order (=> the kernel's spin-waits cannot deadlock), every
tile gets exactly its arrivals, K ranges are covered exactly once, flags are unique.  GPU parity of the kernel itself
is in tests/test_step_program_gpu.py.
"""
import ctypes

import pytest

from teal_b200 import _lib
from teal_b200.program import DEP_ALL, DEP_LANES, DEP_QKV_HEAD, StepProgram, layer_desc, verify_schedule

FAKE = 0x7F0000000000  # plan() never dereferences device pointers: any 256-byte aligned non-null address will do


def _fake_desc(dim, n_head, n_kv, hd, inter, vocab, with_embed=True, with_head=True, **tuning):
    d = _lib.StepDesc()
    d.dim, d.n_head, d.n_kv_head, d.head_dim, d.inter, d.vocab, d.max_seq = dim, n_head, n_kv, hd, inter, vocab, 256
    d.eps = 1e-5
    a = [FAKE + 0x1000000 * i for i in range(1, 20)]
    d.tok_emb = a[0] if with_embed else None
    d.token, d.ssq_init = a[1], a[2]
    d.lm_head = a[3] if with_head else None
    d.final_norm, d.ld_head, d.freqs_cis, d.pos, d.out_tokens, d.argmax_packed = a[4], dim, a[5], a[6], a[7], a[8]
    d.x, d.qkv, d.att, d.h, d.logits, d.scratch = a[9], a[10], a[11], a[12], a[13], a[14]
    d.out_cap = 64
    for k, v in tuning.items():
        setattr(d, k, v)
    return d


def _fake_layer(i, dim, n_head, n_kv, hd, inter, thr=None, attn=True, mlp=True):
    base = FAKE + 0x100000000 * (i + 1)
    nqkv = (n_head + 2 * n_kv) * hd
    kw = dict(attn_norm=base + 0x10, ffn_norm=base + 0x20000, thr=thr)
    if attn:
        kw.update(wqkv=base + 0x1000000, wo=base + 0x2000000, k_cache=base + 0x3000000, v_cache=base + 0x4000000,
                  ld_qkv=nqkv, ld_o=dim)
    if mlp:
        kw.update(w1=base + 0x5000000, w3=base + 0x6000000, w2=base + 0x7000000, ld_1=inter, ld_3=inter, ld_2=dim)
    return layer_desc(**kw)


SHAPES = {
    "tiny": dict(dim=256, n_head=4, n_kv=2, hd=64, inter=704, vocab=512),
    "llama-2-7b": dict(dim=4096, n_head=32, n_kv=32, hd=128, inter=11008, vocab=32000),
    "llama-3-8b": dict(dim=4096, n_head=32, n_kv=8, hd=128, inter=14336, vocab=128256),
    "llama-2-70b": dict(dim=8192, n_head=64, n_kv=8, hd=128, inter=28672, vocab=32000),
}
UNIFORM = dict(q=0.3, k=0.3, v=0.3, o=0.02, gate=0.4, up=0.4, down=0.01)
GREEDY = dict(q=0.31, k=0.2, v=0.1, o=0.02, gate=0.4, up=0.35, down=0.01)


def _plan(shape, n_layer, thr=UNIFORM, **kw):
    s = SHAPES[shape]
    from teal_b200.program import TUNING_KEYS, set_knob
    tuning = {k: kw.pop(k) for k in list(kw) if k in TUNING_KEYS}
    d = _fake_desc(**s, **kw)
    for k, v in tuning.items():
        set_knob(d, k, v)
    layers = [_fake_layer(i, s["dim"], s["n_head"], s["n_kv"], s["hd"], s["inter"], thr) for i in range(n_layer)]
    return StepProgram(d, layers).plan()


@pytest.mark.parametrize("shape,n_layer,thr", [("tiny", 2, UNIFORM), ("tiny", 2, GREEDY), ("llama-2-7b", 1, UNIFORM),
                                               ("llama-3-8b", 2, UNIFORM), ("llama-3-8b", 1, GREEDY),
                                               ("llama-2-70b", 1, UNIFORM)])
def test_schedule_invariants(shape, n_layer, thr):
    prog = _plan(shape, n_layer, thr)
    st = verify_schedule(prog)
    s = SHAPES[shape]
    assert st["phases"] == 1 + 5 * n_layer + 2
    kinds = [st["phase_kind"][p] for p in sorted(st["phase_kind"])]
    assert kinds == [0] + [1, 2, 1, 1, 1] * n_layer + [3, 4]
    deps = [st["phase_dep"][p] for p in sorted(st["phase_dep"])]
    assert deps == [0] + [DEP_ALL, DEP_QKV_HEAD, DEP_LANES, DEP_ALL, DEP_LANES] * n_layer + [DEP_ALL, DEP_ALL]
    # one attention item per query head, one bookkeeping item, lm_head rows all covered
    for p, k in st["phase_kind"].items():
        lo, hi = st["phase_range"][p]
        if k == 2:
            assert hi - lo + 1 == s["n_head"]
        if k == 3:
            assert (hi - lo + 1) * 32 >= s["vocab"] > (hi - lo) * 32
        if k in (0, 4):
            assert hi == lo


def test_dataflow_edges_open_before_the_producer_phase_ends():
    """The point of the ordering (tile groups, k/v tiles first): the first down-projection items and the first
    attention heads depend only on tiles that complete well before their producer phase has issued all its tickets."""
    prog = _plan("llama-3-8b", 1)
    st = verify_schedule(prog)
    ranges, kinds = st["phase_range"], st["phase_kind"]
    # phases: 0 embed, 1 qkv, 2 attn, 3 o, 4 gate/up, 5 down, 6 head, 7 advance
    first_down = prog.item(ranges[5][0])
    latest = max(_owner(prog, f, ranges[4]) for f in list(first_down.deps)[: first_down.n_deps])
    lo, hi = ranges[4]
    assert (latest - lo) / (hi - lo + 1) < 0.35, "first h tile pair should complete within the first third of gate/up"
    first_head = prog.item(ranges[2][0])
    latest = max(_owner(prog, f, ranges[1]) for f in list(first_head.deps)[: first_head.n_deps])
    lo, hi = ranges[1]
    assert (latest - lo) / (hi - lo + 1) < 0.6
    assert kinds[5] == 1 and kinds[2] == 2


def _owner(prog, flag, rng):
    """ticket that raises `flag` (search the producer phase's ticket range)."""
    seen = {}
    for t in range(rng[0], rng[1] + 1):
        it = prog.item(t)
        key = it.pub_index
        seen[key] = seen.get(key, 0) + 1
        if seen[key] == it.arrivals and it.pub_flag == flag:
            return t
    raise AssertionError(f"flag {flag} not raised in {rng}")


@pytest.mark.parametrize("tuning", [dict(rows_small=32, rows_large=64), dict(rows_small=128, rows_large=256),
                                    dict(group_qkv=1, group_gate_up=2), dict(group_qkv=1000, group_gate_up=1000),
                                    dict(wide_melems=1), dict(wide_melems=100000), dict(head_rows=8),
                                    dict(group_gate_up=7), dict(warps=2), dict(warps=2, rows_small=64, rows_large=64),
                                    dict(warps=2, rows_qkv=32, rows_o=64, rows_gate_up=128, rows_down=192),
                                    dict(warps=4, rows_down=256, stagger_ns=200, opts=3), dict(opts=4), dict(warps=2, opts=7),
                                    dict(tiles_per_item=2), dict(tiles_per_item=4, warps=2, rows_large=64),
                                    dict(tiles_per_item=3, group_qkv=5, group_gate_up=6), dict(tiles_per_item=2, wide_melems=1)])
def test_tuning_knobs_keep_the_invariants(tuning):
    verify_schedule(_plan("tiny", 2, GREEDY, **tuning))
    verify_schedule(_plan("llama-2-7b", 1, UNIFORM, **tuning))


def test_partial_programs():
    """MLP-only / attention-only / no-embed / no-head programs (the pieces the GPU tests compare against the
    launch-per-projection kernels)."""
    s = SHAPES["tiny"]
    for attn, mlp, emb, head in [(True, False, True, False), (False, True, False, False), (True, True, False, True),
                                 (False, True, True, True)]:
        d = _fake_desc(**s, with_embed=emb, with_head=head)
        layers = [_fake_layer(i, s["dim"], s["n_head"], s["n_kv"], s["hd"], s["inter"], UNIFORM, attn=attn, mlp=mlp)
                  for i in range(2)]
        st = verify_schedule(StepProgram(d, layers).plan())
        assert st["phases"] == int(emb) + 2 * (3 * int(attn) + 2 * int(mlp)) + 2 * int(head)


def test_rejects_bad_descriptions():
    s = SHAPES["tiny"]
    d = _fake_desc(**s)
    lay = [_fake_layer(0, s["dim"], s["n_head"], s["n_kv"], s["hd"], s["inter"])]
    d.head_dim = 96
    with pytest.raises(_lib.TealLibraryError):
        StepProgram(d, lay)
    d = _fake_desc(**s)
    d.n_kv_head = 3
    with pytest.raises(AssertionError):
        StepProgram(d, lay)
    d = _fake_desc(**s, rows_small=6)
    with pytest.raises(AssertionError):
        StepProgram(d, lay)
    d = _fake_desc(**s, warps=3)
    with pytest.raises(_lib.TealLibraryError):
        StepProgram(d, lay)
    d = _fake_desc(**s, warps=2, rows_small=96)  # above 32 x warps it must be a multiple of 64
    with pytest.raises(AssertionError):
        StepProgram(d, lay)
    d = _fake_desc(**s)
    bad = _fake_layer(0, s["dim"], s["n_head"], s["n_kv"], s["hd"], s["inter"])
    bad.ld_qkv = 8  # < N
    p = StepProgram(d, [bad])
    with pytest.raises(AssertionError):
        p.plan()
    d = _fake_desc(**s)
    d.scratch = None
    with pytest.raises(AssertionError):
        StepProgram(d, lay)


def test_item_out_of_range():
    prog = _plan("tiny", 1)
    info = _lib.ItemInfo()
    assert prog.lib.teal_step_program_item(prog.host, prog.n_items, ctypes.byref(info)) < 0
    assert prog.lib.teal_step_program_item(prog.host, -1, ctypes.byref(info)) < 0
    junk = ctypes.create_string_buffer(4096)
    assert prog.lib.teal_step_program_items(junk) < 0
