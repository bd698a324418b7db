"""The drop-in proven under the reference's OWN caller: the reference's gpt-fast `Transformer` (gpt-fast/model.py,
unmodified -- taken from oracle/_ref, which oracle/fetch_ref.py fills from /root/reference, or from /root/reference
itself) with only `kernels.sparse_gemv` replaced by a module backed by teal_b200.ops.  Its `apply_monkeypatch`
forwards (model.py:163-190,222-224,258-272) then call OUR torch.ops.teal.sparse_gemv / sparse_qkv_gemv:
  (i)  eager decode against OracleDecoder,
  (ii) torch.compile(mode="reduce-overhead", fullgraph=True) of the decode step, as generate.py:420 does.
Skipped when neither copy of the reference is present."""
import importlib.util
import os
import sys
import types

import pytest
import torch

from oracle import teal_oracle as O
from test_decoder_gpu import CFG, THR, _state_dict

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_model_module():
    for base in (os.path.join(ROOT, "oracle", "_ref"), "/root/reference"):
        path = os.path.join(base, "gpt-fast", "model.py")
        if os.path.exists(path):
            break
    else:
        pytest.skip("reference gpt-fast/model.py not present (oracle/_ref not fetched)")
    import teal_b200.ops as ops
    # the ONE replaced module: kernels.sparse_gemv (model.py:16 `from kernels.sparse_gemv import SparseGEMV`)
    pkg = types.ModuleType("kernels")
    pkg.__path__ = []
    mod = types.ModuleType("kernels.sparse_gemv")
    mod.SparseGEMV, mod.SparseQKVGEMV = ops.SparseGEMV, ops.SparseQKVGEMV
    mod.splitk_sparse_gemv, mod.qkv_gemv = ops.splitk_sparse_gemv, ops.qkv_gemv
    saved = {k: sys.modules.get(k) for k in ("kernels", "kernels.sparse_gemv")}
    sys.modules["kernels"], sys.modules["kernels.sparse_gemv"] = pkg, mod
    try:
        spec = importlib.util.spec_from_file_location("reference_gptfast_model", path)
        m = importlib.util.module_from_spec(spec)
        sys.modules["reference_gptfast_model"] = m   # (torch.compile re-imports the module of the code it traces by name)
        spec.loader.exec_module(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return m


def _build(ref):
    from teal_b200.gptfast_compat import install_sparsity
    cfg = ref.ModelArgs(**{k: v for k, v in CFG.items()})
    model = ref.Transformer(cfg)
    model.load_state_dict(_state_dict(CFG), assign=True)
    model = model.to("cuda", torch.float16)
    with torch.device("cuda"):
        model.setup_caches(max_batch_size=1, max_seq_length=64)
    # generate.py:289-321 (monkeypatch_layer): ops installed, weights re-laid out column-major, forwards swapped
    return install_sparsity(model, THR)


def _check_against_oracle(step):
    orc = O.OracleDecoder(CFG, _state_dict(CFG), THR, max_seq=64)
    toks = [3, 17, 42, 8, 77, 300]
    for i, t in enumerate(toks):
        with torch.no_grad():
            got = step(torch.tensor([[t]], device="cuda"), torch.tensor([i], device="cuda")).float().cpu().reshape(-1)
            want = orc.forward(torch.tensor([[t]]), torch.tensor([i])).float().reshape(-1)
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 6e-2 * scale + 2e-2, (i, float((got - want).abs().max()), scale)


def test_reference_transformer_decodes_through_our_ops_eager():
    ref = _reference_model_module()
    model = _build(ref)
    assert type(model.layers[0].feed_forward.gemv1).__name__ == "OpOverloadPacket"  # torch.ops.teal.sparse_gemv
    _check_against_oracle(lambda x, pos: model(x, pos))


def test_reference_transformer_under_torch_compile_fullgraph():
    """generate.py:420: decode_one_token = torch.compile(decode_one_token, mode="reduce-overhead", fullgraph=True)."""
    ref = _reference_model_module()
    model = _build(ref)

    def decode_one_token(x, input_pos):
        return model(x, input_pos)

    compiled = torch.compile(decode_one_token, mode="reduce-overhead", fullgraph=True)
    _check_against_oracle(lambda x, pos: compiled(x, pos).clone())
