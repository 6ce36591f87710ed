"""Host-side logic that needs no GPU: parameter store layout, shapes, scopes."""
import numpy as np
import pytest
import torch

from tacotron_b200.params import ParamStore, model_shapes, cbhg_shapes
from tacotron_b200.models.tacotron import Config
from oracle import tacotron_oracle as O


def test_store_matches_oracle_inventory():
    cfg = Config(r=5, vocab_size=64)
    mine = {n: tuple(s) for n, s, _ in model_shapes(cfg)}
    theirs = {n: tuple(s) for n, s, _ in O.param_shapes(O.OracleConfig(r=5, vocab_size=64))}
    assert mine == theirs


def test_store_views_alias_flat_and_bank_is_contiguous():
    cfg = Config(r=2, vocab_size=20)
    st = ParamStore(model_shapes(cfg), "cpu")
    st.init_tf_default(3)
    span = st.span("enc/cbhg/bank/W1", "enc/cbhg/bank/W16")
    assert span.numel() == 128 * 128 * (16 * 17 // 2)
    off = 0
    for k in range(1, 17):
        w = st[f"enc/cbhg/bank/W{k}"]
        assert torch.equal(span[off:off + w.numel()].view_as(w), w)
        off += w.numel()
    b = st.span("post/cbhg/bank/b1", "post/cbhg/bank/b8")
    assert b.numel() == 8 * 128
    for n in st.names():
        assert st[n].data_ptr() % 16 == 0
    assert float(st["dec/gru1/bg"].min()) == 1.0 and float(st["enc/cbhg/gru_fw/bc"].abs().max()) == 0.0


def test_load_roundtrip_and_version():
    cfg_o = O.OracleConfig(r=2, vocab_size=20)
    p = O.init_params(cfg_o, 1, True)
    st = ParamStore(model_shapes(Config(r=2, vocab_size=20)), "cpu")
    v0 = st.version
    st.load(p)
    assert st.version == v0 + 1
    for n, t in p.items():
        assert torch.equal(st[n], t)
    with pytest.raises(KeyError):
        st.load({})


def test_scopes():
    from tacotron_b200.models import ops
    st = ParamStore(model_shapes(Config(r=2, vocab_size=20)), "cpu")
    with ops.variable_scope(st, "enc") as sc:
        assert sc.name("prenet/W1") == "enc/prenet/W1"
        with ops.variable_scope("cbhg") as sc2:
            assert sc2.p("bank/W3").shape == (3, 128, 128)
    with pytest.raises(RuntimeError):
        ops.current_scope()


def test_model_requires_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tacotron_b200 import Tacotron
    with pytest.raises(RuntimeError):
        Tacotron(Config(), None, train=False)


def test_tf_name_map_covers_every_parameter():
    """Every TF variable the reference's graph code creates (recorded while executing it, tests/golden) maps to
    exactly one parameter of the store, and every parameter is hit."""
    import os
    from tacotron_b200.tf_names import tf_name_to_param
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_wiring_r5.npz"))
    names = [str(x) for x in g["tf_variable_names"]]
    mapped = [tf_name_to_param(n) for n in names]
    store_names = {n for n, _, _ in model_shapes(Config(r=5, vocab_size=20))}
    assert len(set(mapped)) == len(mapped) == len(store_names) and set(mapped) == store_names
    with pytest.raises(KeyError):
        tf_name_to_param("global_step")
