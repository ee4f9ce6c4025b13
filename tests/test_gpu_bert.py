"""GPU parity of the DeBERTa-v2 feature encoder (bert::predict, crates/sbv2_core/src/bert.rs:6-24)
against the HF model the reference exports (oracle/deberta.py).

Two numerics modes (bert_model.cu): "exact" (default; two-term fp16 operand splits, fp32 activations and attention) must
match HF fp32 to max-abs 2e-4 / relative Frobenius 2e-5 — the features feed ceil() downstream; "fp16"
(SBV2_B200_BERT=fp16; single-term fp16 operands, tensor-core attention) to 3e-2 / 5e-3 on activations of magnitude ~4."""
import os

import numpy as np
import pytest
import torch

import util  # noqa: F401  (sys.path)
from oracle import deberta as od

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def S(lib_built):
    import sbv2_b200
    if sbv2_b200.device_count() < 1:
        pytest.fail("GPU tests selected but no B200 is visible")
    return sbv2_b200


@pytest.fixture(scope="module")
def tiny(S):
    from sbv2_b200 import assets
    cfg = od.tiny_config()
    hf = od.build_model(cfg, seed=1)
    model = S.Model(assets.deberta_onnx(od.state_dict_numpy(hf)), bert=True)
    assert model.describe()["numerics"] == "exact"
    return cfg, hf, model


@pytest.fixture(scope="module")
def tiny_fp16(S):
    from sbv2_b200 import assets
    cfg = od.tiny_config()
    hf = od.build_model(cfg, seed=1)
    model = make_model(S, assets.deberta_onnx(od.state_dict_numpy(hf)), "fp16")
    assert model.describe()["numerics"] == "fp16"
    return cfg, hf, model


TOL = {"exact": (2e-4, 2e-5), "fp16": (3e-2, 5e-3)}


def close(got, ref, mode="exact"):
    err = np.abs(got - ref).max()
    rel = np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30)
    assert err <= TOL[mode][0] and rel <= TOL[mode][1], f"[{mode}] max-abs {err:.3e}, rel-fro {rel:.3e} (|ref| max {np.abs(ref).max():.2f})"


def make_model(S, onnx, mode):
    if mode == "exact":
        return S.Model(onnx, bert=True)
    os.environ["SBV2_B200_BERT"] = "fp16"
    try:
        return S.Model(onnx, bert=True)
    finally:
        del os.environ["SBV2_B200_BERT"]


def test_describe_and_live_layers(tiny):
    cfg, hf, model = tiny
    d = model.describe()
    assert d["kind"] == "deberta-v2" and d["hidden_size"] == cfg.hidden_size
    assert d["num_hidden_layers"] == cfg.num_hidden_layers and d["live_layers"] == cfg.num_hidden_layers - 2
    assert model.hidden_size() == cfg.hidden_size


@pytest.mark.parametrize("s", [1, 7, 64, 129, 200, 300])
def test_predict_matches_hf(tiny, s):
    """cfg 1 uses T_tok = 7; s > 129 exercises the log-bucket region of the relative positions."""
    cfg, hf, model = tiny
    g = torch.Generator().manual_seed(100 + s)
    ids = torch.randint(3, cfg.vocab_size, (1, s), generator=g)
    ref = od.predict(hf, ids, torch.ones_like(ids))[0].numpy()
    got = model.predict(ids[0].numpy(), np.ones(s, np.int64))
    assert got.shape == (s, cfg.hidden_size)
    close(got, ref)


@pytest.mark.parametrize("s", [1, 7, 64, 128, 129, 200, 256, 257, 300, 384, 385, 511, 512])
def test_predict_matches_hf_fp16_mode(tiny_fp16, s):
    """The throughput mode: single-tile tensor-core attention up to 128 tokens, the multi-tile kernel (online softmax,
    log-bucket position windows) from 129 to 512."""
    cfg, hf, model = tiny_fp16
    g = torch.Generator().manual_seed(100 + s)
    ids = torch.randint(3, cfg.vocab_size, (1, s), generator=g)
    ref = od.predict(hf, ids, torch.ones_like(ids))[0].numpy()
    got = model.predict(ids[0].numpy(), np.ones(s, np.int64))
    close(got, ref, "fp16")


def test_golden_fixture(tiny):
    cfg, hf, model = tiny
    g = np.load(os.path.join(GOLDEN, "deberta_tiny_s7.npz"))
    got = model.predict(g["ids"][0], np.ones(7, np.int64))
    close(got, g["out"])


def test_batch_with_right_padding_equals_singles(tiny):
    cfg, hf, model = tiny
    g = torch.Generator().manual_seed(5)
    S_, lens = 40, [40, 17, 1, 33]
    ids = torch.randint(3, cfg.vocab_size, (len(lens), S_), generator=g).numpy()
    mask = np.zeros_like(ids)
    for b, l in enumerate(lens):
        mask[b, :l] = 1
    out = model.predict_batch(ids, mask)
    assert out.shape == (len(lens), S_, cfg.hidden_size)
    for b, l in enumerate(lens):
        single = model.predict(ids[b, :l], np.ones(l, np.int64))
        assert np.array_equal(out[b, :l], single)       # bit-identical to the batch-1 call
        assert not out[b, l:].any()                     # padded positions are zeros
        ref = od.predict(hf, torch.from_numpy(ids[b:b + 1, :l]), torch.ones(1, l, dtype=torch.long))[0].numpy()
        close(single, ref)


def test_error_paths(tiny, S):
    cfg, hf, model = tiny
    with pytest.raises(S.Sbv2Error):
        model.predict([cfg.vocab_size + 5], [1])
    with pytest.raises(S.Sbv2Error):
        model.predict_batch(np.ones((1, 4), np.int64), np.array([[1, 0, 1, 1]]))  # hole in the mask
    with pytest.raises(S.Sbv2Error):
        model.synthesize(np.zeros((1024, 3), np.float32), [0, 1, 0], [0], [0, 6, 0], [0, 1, 0], np.zeros(256, np.float32),
                         0.0, 1.0, 0.677, 0.8)
    assert model.predict([5, 6, 7], [1, 1, 1]).shape == (3, cfg.hidden_size)


@pytest.mark.parametrize("mode,smax,lens", [("fp16", 128, [128, 1, 77, 128, 5, 100]), ("exact", 128, [128, 1, 77, 128, 5, 100]),
                                            ("fp16", 512, [512, 129, 300, 1, 128, 257, 400, 385, 256, 511]),
                                            ("exact", 512, [512, 129, 300, 1, 128, 257, 400, 385, 256, 511])])
def test_tensor_core_attention_matches_cuda_core_attention(S, mode, smax, lens):
    """The disentangled attention runs on tcgen05 (bert_attention_tc.cu: one tile up to 128 tokens, 128 x 128 tile pairs
    with gathered log-bucket position windows up to 512); the CUDA-core kernel (SBV2_B200_BERT_ATTN=simt) is the
    cross-check, on a ragged right-padded batch.  Exact mode (the default) has its own tensor-core kernels (two-term fp16
    splits of every operand, fp32 bias tile), again one tile up to 128 tokens and tile pairs beyond."""
    from sbv2_b200 import assets
    cfg = od.tiny_config()
    onnx = assets.deberta_onnx(od.state_dict_numpy(od.build_model(cfg, seed=1)))
    tc = make_model(S, onnx, mode)
    os.environ["SBV2_B200_BERT_ATTN"] = "simt"
    try:
        simt = make_model(S, onnx, mode)
    finally:
        del os.environ["SBV2_B200_BERT_ATTN"]
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(3, cfg.vocab_size, (len(lens), smax), generator=g).numpy()
    mask = np.zeros((len(lens), smax), np.int64)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    a = tc.predict_batch(ids, mask)
    b_ = simt.predict_batch(ids, mask)
    assert a.shape == b_.shape == (len(lens), smax, cfg.hidden_size)
    assert np.isfinite(a).all()
    for i, n in enumerate(lens):
        assert not a[i, n:].any()
        close(a[i, :n], b_[i, :n], mode)
