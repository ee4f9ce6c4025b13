"""CPU tests of the oracle itself: integer alignment restatement vs the float generate_path,
batch-vs-single invariance, golden fixtures, SDP spline sanity."""
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import util
from util import ov

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@settings(max_examples=200, deadline=None)
@given(st.lists(st.floats(min_value=0.0, max_value=9.5, allow_nan=False, width=32), min_size=1, max_size=40))
def test_length_regulator_int_matches_generate_path(ws):
    w = torch.tensor(ws, dtype=torch.float32).view(1, 1, -1)
    w_ceil = torch.ceil(w)
    y_len = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
    x_mask = torch.ones_like(w)
    y_mask = ov.sequence_mask(y_len, None).unsqueeze(1).float()
    attn = ov.generate_path(w_ceil, x_mask.unsqueeze(2) * y_mask.unsqueeze(-1))[0, 0]  # [T_y, T_x]
    d, t_y, f2p = ov.length_regulate_int(np.asarray(ws, np.float32))
    assert t_y == int(y_len[0])
    assert np.array_equal(d, w_ceil[0, 0].numpy().astype(np.int32))
    rows = attn.sum(1).numpy()
    for j in range(t_y):
        if f2p[j] < 0:
            assert rows[j] == 0
        else:
            assert rows[j] == 1 and int(attn[j].argmax()) == f2p[j]


def test_zero_durations_clamp():
    d, t_y, f2p = ov.length_regulate_int(np.zeros(5, np.float32))
    assert t_y == 1 and f2p.tolist() == [-1] and d.sum() == 0


def test_oracle_batch_equals_single():
    """The graph is batch-1 in the reference; padding + masks must not change an utterance."""
    hp = ov.tiny_hparams()
    model = ov.build_model(hp, seed=0)
    us = [util.make_utterance(hp, t, seed=s, sdp_ratio=0.3) for t, s in ((21, 1), (33, 2))]
    singles = [util.oracle_run(model, u)[1] for u in us]
    T = max(u["t_x"] for u in us)
    pad = lambda t, n, dim: torch.nn.functional.pad(t, (0, n - t.shape[dim]))
    x = torch.cat([pad(u["x"], T, 1) for u in us])
    tone = torch.cat([pad(u["tone"], T, 1) for u in us])
    lang = torch.cat([pad(u["lang"], T, 1) for u in us])
    bert = torch.cat([pad(u["bert"], T, 2) for u in us])
    style = torch.cat([u["style"] for u in us])
    nsdp = torch.cat([pad(u["noise_sdp"], T, 2) for u in us])
    F = min(u["noise_zp"].shape[2] for u in us)
    nzp = torch.cat([u["noise_zp"][:, :, :F] for u in us])
    o, inter = model.infer(x, torch.tensor([u["t_x"] for u in us]), torch.tensor([0, 0]), tone, lang, bert, style,
                           noise_sdp=nsdp, noise_zp=nzp, noise_scale=0.677, noise_scale_w=0.8, sdp_ratio=0.3,
                           return_intermediates=True)
    for i, u in enumerate(us):
        wc = inter["w_ceil"][i, 0, :u["t_x"]]
        assert torch.equal(wc, singles[i]["w_ceil"][0, 0])
        ty = int(singles[i]["y_lengths"][0])
        assert torch.allclose(inter["z"][i, :, :ty], singles[i]["z"][0], atol=2e-5)


def test_spline_inverse_is_monotone_and_identity_in_tails():
    torch.manual_seed(0)
    n = 64
    uw, uh, ud = torch.randn(n, 10), torch.randn(n, 10), torch.randn(n, 9)
    x = torch.linspace(-7, 7, n)
    y = ov.unconstrained_rqs_inverse(x.clone(), uw, uh, ud, 5.0)
    assert torch.equal(y[x.abs() > 5], x[x.abs() > 5])
    # same parameters on every row -> monotone map
    uw1, uh1, ud1 = uw[:1].expand(n, -1), uh[:1].expand(n, -1), ud[:1].expand(n, -1)
    y1 = ov.unconstrained_rqs_inverse(x.clone(), uw1, uh1, ud1, 5.0)
    assert (y1[1:] >= y1[:-1] - 1e-5).all()
    assert y1.abs().max() <= 7.0 + 1e-4


@pytest.mark.parametrize("name", ["synth_tiny_tx23"])
def test_golden_fixture_reproduces(name):
    """tests/golden/*.npz were minted by tests/golden/make_golden.py from this oracle; a change in
    the oracle (or in torch's CPU kernels) that moves results shows up here."""
    path = os.path.join(GOLDEN, name + ".npz")
    g = np.load(path)
    hp = ov.tiny_hparams()
    model = ov.build_model(hp, seed=int(g["weights_seed"]))
    u = util.make_utterance(hp, int(g["t_x"]), seed=int(g["input_seed"]), sdp_ratio=float(g["sdp_ratio"]))
    o, inter = util.oracle_run(model, u)
    assert np.array_equal(inter["w_ceil"][0, 0].numpy().astype(np.int32), g["durations"])
    np.testing.assert_allclose(o[0, 0].numpy(), g["audio"], atol=2e-5)
    np.testing.assert_allclose(inter["logw"][0, 0].numpy(), g["logw"], atol=2e-5)


def test_c_restatement_of_length_regulator_matches_numpy():
    """oracle/length_regulator.c (plain C, integers) vs oracle/vits.py (numpy) vs float generate_path."""
    import ctypes as C
    import importlib.util
    spec = importlib.util.spec_from_file_location("oracle_build_c", os.path.join(util.ROOT, "oracle", "build_c.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = C.CDLL(mod.build())
    lib.sbv2_oracle_durations.restype = C.c_int32
    rng = np.random.default_rng(0)
    for t_x in (1, 2, 23, 241, 1801):
        for scale in (0.0, 0.3, 1.0, 2.5):
            w = (np.exp(rng.standard_normal(t_x)) * scale).astype(np.float32)
            d_ref, ty_ref, f_ref = ov.length_regulate_int(w)
            d = np.zeros(t_x, np.int32)
            ty = lib.sbv2_oracle_durations(w.ctypes.data_as(C.POINTER(C.c_float)), t_x, d.ctypes.data_as(C.POINTER(C.c_int32)))
            f = np.zeros(ty, np.int32)
            lib.sbv2_oracle_frame2ph(d.ctypes.data_as(C.POINTER(C.c_int32)), t_x, ty, f.ctypes.data_as(C.POINTER(C.c_int32)))
            assert ty == ty_ref and np.array_equal(d, d_ref) and np.array_equal(f, f_ref)
