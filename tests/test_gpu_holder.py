"""GPU tests of the TTSModelHolder mirror (crates/sbv2_core/src/tts.rs:40-349) above the C ABI:
registry, residency cap / eviction, .sbv2 and .aivmx loading, bert feature expansion
(tts_util.rs:129-154), sentence concatenation with 0.5 s silences and the WAV container."""
import struct

import numpy as np
import pytest
import torch

import util
from util import ov
from oracle import deberta as od

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S(lib_built):
    import sbv2_b200
    if sbv2_b200.device_count() < 1:
        pytest.fail("GPU tests selected but no B200 is visible")
    return sbv2_b200


@pytest.fixture(scope="module")
def blobs(S):
    from sbv2_b200 import assets
    hp = ov.tiny_hparams()
    bert_cfg = od.deberta_config(hidden_size=1024, num_hidden_layers=3, num_attention_heads=16, intermediate_size=256, vocab_size=300)
    bert = od.build_model(bert_cfg, seed=1)
    bert_onnx = assets.deberta_onnx(od.state_dict_numpy(bert))
    out = {"hp": hp, "bert_cfg": bert_cfg, "bert": bert, "bert_onnx": bert_onnx}
    rng = np.random.default_rng(0)
    for name, seed in (("a", 0), ("b", 5)):
        oracle = ov.build_model(hp, seed=seed)
        sd = ov.state_dict_numpy(oracle)
        sv = rng.standard_normal((3, 256)).astype(np.float32) * 0.1
        onnx = assets.synth_onnx(sd, hp.upsample_rates, hp.resblock_dilation_sizes)
        out[name] = dict(oracle=oracle, style=sv, onnx=onnx, sbv2=assets.sbv2_file(onnx, sv),
                         aivmx=assets.synth_onnx(sd, hp.upsample_rates, hp.resblock_dilation_sizes,
                                                 metadata=assets.aivmx_metadata(sv, fortran=True)))
    return out


def line(hp, t_x, seed):
    u = util.make_utterance(hp, t_x, seed)
    return dict(bert=u["bert"][0].numpy(), phones=u["x"][0].numpy(), tones=u["tone"][0].numpy(), lang_ids=u["lang"][0].numpy()), u


def wav_samples(wav):
    pos = wav.index(b"data")
    n, = struct.unpack("<I", wav[pos + 4:pos + 8])
    return np.frombuffer(wav[pos + 8:pos + 8 + n], dtype="<f4")


def test_registry_and_styles(S, blobs):
    h = S.TTSModelHolder(blobs["bert_onnx"], b"{}")
    assert h.models() == []
    h.load_sbv2file("a", blobs["a"]["sbv2"])
    h.load_sbv2file("a", blobs["a"]["sbv2"])          # same ident: ignored (tts.rs:156)
    h.load("b", __import__("sbv2_b200").assets.style_json(blobs["b"]["style"]), blobs["b"]["onnx"])
    assert h.models() == ["a", "b"] and h.loaded_count() == 2
    sv = blobs["a"]["style"]
    np.testing.assert_array_equal(h.get_style_vector("a", 2, 0.5), sv[0] + (sv[2] - sv[0]) * np.float32(0.5))
    with pytest.raises(S.Sbv2Error) as e:
        h.get_style_vector("zzz", 0, 1.0)
    assert e.value.status == S.ERR_MODEL_NOT_FOUND
    assert h.unload("a") is True and h.unload("a") is False
    assert h.models() == ["b"]
    h.close()


def test_aivmx_metadata_styles(S, blobs):
    h = S.TTSModelHolder(blobs["bert_onnx"], b"")
    h.load_aivmx("x", blobs["a"]["aivmx"])
    sv = blobs["a"]["style"]
    np.testing.assert_array_equal(h.get_style_vector("x", 1, 1.0), sv[0] + (sv[1] - sv[0]) * np.float32(1.0))
    # a graph without the metadata key is silently not registered (tts.rs:94)
    h.load_aivmx("y", blobs["a"]["onnx"])
    assert h.models() == ["x"]
    h.close()


def test_residency_cap_and_eviction(S, blobs):
    hp = blobs["hp"]
    h = S.TTSModelHolder(blobs["bert_onnx"], b"", max_loaded_models=1)
    h.load_sbv2file("a", blobs["a"]["sbv2"])
    h.load_sbv2file("b", blobs["b"]["sbv2"])   # cap reached: registered, not resident (tts.rs:157-163)
    assert h.models() == ["a", "b"] and h.loaded_count() == 1
    l0, _ = line(hp, 15, 3)
    wav_b = h.easy_synthesize("b", [l0], 0, 0)   # reloads b from retained bytes, evicts models[0]
    assert h.loaded_count() == 1 and "b" in h.models()
    assert wav_samples(wav_b).size % 512 == 0
    if "a" in h.models():
        wav_a = h.easy_synthesize("a", [l0], 0, 0)
        assert h.loaded_count() == 1 and wav_samples(wav_a).size % 512 == 0
    with pytest.raises(S.Sbv2Error) as e:
        h.easy_synthesize("nope", [l0], 0, 0)
    assert e.value.status == S.ERR_MODEL_NOT_FOUND
    h.close()


def test_easy_synthesize_concatenation(S, blobs):
    hp = blobs["hp"]
    h = S.TTSModelHolder(blobs["bert_onnx"], b"")
    h.load_sbv2file("a", blobs["a"]["sbv2"])
    l0, u0 = line(hp, 21, 7)
    l1, u1 = line(hp, 11, 8)
    style = h.get_style_vector("a", 1, 1.0)
    # expected lengths from the oracle (sdp_ratio 0: durations do not depend on the noise)
    n = []
    for u in (u0, u1):
        u = dict(u, style=torch.from_numpy(style)[None])
        _, inter = util.oracle_run(blobs["a"]["oracle"], u)
        n.append(512 * int(inter["y_lengths"][0]))
    wav = h.easy_synthesize("a", [l0, None, l1], 1, 0)       # "l0\n\nl1"
    x = wav_samples(wav)
    assert x.size == n[0] + 22050 + n[1]
    assert not x[n[0]:n[0] + 22050].any() and np.isfinite(x).all() and np.abs(x).max() > 1e-3
    wav2 = h.easy_synthesize("a", [l0, l1, None], 1, 0)      # trailing empty line: silence after both
    assert wav_samples(wav2).size == n[0] + 22050 + n[1] + 22050
    wav3 = h.easy_synthesize("a", [l1], 1, 0, length_scale=2.0)
    assert wav_samples(wav3).size > n[1]
    assert wav[:4] == b"RIFF" and struct.unpack("<I", wav[4:8])[0] == len(wav) - 8
    h.close()


def test_bert_features_expansion(S, blobs):
    cfg, hf = blobs["bert_cfg"], blobs["bert"]
    h = S.TTSModelHolder(blobs["bert_onnx"], b"")
    ids = np.array([1, 17, 33, 5, 250, 2], np.int64)      # CLS=1 ... SEP=2 (tokenizer.rs:9-21)
    word2ph = np.array([3, 4, 2, 0, 4, 2], np.int32)       # first entry odd as tts_util.rs:109-112 builds it
    got = h.bert_features(ids, np.ones_like(ids), word2ph)
    ref = od.predict(hf, torch.from_numpy(ids)[None], torch.ones(1, len(ids), dtype=torch.long))[0].numpy()
    want = np.repeat(ref, word2ph, axis=0).T               # [1024, sum(word2ph)]
    assert got.shape == want.shape == (1024, int(word2ph.sum()))
    assert np.abs(got - want).max() <= 3e-2
    # the expansion itself is exact: repeated columns are bit-identical
    assert np.array_equal(got[:, 0], got[:, 2]) and np.array_equal(got[:, 3], got[:, 6])
    h.close()


def test_synthesize_from_tokens_matches_host_expansion(S):
    """bert::predict + word2ph repeat + synthesize with the BERT features kept on the device is
    bit-identical to the three separate calls (same generator state), including zero-length word2ph entries."""
    import torch
    from oracle import deberta as od
    from oracle import vits as ov
    from sbv2_b200 import assets
    import util
    hp = ov.tiny_hparams(bert_dim=128)
    _, onnx = util.synth_assets(hp, seed=0)
    synth = S.Model(onnx, bert=False)
    cfg = od.tiny_config()
    bert = S.Model(assets.deberta_onnx(od.state_dict_numpy(od.build_model(cfg, seed=1))), bert=True)
    assert bert.hidden_size() == hp.bert_dim
    word2ph = np.array([1, 2, 3, 0, 2, 4, 1, 2, 4], np.int32)  # sum = 19 (odd, as the frontend's blank interleave gives)
    t_tok, t_x = len(word2ph), int(word2ph.sum())
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(3, cfg.vocab_size, (t_tok,), generator=g).numpy()
    mask = np.ones(t_tok, np.int64)
    x, tone, lang, _, style = ov.synthetic_inputs(hp, t_x, seed=11)
    x, tone, lang, style = x[0].numpy(), tone[0].numpy(), lang[0].numpy(), style[0].numpy()
    feats = bert.predict(ids, mask)                          # [t_tok, 128]
    expanded = np.repeat(feats, word2ph, axis=0).T.copy()    # tts_util.rs:129-154 -> [128, t_x]
    synth.seed(123)
    a = synth.synthesize(expanded, x, [0], tone, lang, style, 0.2, 1.0, 0.677, 0.8).reshape(-1)
    synth.seed(123)
    b = synth.synthesize_from_tokens(bert, ids, mask, word2ph, x, 0, tone, lang, style, 0.2, 1.0, 0.677, 0.8)
    assert a.shape == b.shape and a.size > 0
    assert np.array_equal(a, b)
    # error paths: word2ph that does not cover the phonemes, models swapped
    with pytest.raises(S.Sbv2Error):
        synth.synthesize_from_tokens(bert, ids, mask, word2ph[::-1].copy() * 0 + 1, x, 0, tone, lang, style, 0.2, 1.0, 0.677, 0.8)
    with pytest.raises(S.Sbv2Error):
        bert.synthesize_from_tokens(synth, ids, mask, word2ph, x, 0, tone, lang, style, 0.2, 1.0, 0.677, 0.8)


def _token_setup(S):
    from sbv2_b200 import assets
    hp = ov.tiny_hparams(bert_dim=128)
    oracle, onnx = util.synth_assets(hp, seed=0)
    cfg = od.tiny_config()
    bert_onnx = assets.deberta_onnx(od.state_dict_numpy(od.build_model(cfg, seed=1)))
    return hp, oracle, onnx, cfg, bert_onnx


def _sentence(hp, cfg, t_x, seed):
    w2p = util.word2ph_for(t_x, seed)
    ids = np.random.default_rng(seed).integers(3, cfg.vocab_size, size=w2p.size).astype(np.int64)
    x, tone, lang, _, style = ov.synthetic_inputs(hp, t_x, seed=seed)
    return dict(token_ids=ids, word2ph=w2p, x_tst=x[0].numpy(), tones=tone[0].numpy(), lang_ids=lang[0].numpy(), style_vec=style[0].numpy())


def test_synthesize_from_tokens_batch_matches_singles_and_writes_pauses(S):
    """sbv2_synthesize_from_tokens_batch: one DeBERTa batch + one synthesizer batch for the sentences of a request
    (SURVEY §8f row 2), pauses written on the device.  sdp_ratio = 0 so that durations do not depend on the noise: every
    sentence must have exactly the length of its own sbv2_synthesize_from_tokens call, the pauses must be exact zeros, and
    with batch 1 the audio is bit-identical to the single call (same generator state)."""
    hp, oracle, onnx, cfg, bert_onnx = _token_setup(S)
    synth, bert = S.Model(onnx, bert=False), S.Model(bert_onnx, bert=True)
    sents = [_sentence(hp, cfg, t, 40 + i) for i, t in enumerate((19, 7, 33, 11))]
    singles = []
    for s in sents:
        synth.seed(5)
        singles.append(synth.synthesize_from_tokens(bert, s["token_ids"], np.ones_like(s["token_ids"]), s["word2ph"], s["x_tst"], 0,
                                                    s["tones"], s["lang_ids"], s["style_vec"], 0.0, 1.0, 0.677, 0.8))
    pauses = [22050, 0, 22050, 100]
    synth.seed(5)
    audio, ns = synth.synthesize_from_tokens_batch(bert, sents, pauses)
    assert [int(n) for n in ns] == [a.size for a in singles]
    assert audio.size == int(ns.sum()) + sum(pauses)
    off = 0
    for i, n in enumerate(ns):
        seg = audio[off:off + n]
        assert np.isfinite(seg).all() and np.abs(seg).max() > 1e-4
        off += int(n)
        assert not audio[off:off + pauses[i]].any()      # silence written on the device
        off += pauses[i]
    synth.seed(5)
    one, n1 = synth.synthesize_from_tokens_batch(bert, sents[:1], [7])
    assert np.array_equal(one[:n1[0]], singles[0]) and not one[n1[0]:].any() and one.size == n1[0] + 7
    with pytest.raises(S.Sbv2Error):
        synth.synthesize_from_tokens_batch(bert, [dict(sents[0], word2ph=sents[0]["word2ph"] + 1)])


def test_easy_synthesize_tokens_matches_feature_path_lengths(S):
    hp, oracle, onnx, cfg, bert_onnx = _token_setup(S)
    from sbv2_b200 import assets
    sv = np.random.default_rng(3).standard_normal((2, 256)).astype(np.float32) * 0.1
    h = S.TTSModelHolder(bert_onnx, b"")
    h.load("m", assets.style_json(sv), onnx)
    sents = [_sentence(hp, cfg, t, 60 + i) for i, t in enumerate((15, 9))]
    lines_tok = [dict(token_ids=s["token_ids"], word2ph=s["word2ph"], phones=s["x_tst"], tones=s["tones"], lang_ids=s["lang_ids"]) for s in sents]
    # the reference-shaped path: bert_features per line on the host, then easy_synthesize
    lines_feat = []
    for s in sents:
        f = h.bert_features(s["token_ids"], np.ones_like(s["token_ids"]), s["word2ph"])
        assert f.shape == (128, s["x_tst"].size)     # rows follow the holder's DeBERTa hidden size, not a constant 1024
        lines_feat.append(dict(bert=f, phones=s["x_tst"], tones=s["tones"], lang_ids=s["lang_ids"]))
    a = wav_samples(h.easy_synthesize("m", [lines_feat[0], None, lines_feat[1]], 1, 0))
    b = wav_samples(h.easy_synthesize_tokens("m", [lines_tok[0], None, lines_tok[1]], 1, 0))
    assert a.size == b.size                              # same durations (sdp_ratio 0), same pause layout
    n0 = (a.size - 22050 - 0) // 1
    # locate the pause: identical position in both
    za = np.flatnonzero(np.convolve((a == 0).astype(np.int32), np.ones(22050, np.int32), "valid") == 22050)
    zb = np.flatnonzero(np.convolve((b == 0).astype(np.int32), np.ones(22050, np.int32), "valid") == 22050)
    assert za.size and zb.size and za[0] == zb[0]
    with pytest.raises(S.Sbv2Error) as e:
        h.easy_synthesize_tokens("m", [None, None], 1, 0)  # only empty lines: concatenate of nothing (tts.rs:321-324)
    assert e.value.status == S.ERR_INVALID_ARGUMENT
    with pytest.raises(S.Sbv2Error) as e:
        h.easy_synthesize("m", [None], 1, 0)
    assert e.value.status == S.ERR_INVALID_ARGUMENT
    h.close()


def test_pinned_inputs_give_identical_results(S):
    """Inputs in page-locked memory (sbv2_alloc_pinned) are copied by the DMA engine in place; results are bit-identical
    to the staged path."""
    hp = ov.tiny_hparams()
    oracle, onnx = util.synth_assets(hp, seed=0)
    model = S.Model(onnx, bert=False)
    us = [util.to_api(util.make_utterance(hp, t, seed=70 + i, sdp_ratio=0.3)) for i, t in enumerate((23, 5, 41))]
    a = model.synthesize_batch(us)
    pinned = [dict(u, bert=S.pinned_copy(u["bert"])) for u in us]
    b = model.synthesize_batch(pinned)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    db = S.DeviceBatch(model, pinned)
    for p in pinned:
        p["bert"][...] = 0                               # upload has finished with the caller's buffers when it returns
    db.run()
    for x, y in zip(a, db.download()):
        assert np.array_equal(x, y)


def test_anonymous_exports_load_and_match(S):
    """Models written the way TorchScript + onnxsim leave them (anonymous transposed Linear weights, anonymous
    weight-normed convs; convert_model.py:115-156, convert_deberta.py:36-52) load through the structural binder and give
    bit-identical results; a Linear whose weight is missing is a load error, not a silently skipped layer."""
    from sbv2_b200 import assets
    for flow in (True, False):
        hp = ov.tiny_hparams(use_transformer_flow=flow)
        sd = ov.state_dict_numpy(ov.build_model(hp, seed=2))
        named = S.Model(assets.synth_onnx(sd, hp.upsample_rates, hp.resblock_dilation_sizes), bert=False)
        anon = S.Model(assets.synth_onnx(sd, hp.upsample_rates, hp.resblock_dilation_sizes, anonymize_weight_norm=True,
                                         anonymize_linear=True), bert=False)
        assert anon.describe()["structural_binding"] is True
        u = util.make_utterance(hp, 27, seed=9, sdp_ratio=0.3, sid=1)
        args = (u["bert"][0].numpy(), u["x"][0].numpy(), u["sid"], u["tone"][0].numpy(), u["lang"][0].numpy(), u["style"][0].numpy(),
                u["sdp_ratio"], u["length_scale"], u["noise_scale"], u["noise_scale_w"], u["noise_sdp"][0].numpy(), u["noise_zp"][0].numpy())
        a, b = named.synthesize_with_noise(*args), anon.synthesize_with_noise(*args)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    broken = {k: v for k, v in sd.items() if k != "enc_p.encoder.spk_emb_linear.weight"}
    with pytest.raises(S.Sbv2Error) as e:
        S.Model(assets.synth_onnx(broken, hp.upsample_rates, hp.resblock_dilation_sizes), bert=False)
    assert e.value.status == S.ERR_UNSUPPORTED and "spk_emb_linear" in e.value.message
    cfg = od.tiny_config()
    bsd = od.state_dict_numpy(od.build_model(cfg, seed=1))
    bn, ba = S.Model(assets.deberta_onnx(bsd), bert=True), S.Model(assets.deberta_onnx(bsd, anonymize_linear=True), bert=True)
    assert ba.describe()["structural_binding"] is True and bn.describe()["structural_binding"] is False
    ids = np.arange(3, 40, dtype=np.int64)
    assert np.array_equal(bn.predict(ids, np.ones_like(ids)), ba.predict(ids, np.ones_like(ids)))
    with pytest.raises(S.Sbv2Error) as e:
        S.Model(assets.deberta_onnx({k: v for k, v in bsd.items() if k != "deberta.encoder.layer.0.output.dense.weight"}), bert=True)
    assert e.value.status == S.ERR_UNSUPPORTED


def test_second_device_in_one_process(S):
    """Function attributes (dynamic shared memory opt-in) are per device: a model on device 1 must work after device 0
    was used in the same process.  Needs two GPUs; on a one-GPU box only the error path is checked."""
    hp = ov.tiny_hparams()
    oracle, onnx = util.synth_assets(hp, seed=0)
    u = util.to_api(util.make_utterance(hp, 23, seed=12, sdp_ratio=0.4))
    m0 = S.Model(onnx, bert=False, device=0)
    a = m0.synthesize_batch([u])[0]
    if S.device_count() < 2:
        with pytest.raises(S.Sbv2Error):
            S.Model(onnx, bert=False, device=1)
        return
    m1 = S.Model(onnx, bert=False, device=1)
    b = m1.synthesize_batch([u])[0]
    assert np.array_equal(a, b)
