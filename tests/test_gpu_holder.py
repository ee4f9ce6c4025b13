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
