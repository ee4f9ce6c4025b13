"""CPU tests: the C-ABI library loads and exports every symbol include/sbv2_b200.h declares, the
asset readers (.sbv2, style json, aivmx npy, WAV, ONNX) behave like the reference's, and the error
paths return statuses instead of aborting.  No compute call is made (no GPU here)."""
import base64
import io
import json
import os
import re
import struct
import tarfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def S(lib_built):
    import sbv2_b200
    return sbv2_b200


def test_header_symbols_exported(S):
    hdr = open(os.path.join(ROOT, "include", "sbv2_b200.h")).read()
    declared = set(re.findall(r"\b(sbv2_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"sbv2_status"}
    assert len(declared) > 30
    missing = [s for s in sorted(declared) if not hasattr(S.lib, s)]
    assert not missing, f"declared but not exported: {missing}"
    # and the binding knows all of them
    unbound = sorted(declared - set(S.EXPORTED_SYMBOLS))
    assert not unbound, f"exported but not bound in sbv2_b200/__init__.py: {unbound}"


def test_version_and_no_cpu_fallback(S):
    assert "sm_100a" in S.version()
    if S.device_count() == 0:
        from sbv2_b200 import assets
        onnx = assets.model_proto({"enc_p.emb.weight": np.zeros((4, 4), np.float32)})
        with pytest.raises(S.Sbv2Error) as e:
            S.Model(onnx, bert=False)
        assert e.value.status == S.ERR_CUDA  # fails loudly; never computes on the CPU


def test_sbv2file_roundtrip(S):
    from sbv2_b200 import assets
    sv = np.random.default_rng(0).standard_normal((4, 256)).astype(np.float32)
    onnx = assets.model_proto({"w": np.arange(12, dtype=np.float32).reshape(3, 4)}, metadata={"a": "b"})
    blob = assets.sbv2_file(onnx, sv)
    style_json, model_onnx = S.parse_sbv2file(blob)
    assert model_onnx == onnx
    assert json.loads(style_json)["shape"] == [4, 256]
    np.testing.assert_array_equal(S.load_style(style_json), sv)
    # unknown entries are ignored (sbv2file.rs:24-28)
    blob2 = assets.sbv2_file(onnx, sv, extra={"README": b"hello"})
    assert S.parse_sbv2file(blob2)[1] == onnx


def test_sbv2file_errors(S):
    from sbv2_b200 import assets
    sv = np.zeros((1, 256), np.float32)
    with pytest.raises(S.Sbv2Error) as e:
        S.parse_sbv2file(b"not zstd at all")
    assert e.value.status == S.ERR_PARSE
    with pytest.raises(S.Sbv2Error) as e:
        S.parse_sbv2file(assets.sbv2_file(b"x", sv, omit=["style_vectors.json"]))
    assert e.value.status == S.ERR_MODEL_NOT_FOUND and "style_vectors" in e.value.message
    with pytest.raises(S.Sbv2Error) as e:
        S.parse_sbv2file(assets.sbv2_file(b"x", sv, omit=["model.onnx"]))
    assert e.value.status == S.ERR_MODEL_NOT_FOUND and "vits2" in e.value.message
    # truncated frame
    good = assets.sbv2_file(b"x" * 1000, sv)
    with pytest.raises(S.Sbv2Error):
        S.parse_sbv2file(good[: len(good) // 2])
    with pytest.raises(S.Sbv2Error):
        S.parse_sbv2file(b"")


def test_style_vector_math(S):
    sv = np.random.default_rng(1).standard_normal((5, 256)).astype(np.float32)
    for sid, w in ((0, 1.0), (3, 0.7), (4, 0.0), (2, 2.5)):
        want = sv[0] + (sv[sid] - sv[0]) * np.float32(w)  # style.rs:24-27
        np.testing.assert_array_equal(S.get_style_vector(sv, sid, w), want)
    with pytest.raises(S.Sbv2Error):
        S.get_style_vector(sv, 5, 1.0)
    with pytest.raises(S.Sbv2Error):
        S.get_style_vector(sv, -1, 1.0)


def test_style_json_errors(S):
    with pytest.raises(S.Sbv2Error) as e:
        S.load_style(b'{"shape":[2,2],"data":[[1,2],[3]]}')
    assert e.value.status == S.ERR_INVALID_ARGUMENT  # ShapeError in the reference
    with pytest.raises(S.Sbv2Error) as e:
        S.load_style(b'{"shape":[2,2]}')
    assert e.value.status == S.ERR_PARSE
    with pytest.raises(S.Sbv2Error):
        S.load_style(b"[1,2,3]")
    m = S.load_style(b'{ "data" : [[1.5, -2e-1],[3,4]], "extra": {"x":[1,2]}, "shape":[2,2] }')
    np.testing.assert_array_equal(m, np.array([[1.5, -0.2], [3, 4]], np.float32))


def test_aivmx_style_npy(S):
    from sbv2_b200 import assets
    sv = np.random.default_rng(2).standard_normal((3, 256)).astype(np.float32)
    for fortran in (False, True):
        md = assets.aivmx_metadata(sv, fortran=fortran)
        np.testing.assert_array_equal(S.load_style_npy_base64(md["aivm_style_vectors"].encode()), sv)
    with pytest.raises(S.Sbv2Error):
        S.load_style_npy_base64(base64.b64encode(b"not an npy file"))
    with pytest.raises(S.Sbv2Error):
        S.load_style_npy_base64(b"***")
    b = io.BytesIO()
    np.save(b, np.zeros((2, 2, 2), np.float32))
    with pytest.raises(S.Sbv2Error):  # "expected 2D array" (tts.rs:103)
        S.load_style_npy_base64(base64.b64encode(b.getvalue()))


def test_wav_container(S):
    x = np.linspace(-1, 1, 1000, dtype=np.float32)
    w = S.wav_from_f32(x)
    assert w[:4] == b"RIFF" and w[8:12] == b"WAVE"
    riff_size, = struct.unpack("<I", w[4:8])
    assert riff_size == len(w) - 8
    fmt_tag, ch, sr, byte_rate, block, bits = struct.unpack("<HHIIHH", w[20:36])
    assert (fmt_tag, ch, sr, byte_rate, block, bits) == (0xFFFE, 1, 44100, 44100 * 4, 4, 32)  # tts_util.rs:164-169
    pos = w.index(b"data")
    n, = struct.unpack("<I", w[pos + 4:pos + 8])
    assert n == 4000
    np.testing.assert_array_equal(np.frombuffer(w[pos + 8:], dtype="<f4"), x)
    assert len(S.wav_from_f32(np.zeros(0, np.float32))) == 68


def test_wav_pcm16_container(S):
    """Optional 16-bit PCM container: canonical 44-byte header, clamp + round-to-nearest, readable by the
    standard library's wave module."""
    import io
    import wave
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.25, 2.0, -3.0, np.nan, 1e-6], np.float32)
    w = S.wav_pcm16_from_f32(x)
    assert len(w) == 44 + 2 * x.size and w[:4] == b"RIFF" and w[8:16] == b"WAVEfmt " and w[36:40] == b"data"
    assert struct.unpack("<I", w[4:8])[0] == len(w) - 8
    assert struct.unpack("<IHHIIHH", w[16:36]) == (16, 1, 1, 44100, 88200, 2, 16)
    got = np.frombuffer(w[44:], dtype="<i2")
    assert got.tolist() == [0, 32767, -32767, 16384, -8192, 32767, -32767, 0, 0]
    with wave.open(io.BytesIO(w)) as r:
        assert (r.getnchannels(), r.getsampwidth(), r.getframerate(), r.getnframes()) == (1, 2, 44100, x.size)
    assert len(S.wav_pcm16_from_f32(np.zeros(0, np.float32))) == 44


def test_onnx_reader_rejects_garbage(S):
    if S.device_count() != 0:
        pytest.skip("covered by GPU tests")
    for blob in (b"\x00\x01\x02", b"\xff" * 64):
        with pytest.raises(S.Sbv2Error) as e:
            S.Model(blob, bert=False)
        assert e.value.status in (S.ERR_PARSE, S.ERR_CUDA)
    with pytest.raises(S.Sbv2Error):
        S.Model(b"", bert=False)


# ---- structural binding of anonymous weights (onnx_bind.h) --------------------------------------------

def test_bind_report_deberta_anonymous_linears(S):
    """A TorchScript + onnxsim export (convert_deberta.py:36-52) keeps only biases / embeddings / LayerNorms named: every
    encoder Linear is a transposed onnx::MatMul_N initializer.  The binder must recover all of them through MatMul -> Add."""
    import util  # noqa: F401
    from oracle import deberta as od
    from sbv2_b200 import assets
    cfg = od.tiny_config()
    sd = od.state_dict_numpy(od.build_model(cfg, seed=1))
    named = S.onnx_bind_report(assets.deberta_onnx(sd), True)
    assert named["bound"] == {} and named["unbound_biases"] == []
    anon_bytes = assets.deberta_onnx(sd, anonymize_linear=True)
    assert b"query_proj.weight" not in anon_bytes and b"onnx::MatMul_" in anon_bytes
    anon = S.onnx_bind_report(anon_bytes, True)
    assert len(anon["bound"]) == 6 * cfg.num_hidden_layers and anon["unbound_biases"] == []
    for name, how in anon["bound"].items():
        assert how["transposed"] is True and how["via"] == "MatMul+Add" and how["initializer"].startswith("onnx::MatMul_")
    # a bias whose weight cannot be bound is reported (the loaders turn it into a hard error)
    broken = dict(sd)
    del broken["deberta.encoder.layer.1.intermediate.dense.weight"]
    rep = S.onnx_bind_report(assets.deberta_onnx(broken), True)
    assert rep["unbound_biases"] == ["deberta.encoder.layer.1.intermediate.dense.bias"]


def test_bind_report_synth_anonymous_convs_and_linears(S):
    import util
    from util import ov
    from sbv2_b200 import assets
    for flow in (True, False):
        hp = ov.tiny_hparams(use_transformer_flow=flow)
        sd = ov.state_dict_numpy(ov.build_model(hp, seed=0))
        rep = S.onnx_bind_report(assets.synth_onnx(sd, hp.upsample_rates, hp.resblock_dilation_sizes, anonymize_weight_norm=True,
                                                   anonymize_linear=True), False)
        b = rep["bound"]
        assert rep["unbound_biases"] == []
        assert b["dec.ups.0.weight"]["via"] == "ConvTranspose" and b["dec.resblocks.0.convs1.0.weight"]["via"] == "Conv"
        spk = [k for k in b if k.endswith("spk_emb_linear.weight")]
        assert "enc_p.encoder.spk_emb_linear.weight" in spk and all(b[k]["transposed"] for k in spk)
        if flow:
            assert len(spk) == 1 + hp.n_flow_layer
        else:
            assert any(".enc.in_layers.0.weight" in k for k in b) and any(".enc.cond_layer.weight" in k for k in b)


def test_container_parsers_reject_hostile_input(S):
    """tar size wrap-around, JSON numbers at the end of an unterminated buffer, non-JSON number forms,
    malformed .npy headers."""
    from sbv2_b200 import assets
    sv = np.zeros((2, 256), np.float32)
    onnx = assets.model_proto({"w": np.zeros(4, np.float32)})
    # GNU base-256 size field close to 2^64 in the first tar header
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w") as w:
        ti = tarfile.TarInfo("model.onnx")
        ti.size = len(onnx)
        w.addfile(ti, io.BytesIO(onnx))
    tar = bytearray(buf.getvalue())
    tar[124:136] = b"\x80" + b"\xff" * 11
    with pytest.raises(S.Sbv2Error) as e:
        S.parse_sbv2file(assets.zstd_compress(bytes(tar)))
    assert e.value.status == S.ERR_PARSE and "exceeds archive" in e.value.message
    # numbers: only the JSON grammar; the buffer is not NUL-terminated (bytes object sliced mid-number is still bounded)
    for bad in (b'{"shape":[1,2],"data":[[nan,1]]}', b'{"shape":[1,2],"data":[[0x10,1]]}', b'{"shape":[1,2],"data":[[inf,1]]}',
                b'{"shape":[1,2],"data":[[1.,1]]}', b'{"shape":[1,2],"data":[[1e,1]]}', b'{"shape":[1,2],"data":[[' + b"1" * 80 + b',1]]}'):
        with pytest.raises(S.Sbv2Error) as e:
            S.load_style(bad)
        assert e.value.status == S.ERR_PARSE, bad
    np.testing.assert_array_equal(S.load_style(b'{"shape":[1,3],"data":[[-1.5e-1, 2E+0, 3]]}'), np.array([[-0.15, 2.0, 3.0]], np.float32))
    with pytest.raises(S.Sbv2Error):
        S.load_style(b'{"shape":[1,2],"data":[[1,2')       # truncated right after a number
    with pytest.raises(S.Sbv2Error):
        S.load_style(b'{"shape":[4611686018427387904,4],"data":[[1,2]]}')  # rows*cols would wrap
    # .npy headers without the expected quotes / parentheses, and shapes whose product overflows
    def npy(header: bytes, payload: bytes = b"") -> bytes:
        h = header + b" " * ((64 - (10 + len(header) + 1) % 64) % 64) + b"\n"
        return base64.b64encode(b"\x93NUMPY\x01\x00" + struct.pack("<H", len(h)) + h + payload)
    for hdr in (b"{'descr': <f4, 'fortran_order': False, 'shape': (1, 2), }", b"{'descr': '<f4', 'fortran_order': False, 'shape': 1, 2, }",
                b"{'descr': '<f4', 'fortran_order':", b"{'descr' '<f4', 'fortran_order': False, 'shape': (1, 2), }",
                b"{'descr': '<f4', 'fortran_order': False, 'shape': (4611686018427387904, 8), }",
                b"{'descr': '<f4', 'fortran_order': False, 'shape': (-1, 2), }"):
        with pytest.raises(S.Sbv2Error):
            S.load_style_npy_base64(npy(hdr, b"\0" * 8))
    ok = S.load_style_npy_base64(npy(b"{'descr': '<f4', 'fortran_order': False, 'shape': (1, 2), }", np.array([1, 2], "<f4").tobytes()))
    np.testing.assert_array_equal(ok, [[1.0, 2.0]])


def test_zstd_streaming_frame_without_content_size(S):
    """Frames written by a streaming compressor carry no content size; a frame whose decompressed size is an exact
    multiple of the 1 MiB output chunk used to be reported as truncated."""
    import ctypes
    from sbv2_b200 import assets
    z = ctypes.CDLL("libzstd.so.1")
    z.ZSTD_createCStream.restype = ctypes.c_void_p
    z.ZSTD_initCStream.argtypes = [ctypes.c_void_p, ctypes.c_int]
    z.ZSTD_compressStream.restype = ctypes.c_size_t
    z.ZSTD_endStream.restype = ctypes.c_size_t

    class Buf(ctypes.Structure):
        _fields_ = [("p", ctypes.c_void_p), ("size", ctypes.c_size_t), ("pos", ctypes.c_size_t)]
    z.ZSTD_compressStream.argtypes = [ctypes.c_void_p, ctypes.POINTER(Buf), ctypes.POINTER(Buf)]
    z.ZSTD_endStream.argtypes = [ctypes.c_void_p, ctypes.POINTER(Buf)]

    sv = np.zeros((1, 256), np.float32)
    onnx = assets.model_proto({"w": np.arange(300000, dtype=np.float32)})
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w", format=tarfile.USTAR_FORMAT) as w:
        for name, b in (("model.onnx", onnx), ("style_vectors.json", assets.style_json(sv))):
            ti = tarfile.TarInfo(name)
            ti.size = len(b)
            w.addfile(ti, io.BytesIO(b))
    tar = buf.getvalue()
    tar += b"\0" * ((-len(tar)) % (1 << 20))          # decompressed size = k * 1 MiB exactly
    cs = z.ZSTD_createCStream()
    z.ZSTD_initCStream(cs, 3)
    dst = ctypes.create_string_buffer(len(tar) + 4096)
    src = ctypes.create_string_buffer(tar, len(tar))
    o = Buf(ctypes.cast(dst, ctypes.c_void_p), len(dst), 0)
    i = Buf(ctypes.cast(src, ctypes.c_void_p), len(tar), 0)
    while i.pos < i.size:
        z.ZSTD_compressStream(cs, ctypes.byref(o), ctypes.byref(i))
    while z.ZSTD_endStream(cs, ctypes.byref(o)) != 0:
        pass
    frame = dst.raw[:o.pos]
    style_json, model_onnx = S.parse_sbv2file(frame)
    assert model_onnx == onnx
    with pytest.raises(S.Sbv2Error):
        S.parse_sbv2file(frame[:-5])                  # a really truncated frame still fails
