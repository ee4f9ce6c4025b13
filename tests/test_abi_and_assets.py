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


def test_onnx_reader_rejects_garbage(S):
    if S.device_count() != 0:
        pytest.skip("covered by GPU tests")
    for blob in (b"\x00\x01\x02", b"\xff" * 64):
        with pytest.raises(S.Sbv2Error) as e:
            S.Model(blob, bert=False)
        assert e.value.status in (S.ERR_PARSE, S.ERR_CUDA)
    with pytest.raises(S.Sbv2Error):
        S.Model(b"", bert=False)
