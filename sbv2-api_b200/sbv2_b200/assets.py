"""Writers for the asset formats the backend loads: ONNX ModelProto (hand-encoded protobuf — the
``onnx`` package is not available offline), ``.sbv2`` (zstd(tar)), ``style_vectors.json`` and the
``.aivmx`` metadata entry.  Used to mint synthetic models for tests and benchmarks; the layouts
follow /root/reference/scripts/convert/convert_model.py:39-46,115-174 and
crates/sbv2_core/src/tts.rs:93-108.
"""
from __future__ import annotations

import base64
import ctypes
import io
import json
import struct
import tarfile
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

# ---- protobuf wire helpers -------------------------------------------------------------------


def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wt: int) -> bytes:
    return _varint((field << 3) | wt)


def _ld(field: int, payload: bytes) -> bytes:
    return _key(field, 2) + _varint(len(payload)) + payload


def _vi(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(v)


def _s(field: int, s: str) -> bytes:
    return _ld(field, s.encode("utf-8"))


_DTYPES = {np.dtype("float32"): 1, np.dtype("int64"): 7, np.dtype("float16"): 10, np.dtype("int32"): 6,
           np.dtype("float64"): 11}


def tensor_proto(name: str, arr: np.ndarray, raw: bool = True) -> bytes:
    arr = np.ascontiguousarray(arr)
    out = b"".join(_vi(1, int(d)) for d in arr.shape)
    out += _vi(2, _DTYPES[arr.dtype])
    out += _s(8, name)
    if raw:
        out += _ld(9, arr.tobytes())
    else:
        assert arr.dtype == np.float32
        out += _ld(4, arr.tobytes())  # packed float_data
    return out


def attribute_ints(name: str, vals: Sequence[int]) -> bytes:
    out = _s(1, name)
    out += _ld(8, b"".join(_varint(int(v)) for v in vals))
    out += _vi(20, 7)  # AttributeType.INTS
    return out


def attribute_int(name: str, v: int) -> bytes:
    return _s(1, name) + _vi(3, int(v)) + _vi(20, 2)


def node_proto(op_type: str, inputs: Sequence[str], outputs: Sequence[str], name: str = "",
               attrs: Iterable[bytes] = ()) -> bytes:
    out = b"".join(_s(1, i) for i in inputs)
    out += b"".join(_s(2, o) for o in outputs)
    if name:
        out += _s(3, name)
    out += _s(4, op_type)
    out += b"".join(_ld(5, a) for a in attrs)
    return out


def value_info(name: str) -> bytes:
    return _s(1, name)


def model_proto(initializers: Dict[str, np.ndarray], nodes: Sequence[bytes] = (), inputs: Sequence[str] = (),
                outputs: Sequence[str] = (), metadata: Optional[Dict[str, str]] = None,
                producer: str = "sbv2_b200.assets", packed_float_names: Sequence[str] = ()) -> bytes:
    parts = [_ld(1, n) for n in nodes]
    parts.append(_s(2, "main_graph"))
    for name, arr in initializers.items():
        parts.append(_ld(5, tensor_proto(name, arr, raw=name not in packed_float_names)))
    parts += [_ld(11, value_info(n)) for n in inputs]
    parts += [_ld(12, value_info(n)) for n in outputs]
    g = b"".join(parts)
    mparts = [_vi(1, 8), _s(2, producer), _ld(7, g), _ld(8, _s(1, "") + _vi(2, 17))]  # ir_version, producer, graph, opset
    for k, v in (metadata or {}).items():
        mparts.append(_ld(14, _s(1, k) + _s(2, v)))
    m = b"".join(mparts)
    return m


# ---- synthesizer / deberta graphs --------------------------------------------------------------

SYNTH_INPUTS = ["x_tst", "x_tst_lengths", "sid", "tones", "language", "bert", "style_vec", "length_scale",
                "sdp_ratio", "noise_scale", "noise_scale_w"]  # convert_model.py:141-153


def synth_onnx(state: Dict[str, np.ndarray], upsample_rates: Sequence[int], resblock_dilations: Sequence[Sequence[int]],
               anonymize_weight_norm: bool = False, metadata: Optional[Dict[str, str]] = None,
               anonymize_linear: bool = False) -> bytes:
    """ModelProto for a JP-Extra synthesizer whose initializers are ``state`` (upstream state_dict
    names, weight-norm folded).  The decoder's Conv / ConvTranspose nodes are emitted in execution
    order with their strides / dilations, as a traced export carries them; with
    ``anonymize_weight_norm`` the weight-normed convs (decoder ups / resblocks, WN flow layers) get
    ``onnx::Conv_N`` names like a real export after constant folding, and with ``anonymize_linear``
    every nn.Linear the graph applies to a 3-D input (``spk_emb_linear`` of the encoders) becomes a
    TRANSPOSED ``onnx::MatMul_N`` initializer feeding MatMul -> Add with the still-named bias
    (scripts/convert/convert_model.py:115-156)."""
    inits: Dict[str, np.ndarray] = {}
    rename: Dict[str, str] = {}
    transposed: set = set()
    counter = [1000]

    def reg(name: str) -> str:
        if anonymize_weight_norm and (name.startswith("dec.ups") or name.startswith("dec.resblocks")) \
                and name.endswith(".weight"):
            op = "ConvTranspose" if name.startswith("dec.ups") else "Conv"
            counter[0] += 1
            rename[name] = f"onnx::{op}_{counter[0]}"
            return rename[name]
        return name

    nodes: List[bytes] = []
    if anonymize_weight_norm:
        # WN flow variant: in_layers / res_skip_layers / cond_layer are weight-normed Conv1d
        for k_ in sorted(state):
            if k_.startswith("flow.flows.") and ".enc." in k_ and k_.endswith(".weight") and \
                    any(t in k_ for t in (".in_layers.", ".res_skip_layers.", ".cond_layer.")):
                counter[0] += 1
                rename[k_] = f"onnx::Conv_{counter[0]}"
                kk = state[k_].shape[2]
                nodes.append(node_proto("Conv", ["wn_in", rename[k_], k_[:-7] + ".bias"], [k_ + "_out"], "/" + k_[:-7].replace(".", "/") + "/Conv",
                                        [attribute_ints("kernel_shape", [kk]), attribute_ints("strides", [1]), attribute_int("group", 1)]))
    if anonymize_linear:
        for k_ in sorted(state):
            if k_.endswith(".spk_emb_linear.weight"):
                counter[0] += 1
                rename[k_] = f"onnx::MatMul_{counter[0]}"
                transposed.add(k_)
                base = k_[:-7]
                nodes.append(node_proto("MatMul", ["g_t", rename[k_]], [base + "_mm"], "/" + base.replace(".", "/") + "/MatMul"))
                nodes.append(node_proto("Add", [base + ".bias", base + "_mm"], [base + "_out"], "/" + base.replace(".", "/") + "/Add"))
        if "enc_p.style_proj.weight" in state:  # Linear on a 2-D input exports as Gemm(transB=1) and keeps its name
            nodes.append(node_proto("Gemm", ["style_vec", "enc_p.style_proj.weight", "enc_p.style_proj.bias"], ["style_emb"],
                                    "/enc_p/style_proj/Gemm", [attribute_int("transB", 1)]))
    n_ups = sum(1 for k in state if k.startswith("dec.ups.") and k.endswith(".weight"))
    n_res_per = len(resblock_dilations)
    cur = "dec_in"
    nodes.append(node_proto("Conv", [cur, "dec.conv_pre.weight", "dec.conv_pre.bias"], ["dec_pre"], "/dec/conv_pre/Conv",
                            [attribute_ints("dilations", [1]), attribute_ints("kernel_shape", [7]),
                             attribute_ints("pads", [3, 3]), attribute_ints("strides", [1]), attribute_int("group", 1)]))
    cur = "dec_pre"
    for i in range(n_ups):
        w = state[f"dec.ups.{i}.weight"]
        k, u = w.shape[2], int(upsample_rates[i])
        p = (k - u) // 2
        nodes.append(node_proto("ConvTranspose", [cur, reg(f"dec.ups.{i}.weight"), f"dec.ups.{i}.bias"], [f"up{i}"],
                                f"/dec/ups.{i}/ConvTranspose",
                                [attribute_ints("dilations", [1]), attribute_ints("kernel_shape", [k]),
                                 attribute_ints("pads", [p, p]), attribute_ints("strides", [u]), attribute_int("group", 1)]))
        cur = f"up{i}"
        for j in range(n_res_per):
            rb = i * n_res_per + j
            x = cur
            for l, d in enumerate(resblock_dilations[j]):
                for which, dd in (("convs1", d), ("convs2", 1)):
                    nm = f"dec.resblocks.{rb}.{which}.{l}"
                    kk = state[nm + ".weight"].shape[2]
                    pad = (kk * dd - dd) // 2
                    nodes.append(node_proto("Conv", [x, reg(nm + ".weight"), nm + ".bias"], [nm + "_out"],
                                            f"/dec/resblocks.{rb}/{which}.{l}/Conv",
                                            [attribute_ints("dilations", [dd]), attribute_ints("kernel_shape", [kk]),
                                             attribute_ints("pads", [pad, pad]), attribute_ints("strides", [1]),
                                             attribute_int("group", 1)]))
                    x = nm + "_out"
    for k_, v in state.items():
        v = np.asarray(v)
        inits[rename.get(k_, k_)] = np.ascontiguousarray(v.T) if k_ in transposed else v
    return model_proto(inits, nodes, SYNTH_INPUTS, ["output"], metadata)


def deberta_onnx(state: Dict[str, np.ndarray], metadata: Optional[Dict[str, str]] = None,
                 anonymize_linear: bool = False) -> bytes:
    """ModelProto for the DeBERTa-v2 feature encoder; initializers keep HF ``state_dict`` names
    (``deberta.embeddings.word_embeddings.weight`` ...). scripts/convert/convert_deberta.py:47-48.

    ``anonymize_linear`` reproduces what the TorchScript export + onnxsim of convert_deberta.py:36-52 does to the
    graph: every encoder Linear (query/key/value_proj, attention.output.dense, intermediate.dense, output.dense) is
    constant-folded into a TRANSPOSED [in, out] ``onnx::MatMul_N`` initializer consumed by MatMul -> Add(bias), only
    the biases / embeddings / LayerNorm parameters / the ConvLayer weight keep their names."""
    inits: Dict[str, np.ndarray] = {}
    nodes: List[bytes] = []
    counter = 2000
    linear = (".attention.self.query_proj", ".attention.self.key_proj", ".attention.self.value_proj",
              ".attention.output.dense", ".intermediate.dense", ".output.dense")
    for k, v in state.items():
        v = np.asarray(v)
        if anonymize_linear and k.endswith(".weight") and ".encoder.layer." in k and k[:-7].endswith(linear) and v.ndim == 2:
            counter += 1
            anon = f"onnx::MatMul_{counter}"
            base = k[:-7]
            inits[anon] = np.ascontiguousarray(v.T)
            nodes.append(node_proto("MatMul", [base + "_in", anon], [base + "_mm"], "/" + base.replace(".", "/") + "/MatMul"))
            # TorchScript emits Add(bias, matmul) for Linear on 3-D inputs
            nodes.append(node_proto("Add", [base + ".bias", base + "_mm"], [base + "_out"], "/" + base.replace(".", "/") + "/Add"))
        else:
            inits[k] = v
    return model_proto(inits, nodes, ["input_ids", "token_type_ids", "attention_mask"], ["output"], metadata)


# ---- containers --------------------------------------------------------------------------------

_zstd = None


def _libzstd():
    global _zstd
    if _zstd is None:
        lib = ctypes.CDLL("libzstd.so.1")
        lib.ZSTD_compressBound.restype = ctypes.c_size_t
        lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        lib.ZSTD_compress.restype = ctypes.c_size_t
        lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.ZSTD_isError.restype = ctypes.c_uint
        lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
        _zstd = lib
    return _zstd


def zstd_compress(data: bytes, level: int = 3) -> bytes:
    lib = _libzstd()
    bound = lib.ZSTD_compressBound(len(data))
    dst = ctypes.create_string_buffer(bound)
    n = lib.ZSTD_compress(dst, bound, data, len(data), level)
    if lib.ZSTD_isError(n):
        raise RuntimeError("zstd compress failed")
    return dst.raw[:n]


def style_json(style_vectors: np.ndarray) -> bytes:
    arr = np.asarray(style_vectors, dtype=np.float32)
    return json.dumps({"data": arr.tolist(), "shape": list(arr.shape)}).encode("utf-8")


def sbv2_file(onnx_bytes: bytes, style_vectors: np.ndarray, level: int = 3, extra: Optional[Dict[str, bytes]] = None,
              omit: Sequence[str] = ()) -> bytes:
    """zstd(tar{version.txt, model.onnx, style_vectors.json}) — convert_model.py:159-174."""
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w") as w:
        def add(name: str, b: bytes):
            ti = tarfile.TarInfo(name)
            ti.size = len(b)
            w.addfile(ti, io.BytesIO(b))
        entries = {"version.txt": b"1", "model.onnx": onnx_bytes, "style_vectors.json": style_json(style_vectors)}
        entries.update(extra or {})
        for k, v in entries.items():
            if k not in omit:
                add(k, v)
    return zstd_compress(buf.getvalue(), level)


def aivmx_metadata(style_vectors: np.ndarray, fortran: bool = False) -> Dict[str, str]:
    arr = np.asarray(style_vectors, dtype=np.float32)
    if fortran:
        arr = np.asfortranarray(arr)
    b = io.BytesIO()
    np.save(b, arr)
    return {"aivm_style_vectors": base64.b64encode(b.getvalue()).decode("ascii")}
