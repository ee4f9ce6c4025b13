"""Probe of cp.async.bulk ... .multicast::cluster on this GPU (debug library): what lands where."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sbv2-api_b200"))
import sbv2_b200 as S  # noqa: E402

fn = S.debug_lib().sbv2_debug_multicast_probe
fn.restype = C.c_int
pu = C.POINTER(C.c_uint32)
fn.argtypes = [C.c_int, C.c_int, C.c_int, pu, pu]
for nc in (2, 4):
    words, blocks = 1024, 2 * nc
    out = np.zeros(blocks * words, np.uint32)
    info = np.zeros(blocks * 4, np.uint32)
    st = fn(nc, words, blocks, out.ctypes.data_as(pu), info.ctypes.data_as(pu))
    if st:
        print("ERROR", S.debug_lib().sbv2_last_error().decode())
        continue
    out = out.reshape(blocks, words)
    for b in range(blocks):
        ok = np.array_equal(out[b], np.arange(words, dtype=np.uint32))
        parts = [bool(np.array_equal(out[b, r * words // nc:(r + 1) * words // nc], np.arange(r * words // nc, (r + 1) * words // nc, dtype=np.uint32)))
                 for r in range(nc)]
        print(f"nc={nc} block {b}: rank {info[b * 4]} smem 0x{info[b * 4 + 1]:08x} bar 0x{info[b * 4 + 2]:08x} all-ok {ok} slices-ok {parts} "
              f"first words {out[b, :2]} {out[b, words // nc:words // nc + 2]}")
