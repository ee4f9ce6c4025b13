"""tcgen05.mma issue rate with epilogue-like background traffic from the other warps (see umma_microbench.cu)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sbv2-api_b200"))
import sbv2_b200 as S
fn = S.debug_lib().sbv2_debug_mma_rate2
fn.restype = C.c_int
fn.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_longlong)]
out = (C.c_longlong * 2)()
it = 4400
if len(sys.argv) > 1 and sys.argv[1] == "conv":
    for n in (16, 32, 64):
        for ra in (1040,):
            for noise in (8, 32, 32 + 7):
                st = fn(n, ra, it, 8, 148, noise, out)
                if st: print("ERR", S.debug_lib().sbv2_last_error().decode()); continue
                print(f"conv-like N={n:3d} RA={ra:4d} noise={noise}: issue {out[0]/it:6.1f} complete {out[1]/it:6.1f} cyc/MMA")
    sys.exit(0)
for blocks in (1, 148):
    for n in (16, 64, 128):
        for ra in (306, 1074):
            for nacc in (1, 4):
                if nacc * n > 256: continue
                for noise in (0, 1, 2, 4, 7):
                    st = fn(n, ra, it, nacc, blocks, noise, out)
                    if st: print("ERR", S.debug_lib().sbv2_last_error().decode()); continue
                    print(f"blocks={blocks:3d} N={n:3d} RA={ra:4d} nacc={nacc} noise={noise}: issue {out[0]/it:6.1f} complete {out[1]/it:6.1f} cyc/MMA")
