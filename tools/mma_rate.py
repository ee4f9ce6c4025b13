import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sbv2-api_b200"))
import sbv2_b200 as S
fn = S.debug_lib().sbv2_debug_mma_rate
fn.restype = C.c_int
fn.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_longlong)]
out = (C.c_longlong * 2)()
for blocks in (1,):
    for layout in (0, 1):
        for n in (16, 32, 64, 128, 256):
            for shift in (0, 3):
                for nacc in (1, 2):
                    if nacc * n > 512: continue
                    it = 4000
                    st = fn(n, layout, shift, it, nacc, blocks, out)
                    if st: print("ERR", S.debug_lib().sbv2_last_error().decode()); continue
                    print(f"blocks={blocks:3d} layout={'none ' if layout==0 else 'sw128'} N={n:3d} shift={shift} nacc={nacc}: issue {out[0]/it:6.1f} cyc/MMA, complete {out[1]/it:6.1f} cyc/MMA (ideal {max(n,8)/2:.0f})")
