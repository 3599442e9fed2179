"""Developer tool: run the FITS decode kernel a few times (for ncu): f32 in place with a uint8
zero plane (the weight -> DQ read of align.py:308-312) and int16 -> float32 scaled."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from subpixal_b200 import _lib as L, batch
lib = L.load()
n = 8192 * 8192
raw = torch.randint(0, 255, (n * 4,), dtype=torch.uint8, device="cuda")
out = torch.empty(n * 4, dtype=torch.uint8, device="cuda")
msk = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    L.check(lib.spx_fits_decode(batch._ptr(raw), n, -32, 1.0, 0.0, 0.0, L.T_F32, batch._ptr(raw), L.T_I8,
                                batch._ptr(msk), batch._stream()))
    L.check(lib.spx_fits_decode(batch._ptr(raw), n, 16, 0.01, 100.5, 0.0, L.T_F32, batch._ptr(out), 0,
                                C.c_void_p(0), batch._stream()))
torch.cuda.synchronize()
