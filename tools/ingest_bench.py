"""Throughput of the FITS decode kernel (HBM-bound byte work) and of the whole file ->
device ingest.  Writes gpurun_out/ingest_bench.json.  Developer / profiling tool."""
import ctypes as C
import json, os, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from subpixal_b200 import _lib as L, batch, ingest
from fits_writer import write_fits


def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    peak = 6540.8
    try: peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception: pass
    lib = L.load()
    out = []
    n = 8192 * 8192                                       # 268 MB of float32: >> L2
    raw = torch.randint(0, 255, (n * 8,), dtype=torch.uint8, device="cuda")
    dst = torch.empty(n * 8, dtype=torch.uint8, device="cuda")
    msk = torch.empty(n * 8, dtype=torch.uint8, device="cuda")
    cases = [("f32 -> f32 in place", -32, L.T_F32, 4, 4, True, 0, 0, 1.0, 0.0, 0.0),
             ("f32 -> f32 + sky", -32, L.T_F32, 4, 4, False, 0, 0, 1.0, 0.0, 0.37),
             ("f32 -> f32 + zero mask u8 (wht -> dq)", -32, L.T_F32, 4, 4, True, L.T_I8, 1, 1.0, 0.0, 0.0),
             ("f32 -> f32 + zero mask i64", -32, L.T_F32, 4, 4, True, L.T_I64, 8, 1.0, 0.0, 0.0),
             ("i16 -> i16 in place (DQ)", 16, L.T_I16, 2, 2, True, 0, 0, 1.0, 0.0, 0.0),
             ("i16 -> f32 scaled", 16, L.T_F32, 2, 4, False, 0, 0, 0.01, 100.5, 0.0),
             ("i32 -> i32 in place (segmentation)", 32, L.T_I32, 4, 4, True, 0, 0, 1.0, 0.0, 0.0),
             ("f64 -> f64 in place", -64, L.T_F64, 8, 8, True, 0, 0, 1.0, 0.0, 0.0)]
    for name, bp, kind, sin, sout, inplace, mkind, msize, bs, bz, sky in cases:
        o = raw if inplace else dst
        fn = lambda: L.check(lib.spx_fits_decode(batch._ptr(raw), n, bp, bs, bz, sky, kind, batch._ptr(o),
                                                 mkind, batch._ptr(msk) if msize else C.c_void_p(0),
                                                 batch._stream()))
        ms = timeit(fn)
        by = n * (sin + sout + msize)
        out.append({"kernel": "fits_decode " + name, "elements": n, "ms": ms, "algorithmic_GB": by / 1e9,
                    "GBps": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak})
        print(out[-1], flush=True)
    # whole ingest: file (page cache) -> pinned staging -> device -> decode
    a = np.random.default_rng(0).standard_normal((4096, 4096)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "x.fits")
        write_fits(path, [dict(data=None), dict(data=a, extname="SCI", extver=1)])
        ingest.read_image(path + "[sci,1]"); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            t = ingest.read_image(path + "[sci,1]")
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        assert np.array_equal(t.cpu().numpy(), a)
        t0 = time.perf_counter()
        for _ in range(5):
            with open(path, "rb") as f:
                f.seek(5760)
                host = np.frombuffer(f.read(a.nbytes), dtype=">f4").astype(np.float32)
            t2 = torch.from_numpy(host).cuda()
        torch.cuda.synchronize()
        dt_host = (time.perf_counter() - t0) / 5
        out.append({"path": "FITS file (page cache) -> device float32 [4096,4096]", "ms": dt * 1e3,
                    "GBps_file": a.nbytes / dt / 1e9,
                    "ms_numpy_decode_then_upload": dt_host * 1e3})
        print(out[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ingest_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
