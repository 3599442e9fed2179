"""Throughput of the batched blot kernel (4 half-pixel phases per source).
Writes gpurun_out/blot_bench.json.  Developer / profiling tool."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from subpixal_b200 import batch

def fit_map(f, nx, ny):
    xc, yc = 0.5 * (nx - 1), 0.5 * (ny - 1)
    uu, vv = np.meshgrid(np.array([-xc, 0, xc]), np.array([-yc, 0, yc]), indexing="ij")
    xs, ys = f(uu + xc, vv + yc)
    def diff(g, h, ax):
        fm, f0, fp = np.moveaxis(g, ax, 0)
        return np.moveaxis(np.stack([f0, (fp - fm) / (2 * h), ((fp - f0) - (f0 - fm)) / (2 * h * h)]), 0, ax)
    return np.concatenate([[xc, yc], diff(diff(xs, xc, 0), yc, 1).ravel(), diff(diff(ys, xc, 0), yc, 1).ravel()])

def main():
    peak = 6540.8
    try: peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception: pass
    out = []
    g = torch.Generator(device="cuda").manual_seed(0)
    cases = ((64, 10000), (128, 5000), (32, 50000))
    variants = (("poly5", True), ("poly5", False), ("linear", False))
    if len(sys.argv) > 1:            # e.g.  blot_bench.py 64 poly5 1   (for ncu)
        cases = tuple(c for c in cases if c[0] == int(sys.argv[1]))
        variants = ((sys.argv[2], sys.argv[3] == "1"),)
    for n_side, n in cases:
        hs = n_side + 16                     # source cutouts a little larger than the targets
        src = torch.randn((n * hs * hs,), device="cuda", generator=g)
        rng = np.random.default_rng(1)
        maps = np.stack([fit_map(lambda x, y, t=rng.uniform(7, 9, 2), a=rng.uniform(-2e-3, 2e-3):
                                 (x + a * y + t[0], y - a * x + t[1]), n_side, n_side) for _ in range(n)])
        s_shp = np.full((n, 2), hs); t_shp = np.full((n, 2), n_side)
        s_off = np.arange(n) * hs * hs
        for interp, prep in variants:
            fn = lambda: batch.blot_phases(src, s_off, s_shp, maps, t_shp, interp=interp, prepare=prep)
            for _ in range(2): fn()
            torch.cuda.synchronize()
            # device time of the kernel alone: the host-side table upload is outside the events
            import ctypes as C
            from subpixal_b200 import _lib as L
            lib = L.load()
            dev = src.device
            keep = [torch.as_tensor(s_off).to(dev), torch.as_tensor(s_shp, dtype=torch.int32).to(dev),
                    torch.as_tensor(maps).to(dev), torch.as_tensor(t_shp, dtype=torch.int32).to(dev),
                    (torch.arange(n, dtype=torch.int64) * 4 * n_side * n_side).to(dev)]
            o = torch.empty(n * 4 * n_side * n_side, device=dev)
            a = L.BlotArgs()
            a.src = src.data_ptr(); a.src_offset = keep[0].data_ptr(); a.src_shape = keep[1].data_ptr()
            a.map = keep[2].data_ptr(); a.tgt_shape = keep[3].data_ptr(); a.out_offset = keep[4].data_ptr()
            a.out = o.data_ptr()
            ph = (C.c_double * 8)(0, 0, .5, 0, 0, .5, .5, .5)
            a.phases_host = C.cast(ph, C.POINTER(C.c_double))
            a.n, a.nph, a.interp, a.max_ny, a.max_nx = n, 4, {"poly5": 2, "linear": 1}[interp], n_side, n_side
            if prep:
                ws = torch.empty(n * (hs + 5) * (hs + 1) * 4, device=dev)
                wo = (torch.arange(n, dtype=torch.int64) * (hs + 5) * (hs + 1)).to(dev)
                a.workspace, a.ws_offset, a.max_src_h, a.max_src_w = ws.data_ptr(), wo.data_ptr(), hs, hs
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            for _ in range(3): L.check(lib.spx_blot_phases(C.byref(a), st))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): L.check(lib.spx_blot_phases(C.byref(a), st))
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            by = n * (hs * hs * 4 + 4 * n_side * n_side * 4)
            out.append({"kernel": "blot_kernel<%s%s> x4 phases" % (interp, ", prepared" if prep else ""), "n": n, "target": n_side, "source": hs,
                        "ms": ms, "sources_per_s": n / ms * 1e3, "Mpix_per_s": n * 4 * n_side * n_side / ms / 1e3,
                        "algorithmic_GB": by / 1e9, "GBps": by / ms / 1e6, "hbm_frac": by / ms / 1e6 / peak})
            print(out[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/blot_bench.json", "w"), indent=1)

main()
