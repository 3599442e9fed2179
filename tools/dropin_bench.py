"""Developer tool: the reference-signature builders and the batched find_linear_fit on a
realistic scene (2048 x 2048 frames, 2000 segmented sources, affine WCS pair): wall time of
create_primary_cutouts, create_cutouts and align.find_linear_fit (GPU blot), per source.
Writes gpurun_out/dropin_bench.json."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from wcs_stubs import AffineWCS
from subpixal_b200 import cutout, align


class LinWCS(AffineWCS):
    """AffineWCS + what find_linear_fit asks of a tangent-plane WCS."""
    sip = cpdis1 = cpdis2 = det2im1 = det2im2 = None

    class wcs:
        crpix = [1024.0, 1024.0]

    def deepcopy(self):
        import copy
        return copy.deepcopy(self)

    def wcs_world2pix(self, *a, **k):
        return self.all_world2pix(*a, **k)


def main(nsrc=2000, n=2048, seed=1):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n]
    drz = 0.05 * rng.standard_normal((n, n)).astype(np.float32)
    seg = np.zeros((n, n), dtype=np.int32)
    ids, xs, ys = [], [], []
    for k in range(nsrc):
        r = int(rng.integers(6, 16))
        cx, cy = rng.uniform(40, n - 40, 2)
        x1, y1 = int(cx) - r, int(cy) - r
        sl = np.s_[y1:y1 + 2 * r, x1:x1 + 2 * r]
        if seg[sl].any():
            continue
        g = np.exp(-((xx[sl] - cx) ** 2 + (yy[sl] - cy) ** 2) / (2 * (r / 3.0) ** 2))
        drz[sl] += (200 * g).astype(np.float32)
        seg[sl][g > 0.02] = k + 1
        ids.append(k + 1); xs.append(cx); ys.append(cy)
    cat = {"id": np.array(ids), "x": np.array(xs), "y": np.array(ys), "weight": np.ones(len(ids))}
    flt = np.roll(drz, (1, -1), axis=(0, 1)) + 0.05 * rng.standard_normal((n, n)).astype(np.float32)
    drz_wcs = LinWCS(np.eye(2), [0.0, 0.0])
    flt_wcs = LinWCS(np.eye(2), [-1.3, 0.8])
    out = {"sources": len(ids), "frame": [n, n]}
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        prim = cutout.create_primary_cutouts(cat, seg, drz, drz_wcs, pad=2)
        t1 = time.perf_counter()
        flt_cts, drz_cts = cutout.create_cutouts(prim, seg, drz, drz_wcs, flt, flt_wcs, pad=3)
        t2 = time.perf_counter()
        fit, iccs, blts = align.find_linear_fit(flt_cts, drz_cts, fitgeom="rscale")
        torch.cuda.synchronize(); t3 = time.perf_counter()
    out.update({"cutouts": len(flt_cts), "create_primary_cutouts_ms": (t1 - t0) * 1e3,
                "create_cutouts_ms": (t2 - t1) * 1e3, "find_linear_fit_ms": (t3 - t2) * 1e3,
                "fit_offset": [float(v) for v in fit["offset"]], "fit_rmse": float(fit["rmse"]),
                "us_per_source_total": (t3 - t0) * 1e6 / max(1, len(flt_cts))})
    print(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dropin_bench.json"), "w"), indent=1)


main()
