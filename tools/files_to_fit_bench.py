"""Developer tool: the whole chain from FILES to a linear fit on the device-level API -
FITS ingest (drizzled image + weight image, input image minus sky, segmentation image,
SExtractor catalogue), segment bounding boxes, the per-image measurement
(align.frame_displacements: 2 gathers, blot x 4, CC + fit) and the linear fit.
Writes gpurun_out/files_to_fit.json."""
import json, os, sys, tempfile, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fits_writer import write_fits
from subpixal_b200 import align, batch, ingest


def scene(n=4096, nsrc=12000, seed=3):
    rng = np.random.default_rng(seed)
    drz = (0.05 * rng.standard_normal((n, n))).astype(np.float32)
    seg = np.zeros((n, n), dtype=np.int32)
    rows = []
    yy, xx = np.mgrid[0:40, 0:40]
    for k in range(nsrc):
        r = int(rng.integers(6, 16))
        cx, cy = rng.uniform(40, n - 40, 2)
        x1, y1 = int(cx) - r, int(cy) - r
        sl = np.s_[y1:y1 + 2 * r, x1:x1 + 2 * r]
        if seg[sl].any():
            continue
        g = np.exp(-((xx[:2 * r, :2 * r] + x1 - cx) ** 2 + (yy[:2 * r, :2 * r] + y1 - cy) ** 2)
                   / (2 * (r / 3.0) ** 2))
        amp = float(rng.uniform(50, 400))
        drz[sl] += (amp * g).astype(np.float32)
        seg[sl][g > 0.02] = k + 1
        rows.append((k + 1, cx + 1, cy + 1, amp * 30, amp * 0.3, r / 3.0, r / 3.5, 2.35 * r / 3.0,
                     float(rng.uniform(0, 1))))
    flt = np.roll(drz, (1, -1), axis=(0, 1)) + (0.05 * rng.standard_normal((n, n))).astype(np.float32) + 7.5
    wht = np.where(rng.random((n, n)) < 0.01, 0.0, 1.0).astype(np.float32)
    return drz, flt.astype(np.float32), wht, seg, rows


def main():
    drz, flt, wht, seg, rows = scene()
    out = {"frame": list(drz.shape), "catalogue_rows": len(rows)}
    with tempfile.TemporaryDirectory() as d:
        p = lambda name: os.path.join(d, name)                       # noqa: E731
        write_fits(p("drz_sci.fits"), [dict(data=drz, extra=[("EXPTIME", 1200.0), ("BUNIT", "ELECTRONS/S")])])
        write_fits(p("drz_wht.fits"), [dict(data=wht)])
        write_fits(p("img_flt.fits"), [dict(data=None, extra=[("EXPTIME", 400.0)]),
                                       dict(data=flt, extname="SCI", extver=1, extra=[("BUNIT", "ELECTRONS/S")])])
        write_fits(p("seg.fits"), [dict(data=seg)])
        names = ["NUMBER", "X_IMAGE", "Y_IMAGE", "FLUX_ISO", "FLUXERR_ISO", "A_IMAGE", "B_IMAGE",
                 "FWHM_IMAGE", "CLASS_STAR"]
        with open(p("src.cat"), "w") as f:
            for k, nme in enumerate(names):
                f.write("# %3d %-14s x\n" % (k + 1, nme))
            for r in rows:
                f.write("%d " % r[0] + " ".join(repr(float(v)) for v in r[1:]) + "\n")
        for rep in range(2):                                         # second pass: warm caches / buffers
            torch.cuda.synchronize(); t0 = time.perf_counter()
            dz = ingest.load_drizzled(p("drz_sci.fits"), p("drz_wht.fits"), dq_dtype=torch.uint8)
            im = ingest.load_image_ext(p("img_flt.fits"), ("SCI", 1), sky=7.5)
            sg = ingest.load_segmentation(p("seg.fits"))
            torch.cuda.synchronize(); t1 = time.perf_counter()
            cat = ingest.select_sources(ingest.read_sextractor_catalog(p("src.cat")))
            t2 = time.perf_counter()
            bx = batch.segment_boxes(sg).cpu().numpy()                # [label] -> x1, x2, y1, y2
            ids = cat["id"].astype(np.int64)
            b = bx[ids]
            ok = b[:, 1] >= 0
            ids, b = ids[ok], b[ok]
            pad = 3
            db = np.stack([b[:, 0] - pad, b[:, 2] - pad, b[:, 1] + pad, b[:, 3] + pad], 1)
            ib = db + np.array([-1, 1, -1, 1]) + np.array([-2, -2, 2, 2])
            t3 = time.perf_counter()
            sh, pk, st = align.frame_displacements(
                im["data"], dz["data"], ib, db, lambda x, y: (x + 1.0, y - 1.0),
                drz_seg=sg, src_ids=ids, drz_dq=dz["whtdq"], drz_dqbitmask=0)
            shifts = sh.cpu().numpy()
            torch.cuda.synchronize(); t4 = time.perf_counter()
            xy = np.stack([0.5 * (ib[:, 0] + ib[:, 2]), 0.5 * (ib[:, 1] + ib[:, 3])], 1)
            good = np.isfinite(shifts).all(1)
            fit = align.iter_linear_fit(xy[good], xy[good] + shifts[good], fitgeom="rscale",
                                        center=np.array(drz.shape[::-1]) / 2.0)
            t5 = time.perf_counter()
        out.update({"sources": int(len(ids)), "ms_ingest_4_files": (t1 - t0) * 1e3,
                    "ingest_bytes": int(drz.nbytes + wht.nbytes + flt.nbytes + seg.nbytes),
                    "ms_catalogue": (t2 - t1) * 1e3, "ms_segment_boxes_and_tables": (t3 - t2) * 1e3,
                    "ms_frame_displacements": (t4 - t3) * 1e3, "ms_linear_fit": (t5 - t4) * 1e3,
                    "ms_total": (t5 - t0) * 1e3, "fit_offset": [float(v) for v in fit["offset"]],
                    "fit_rmse": float(fit["rmse"])})
    print(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "files_to_fit.json"), "w"), indent=1)


main()
