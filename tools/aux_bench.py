"""Throughput of the gather and find_peak kernels (HBM-bound byte work).
Writes gpurun_out/aux_bench.json.  Developer / profiling tool."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from subpixal_b200 import batch, synthetic as syn

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    peak = 6540.8
    try: peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception: pass
    out = []
    H, W = 4096, 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    frame = torch.randn((H, W), device="cuda", generator=g)
    dq = torch.randint(0, 4, (H, W), device="cuda", dtype=torch.int16, generator=g)
    seg = torch.randint(0, 50000, (H, W), device="cuda", dtype=torch.int32, generator=g)
    for n_side, n in ((64, 50000), (128, 20000), (32, 200000)):
        rng = np.random.default_rng(1)
        x1 = rng.integers(0, W - n_side, n); y1 = rng.integers(0, H - n_side, n)
        boxes = np.stack([x1, y1, x1 + n_side - 1, y1 + n_side - 1], 1)
        ms = timeit(lambda: batch.gather_cutouts(frame, boxes, fillval=0.0))
        # host-side table prep is inside gather_cutouts; time the launch alone as well
        by = n * n_side * n_side * (4 + 4 + 1)
        out.append({"kernel": "gather (data+mask)", "n": n, "cutout": n_side, "ms_incl_host_prep": ms,
                    "algorithmic_GB": by / 1e9, "GBps_incl_host_prep": by / ms / 1e6})
        print(out[-1], flush=True)
    # device-only timing of the gather launch through the C ABI path with prepared tables
    import ctypes as C
    from subpixal_b200 import _lib as L
    lib = L.load()
    for n_side, n in ((64, 50000), (128, 20000)):
        rng = np.random.default_rng(1)
        x1 = rng.integers(0, W - n_side, n); y1 = rng.integers(0, H - n_side, n)
        bx = torch.from_numpy(np.stack([x1, y1, x1 + n_side - 1, y1 + n_side - 1], 1).astype(np.int32)).cuda()
        off = (torch.arange(n, dtype=torch.int64) * n_side * n_side).cuda()
        ids = torch.randint(0, 50000, (n,), device="cuda", dtype=torch.int64)
        for label, use_dq in (("data+mask", False), ("data+mask+dq+seg", True)):
            data = torch.empty(n * n_side * n_side, device="cuda"); mask = torch.empty(n * n_side * n_side, device="cuda", dtype=torch.uint8)
            dqo = torch.empty(n * n_side * n_side, device="cuda", dtype=torch.int16)
            a = L.GatherArgs()
            a.frame, a.frame_kind, a.frame_h, a.frame_w, a.frame_pitch = frame.data_ptr(), 0, H, W, W
            if use_dq:
                a.dq, a.dq_kind, a.dq_out = dq.data_ptr(), L.T_I16, dqo.data_ptr()
                a.seg, a.seg_kind, a.src_ids = seg.data_ptr(), L.T_I32, ids.data_ptr()
                a.dq_use_bits, a.dq_good_bits = 1, 2
            a.boxes, a.out_offset, a.n = bx.data_ptr(), off.data_ptr(), n
            a.max_h = a.max_w = n_side; a.fillval = 0.0; a.zero_masked = 1
            a.data_out, a.mask_out = data.data_ptr(), mask.data_ptr()
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            ms = timeit(lambda: L.check(lib.spx_gather_cutouts(C.byref(a), st)))
            by = n * n_side * n_side * ((4 + 4 + 1) + ((2 + 2 + 4) if use_dq else 0))
            out.append({"kernel": "gather " + label, "n": n, "cutout": n_side, "ms": ms,
                        "algorithmic_GB": by / 1e9, "GBps": by / ms / 1e6, "hbm_frac": by / ms / 1e6 / peak})
            print(out[-1], flush=True)
    # find_peak on 128x128 fp32 surfaces
    n = 50000
    surf = torch.randn((n, 128, 128), device="cuda", generator=g)
    ms = timeit(lambda: batch.find_peak_batch(surf))
    by = n * 128 * 128 * 4
    out.append({"kernel": "find_peak 128x128 f32", "n": n, "ms": ms, "GBps": by / ms / 1e6, "hbm_frac": by / ms / 1e6 / peak})
    print(out[-1], flush=True)
    ms = timeit(lambda: batch.segment_boxes(seg))
    by = H * W * 4
    out.append({"kernel": "segment_boxes 4096^2 int32 (every pixel a random label of 50k: worst case, 4 atomics per pixel)",
                "ms": ms, "GBps": by / ms / 1e6, "hbm_frac": by / ms / 1e6 / peak})
    print(out[-1], flush=True)
    # a realistic segmentation image: 50 000 compact segments (discs of radius 3-7 px) on background 0
    rng = np.random.default_rng(5)
    segr = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    cx = torch.from_numpy(rng.integers(8, W - 8, 50000)).cuda(); cy = torch.from_numpy(rng.integers(8, H - 8, 50000)).cuda()
    rad = torch.from_numpy(rng.integers(3, 8, 50000)).cuda()
    dy, dx = torch.meshgrid(torch.arange(-7, 8, device="cuda"), torch.arange(-7, 8, device="cuda"), indexing="ij")
    inside = (dx[None] ** 2 + dy[None] ** 2) <= (rad[:, None, None] ** 2)
    yy = (cy[:, None, None] + dy[None]).expand_as(inside)[inside]; xx = (cx[:, None, None] + dx[None]).expand_as(inside)[inside]
    lab = (torch.arange(1, 50001, device="cuda", dtype=torch.int32)[:, None, None]).expand_as(inside)[inside]
    segr[yy, xx] = lab
    ms = timeit(lambda: batch.segment_boxes(segr))
    out.append({"kernel": "segment_boxes 4096^2 int32 (50k compact segments on background)", "ms": ms,
                "GBps": by / ms / 1e6, "hbm_frac": by / ms / 1e6 / peak})
    print(out[-1], flush=True)
    json.dump(out, open("gpurun_out/aux_bench.json", "w"), indent=1)

main()
