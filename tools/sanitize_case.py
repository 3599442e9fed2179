"""Small end-to-end run of every kernel (for compute-sanitizer)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from subpixal_b200 import batch, synthetic as syn
sc = syn.make_scene(40, 300, 400, 64, seed=3)
frames, _ = syn.make_frames(sc, "cuda")
boxes = syn.make_boxes(sc)
boxes[:3, [0, 2]] -= boxes[:3, [0]] + 10          # partly outside the frame
boxes[3:5, [1, 3]] += 300 - boxes[3:5, [3]] + 5
frames[2, 50:60, 50:70] = float("nan")
tb = torch.from_numpy(boxes).cuda()
for t in ("NCC", "ZNCC", "CC"):
    out = batch.cc_fit_frames(frames, tb, 64, 64, t, want_icc=(t == "NCC"))
g = torch.Generator(device="cuda").manual_seed(0)
for (ny, nx, n) in ((32, 32, 12), (20, 37, 6), (128, 128, 3), (100, 90, 2), (256, 256, 1), (3, 5, 4)):
    st = torch.randn((n, 5, ny, nx), device="cuda", generator=g)
    for t in ("NCC", "ZNCC"):
        batch.cc_fit_stack(st, t, want_icc=True)
        batch.cc_fit_stack(st, t)
res = batch.gather_cutouts(frames[0], np.array([[-5, -5, 40, 30], [380, 280, 420, 310], [10, 10, 73, 73]]),
                           dq=torch.zeros((300, 400), dtype=torch.int16, device="cuda"),
                           seg=torch.ones((300, 400), dtype=torch.int32, device="cuda"), src_ids=[1, 1, 0],
                           dq_good_bits=0, zero_masked=True)
batch.find_peak_batch(torch.randn((5, 33, 47), device="cuda", generator=g))
batch.find_peak_batch(torch.randn((5, 33, 47), device="cuda", generator=g, dtype=torch.float64),
                      guess=torch.tensor([[3.0, 4.0]] * 5), search_box=(5, 7), fit_box=(3, 5))
batch.segment_boxes(torch.randint(0, 9, (70, 90), device="cuda", dtype=torch.int32, generator=g))
# new size classes, ragged batches, blot (direct / prepared / linear / nearest), device pipeline
for (ny, nx, n) in ((20, 20, 5), (13, 19, 3), (42, 42, 4), (33, 40, 3), (84, 84, 2), (70, 66, 2)):
    st = torch.randn((n, 5, ny, nx), device="cuda", generator=g)
    batch.cc_fit_stack(st, "NCC", want_icc=True)
shapes = [(3, 5), (20, 37), (64, 64), (100, 90), (21, 19), (85, 60), (130, 140)]
flat = torch.randn(sum(5 * a * b for a, b in shapes), device="cuda", generator=g)
offs = np.concatenate([[0], np.cumsum([5 * a * b for a, b in shapes])])[:-1]
batch.cc_fit_ragged(flat, offs, shapes, "ZNCC", want_icc=True)
batch.cc_fit_ragged(flat, offs, shapes, "NCC")
from subpixal_b200 import align, blot
rng = np.random.default_rng(0)
srcs = [(5, 4), (6, 6), (40, 33), (64, 80), (3, 70)]
sflat = torch.randn(sum(a * b for a, b in srcs), device="cuda", generator=g)
soff = np.concatenate([[0], np.cumsum([a * b for a, b in srcs])])[:-1]
tg = [(7, 9), (6, 6), (32, 32), (64, 64), (16, 50)]
def mk(k):
    h, w = srcs[k]; ny, nx = tg[k]
    return blot.fit_pixel_map(lambda x, y: (1 + (x - 1) * (w + 3.0) / nx - 1.5, 1 + (y - 1) * (h + 3.0) / ny - 1.5), nx, ny)
maps = np.stack([mk(k) for k in range(len(srcs))])
for interp, prep in (("poly5", True), ("poly5", False), ("linear", False), ("nearest", False)):
    batch.blot_phases(sflat, soff, srcs, maps, tg, interp=interp, prepare=prep)
ib = np.array([[-5, -5, 26, 20], [380, 270, 420, 310], [10, 10, 73, 73], [100, 100, 119, 119]])
db = ib + np.array([-6, -6, 6, 6])
seg = torch.randint(0, 3, (300, 400), device="cuda", dtype=torch.int32, generator=g)
align.frame_displacements(frames[0], frames[1], ib, db, lambda x, y: (x + 0.3, y - 0.2),
                          drz_seg=seg, src_ids=[1, 1, 2, 0], full_output=True)
align.frame_displacements(frames[0], frames[1], ib, db, lambda x, y: (x + 0.3, y - 0.2))
torch.cuda.synchronize()
print("sanitize case done", float(out[0].sum()))
