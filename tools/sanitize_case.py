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
torch.cuda.synchronize()
print("sanitize case done", float(out[0].sum()))
