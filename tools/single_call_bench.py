"""Developer tool: latency of the drop-in entry points called the way align.py calls them -
ONE pair per call from numpy arrays (cc.find_displacement, align.py:682-685) - next to the
reference's own function on one host core.  Writes gpurun_out/single_call.json."""
import json, os, sys, time, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload, subpixal_oracle as orc
from subpixal_b200 import cc

out = []
ref_cc = refload.load("cc") if refload.available() else None
for n in (32, 64, 128):
    stacks, _ = orc.random_pairs(8, n, n, seed=n)
    for full in (False, True):
        for k in range(8):
            cc.find_displacement(*stacks[k], cc_type="NCC", full_output=full)
        torch.cuda.synchronize()
        reps = 200
        t0 = time.perf_counter()
        for r in range(reps):
            cc.find_displacement(*stacks[r % 8], cc_type="NCC", full_output=full)
        dt = (time.perf_counter() - t0) / reps
        row = {"cutout": n, "full_output": full, "b200_us_per_call": dt * 1e6}
        if ref_cc is not None:
            with warnings.catch_warnings(), np.errstate(all="ignore"):
                warnings.simplefilter("ignore")
                t0 = time.perf_counter()
                for r in range(40):
                    ref_cc.find_displacement(*stacks[r % 8], cc_type="NCC", full_output=full)
                row["reference_us_per_call_1core"] = (time.perf_counter() - t0) / 40 * 1e6
        out.append(row)
        print(row, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "single_call.json"), "w"), indent=1)
