"""Throughput of the fused CC+fit kernel on [n,5,N,N] stacks for every supported
size class and cc_type (stack mode, inputs resident in HBM, pool > L2).
Writes gpurun_out/size_sweep.json.  Developer / profiling tool."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from subpixal_b200 import batch

def main():
    dev = torch.device("cuda")
    peak = 6540.8
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        pass
    out = []
    gen = torch.Generator(device="cuda").manual_seed(1)
    cases = ((32, 100000), (64, 25000), (128, 6000), (256, 1500), (48, 40000), (100, 8000),
             (20, 200000), (42, 50000), (84, 12000), (170, 3000))
    if len(sys.argv) > 1:
        want = [int(v) for v in sys.argv[1].split(",")]
        cases = tuple(c for c in cases if c[0] in want)
    for n_side, npairs in cases:
        stack = torch.randn((npairs, 5, n_side, n_side), device=dev, generator=gen)
        yy, xx = torch.meshgrid(torch.arange(n_side, device=dev), torch.arange(n_side, device=dev), indexing="ij")
        blob = 50 * torch.exp(-((xx - n_side / 2) ** 2 + (yy - n_side / 2) ** 2) / 18.0)
        stack += blob
        for t in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("NCC", "CC", "ZNCC")):
            for _ in range(2):
                batch.cc_fit_stack(stack, t)
            torch.cuda.synchronize()
            reps = 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                batch.cc_fit_stack(stack, t)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            rate = npairs / (ms / 1e3)
            gbs = rate * (20 * n_side * n_side + 8) / 1e9
            out.append({"n": n_side, "pairs": npairs, "cc_type": t, "ms": ms, "pairs_per_s": rate,
                        "hbm_gbs": gbs, "hbm_frac": gbs / peak})
            print(out[-1], flush=True)
        del stack
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/size_sweep.json", "w"), indent=1)

if __name__ == "__main__":
    main()
