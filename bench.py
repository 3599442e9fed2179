#!/usr/bin/env python
"""Benchmark of the subpixal hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): 10 000 cutout pairs of 64x64 fp32 taken
from one synthetic 4096x2048 ACS/WFC-like frame (+ the four reference-phase
frames), cc_type='NCC', full_output=False.  One "step" = one pass of the fused
gather + cross-correlation + sub-pixel-fit kernel over the 10 000 pairs (per
GPU; weak scaling), followed for N > 1 by the single NCCL all-gather of the
shift table.  Prints ONE JSON line on rank 0.

The line also carries a ``sizes`` block - the other half of the metric and
configs[2] / configs[4]:

* ``"128"``: config 3 - 100 000 pairs 128x128 in stack mode, pixels outside a disc of
  radius 0.4 N zeroed in the four phase tiles + 0.1 % zeroed phase pixels, NCC and ZNCC;
* ``"32"``, ``"64"``, ``"256"``: config 5 - 1 M pairs per size in stack mode, cycled over a
  device-resident pool of >= 8 GB (>> 126 MB of L2);

each with pairs/s, roofline (HBM fraction, ncu DRAM traffic from profiles/), the fp32
FLOP fraction for the transform size used, and a parity block against the CPU
reference on the same tiles.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cutout pairs/sec (CC+subpixel fit)"
UNIT = "pairs/s"
FRAME_H, FRAME_W = 2048, 4096
L2_FLUSH_BYTES = 256 << 20

# fp32 peak of the CUDA cores: 148 SMs x 128 lanes x 2 flop (FMA) x SM clock
FP32_LANES = 148 * 128
# flop per pair of the transforms at the alias-free size P = 1.5 N
MFLOP_PER_PAIR = {32: 0.61, 64: 2.84, 128: 13.0, 256: 58.7}
# pairs compared with the CPU reference per size
PARITY_PAIRS = {32: 2000, 64: 2000, 128: 500, 256: 100}
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def algorithmic_bytes(n):
    """five NxN fp32 tiles read once, (dx,dy) written."""
    return 20 * n * n + 8


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(n, mode, pairs):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch of
    ``pairs`` pairs, scaled from the committed ``ncu --set full`` capture of that size
    class (profiles/ncu_traffic.json names the capture); None when there is none."""
    try:
        with open(TRAFFIC_FILE) as f:
            ent = json.load(f)["%d:%s" % (n, mode)]
        return int(round(ent["dram_bytes_per_pair"] * pairs)), ent["source"]
    except Exception:
        return None, None


# ----------------------------------------------------------------------------
# CPU arm: the reference's own cc.find_displacement (UNMODIFIED sources, loaded by
# oracle/refload.py from /root/reference or from the staged copy oracle/_ref made by
# oracle/make_ref.py) on all host cores - kind "reference"; the oracle port
# (numpy + scipy.signal.fftconvolve + numpy.linalg.lstsq, the libraries the reference
# calls) when neither tree is present - kind "port".
# ----------------------------------------------------------------------------
_W_STACK = None
_W_FD = None


def cpu_kind():
    from oracle import refload
    return "reference" if refload.available() else "port"


def _pool_init(path):
    global _W_STACK, _W_FD
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    import numpy as np
    from oracle import refload
    if refload.available():
        _W_FD = refload.load("cc").find_displacement
    else:
        from oracle import subpixal_oracle as orc
        _W_FD = orc.find_displacement
    _W_STACK = np.load(path, mmap_mode="r")


def _pool_work(job):
    import warnings
    import numpy as np
    lo, hi, cc_type, full = job
    out = []
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for k in range(lo, hi):
            st = np.asarray(_W_STACK[k])
            if not full:
                out.append(tuple(float(v) for v in _W_FD(st[0], st[1], st[2], st[3], st[4],
                                                         cc_type=cc_type)))
                continue
            dx, dy, icc, _ = _W_FD(st[0], st[1], st[2], st[3], st[4], cc_type=cc_type,
                                   full_output=True)
            flat = icc.ravel().astype(np.float64)
            am = int(np.argmax(flat))
            top = flat[am]
            flat[am] = -np.inf
            second = float(np.max(flat))
            tie = bool(top - second <= 1e-5 * abs(top))
            out.append((float(dx), float(dy), am, tie))
    return out


class CpuArm(object):
    def __init__(self, stack):
        import multiprocessing as mp
        import numpy as np
        self.cores = os.cpu_count() or 1
        self.kind = cpu_kind()
        shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
        self.path = os.path.join(shm, "spx_bench_sample_%d_%d.npy" % (os.getpid(), id(self)))
        np.save(self.path, stack)
        self.n = stack.shape[0]
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_pool_init,
                                                 initargs=(self.path,))
        self.run(0, min(self.n, 2 * self.cores))          # import + warm caches

    def run(self, lo, hi, cc_type="NCC", full=False):
        """Process pairs [lo, hi) on all cores; returns (seconds, results)."""
        chunk = max(1, (hi - lo + 4 * self.cores - 1) // (4 * self.cores))
        tasks = [(a, min(hi, a + chunk), cc_type, full) for a in range(lo, hi, chunk)]
        t0 = time.perf_counter()
        res = self.pool.map(_pool_work, tasks)
        dt = time.perf_counter() - t0
        return dt, [r for part in res for r in part]

    def close(self):
        self.pool.close()
        self.pool.join()
        try:
            os.remove(self.path)
        except OSError:
            pass

    def describe(self):
        return ("unmodified reference cc.find_displacement (oracle/refload.py)"
                if self.kind == "reference" else
                "oracle port = numpy + scipy.signal.fftconvolve + numpy.linalg.lstsq")


# ----------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ----------------------------------------------------------------------------
def run_reference(args, rank, world, emit):
    """--impl reference: the reference's CPU path on all host cores (rank 0 only)."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from subpixal_b200 import synthetic as syn
    n = args.size
    cores = os.cpu_count() or 1
    per_step = args.ref_pairs_per_core * cores
    need = min(args.pairs, per_step * 2)
    sc = syn.make_scene(args.pairs, FRAME_H, FRAME_W, n, seed=args.seed)
    # render only what the sample needs: the scene restricted to a band of sources
    # would change the neighbours, so render the full frames once (CPU, torch)
    frames, _ = syn.make_frames(sc, "cpu")
    boxes = syn.make_boxes(sc)
    stack = syn.cut_stack(frames.numpy(), boxes[:need])
    arm = CpuArm(stack)
    try:
        for w in range(args.warmup):
            arm.run(0, min(need, per_step))
        total_t, total_n = 0.0, 0
        for k in range(args.steps):
            lo = (k % 2) * per_step if need >= 2 * per_step else 0
            hi = min(need, lo + per_step)
            dt, _ = arm.run(lo, hi)
            total_t += dt
            total_n += hi - lo
    finally:
        arm.close()
    value = total_n / total_t
    sample = "%d pairs of %dx%d per step on %d cores (multiprocessing pool; %s)" % (
        min(need, per_step), n, n, cores, arm.describe())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": arm.kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world):
    return {
        "workload": "%d cutout pairs %dx%d fp32 from one synthetic %dx%d frame "
                    "(+4 reference-phase frames), cc_type=NCC, full_output=False; "
                    "per GPU" % (args.pairs, args.size, args.size, FRAME_W, FRAME_H),
        "pairs_per_gpu": args.pairs, "cutout": args.size, "cc_type": "NCC",
        "frame": [FRAME_H, FRAME_W], "sharding": "sources sharded by rank, frames per rank",
        "collective": "1 all_gather of the shift table per step" if world > 1 else "none",
        "l2": ("256 MiB buffer written between timed steps (outside the CUDA-event windows); "
               "the five frames (168 MB) also exceed the 126 MB L2") if world == 1 else
              ("inputs larger than L2: the five frames (168 MB) exceed the 126 MB L2; the K "
               "steps are one timed region, the all-gather of step i overlapping the kernel "
               "of step i+1"),
        "seed": args.seed,
    }


# ----------------------------------------------------------------------------
# sizes block: configs 3 and 5
# ----------------------------------------------------------------------------
SIZE_PLAN = {
    # size: (config, total pairs per GPU, cc_types, disc mask, zeroed fraction)
    128: ("config 3: 100k pairs 128x128 with per-cutout masking/normalisation", 100000,
          ("NCC", "ZNCC"), True, 0.001),
    32: ("config 5: 1M pairs 32x32 (stack mode, pool cycled)", 1000000, ("NCC",), False, 0.0),
    64: ("config 5: 1M pairs 64x64 (stack mode, pool cycled)", 1000000, ("NCC",), False, 0.0),
    256: ("config 5: 1M pairs 256x256 (stack mode, pool cycled)", 1000000, ("NCC",), False, 0.0),
}


def bench_size(n, args, rank, world, dev, get_gather, sm_mhz):
    """One entry of the ``sizes`` block (every rank takes part; rank 0 returns the dict)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from subpixal_b200 import batch, synthetic as syn
    label, total, cc_types, disc, zfrac = SIZE_PLAN[n]
    total = max(1, int(total * args.sizes_scale))
    bpp = algorithmic_bytes(n)
    pool_pairs = int(min(total, max(PARITY_PAIRS[n], -(-args.pool_bytes // bpp))))
    stack, _ = syn.make_pair_pool(pool_pairs, n, args.seed + 1000 * n + rank, dev,
                                  disc_mask=disc, zero_frac=zfrac)
    launches = -(-total // pool_pairs)              # the pool is cycled until `total` pairs
    done = launches * pool_pairs
    peak, peak_src = measured_peak()
    entry = {"config": label, "mode": "stack [n,5,N,N], device resident",
             "pairs_per_gpu": done, "pool_pairs": pool_pairs,
             "pool_bytes": int(stack.numel() * 4), "launches_per_pass": launches,
             "l2": "pool >> 126 MB L2: every launch streams its inputs from HBM",
             "masking": ("phase tiles zero outside a disc of radius 0.4 N + %.1f %% zeroed "
                         "phase pixels" % (100 * zfrac)) if disc else "none",
             "cc": {}}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in cc_types:
        out = None
        for _ in range(2):                                       # warm-up launches
            out = batch.cc_fit_stack(stack[:min(pool_pairs, 4096)], t)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []                                   # N = 1: one event per launch (spread of the launches)
        sampler = ClockSampler(physical_gpu_index(int(os.environ.get("LOCAL_RANK", 0)))) if rank == 0 else None
        barrier()
        if sampler:
            sampler.start()
        e0.record()
        g = get_gather(pool_pairs) if world > 1 else None
        for _ in range(launches):
            # N > 1: the kernel writes this rank's packed result rows into the persistent
            # send buffer, one all-gather returns the full table to every rank
            out = batch.cc_fit_stack(stack, t, table=g.local if g else None)
            if g:
                g.gather()
            elif launches > 1:
                marks.append(torch.cuda.Event(enable_timing=True))
                marks[-1].record()
        e1.record()
        barrier()
        size_clocks = sampler.finish() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        rate = world * done / (ms / 1e3)
        per_gpu = rate / world
        achieved = per_gpu * bpp / 1e9
        traffic, tsrc = ncu_traffic(n, "stack", pool_pairs)
        flop = MFLOP_PER_PAIR[n] * 1e6
        fp32_peak = FP32_LANES * 2 * (sm_mhz or 1965.0) * 1e6
        per_launch = ([e0.elapsed_time(marks[0])] +
                      [a.elapsed_time(b) for a, b in zip(marks[:-1], marks[1:])]) if marks else None
        res = {"pairs_per_s": rate, "ms_per_launch": ms / launches,
               "ms_per_launch_min_max": [min(per_launch), max(per_launch)] if per_launch else None,
               "clocks": size_clocks,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc,
                            "kernel": "cc_fit_kernel<Cfg%d>" % n, "bytes_per_pair": bpp,
                            "pairs_per_launch": pool_pairs, "peak_source": peak_src},
               "fp32_flop_frac": per_gpu * flop / fp32_peak,
               "fp32": {"mflop_per_pair": MFLOP_PER_PAIR[n], "transform_size": 3 * n // 2,
                        "peak_tflops": fp32_peak / 1e12,
                        "peak_formula": "148 SMs x 128 lanes x 2 x %.0f MHz" % (sm_mhz or 1965.0)}}
        # ---- parity against the CPU reference on the SAME tiles (rank 0) ----------
        if rank == 0 and not args.no_cpu_baseline:
            m = min(pool_pairs, PARITY_PAIRS[n])
            host = stack[:m].cpu().numpy()
            sh = out[0][:m].cpu().numpy()
            pk = out[1][:m].cpu().numpy()
            arm = CpuArm(host)
            try:
                dt, ref = arm.run(0, m, cc_type=t, full=True)
            finally:
                arm.close()
            rdx = np.array([[r[0], r[1]] for r in ref])
            ram = np.array([r[2] for r in ref])
            tie = np.array([r[3] for r in ref])
            same = ram == pk
            diff = np.abs(sh - rdx)
            res["parity"] = {
                "n_compared": int(m), "against": arm.kind,
                "max_abs_shift_diff": float(diff[same].max()) if same.any() else None,
                "argmax_mismatches": int((~same).sum()),
                "argmax_mismatches_not_near_tie": int((~same & ~tie).sum()),
                "near_tie_rate": float(tie.mean()),
                "tolerance_px": 1e-3,
                "ok": bool((~same & ~tie).sum() == 0 and
                           (not same.any() or float(diff[same].max()) < 1e-3)),
                "cpu_pairs_per_s": m / dt, "cpu_cores": arm.cores}
        entry["cc"][t] = res
    first = entry["cc"][cc_types[0]]
    for k in ("pairs_per_s", "roofline", "fp32_flop_frac", "parity"):
        if k in first:
            entry[k] = first[k]                     # the NCC figures at the top level
    del stack
    torch.cuda.empty_cache()
    return entry


# ----------------------------------------------------------------------------
# config 4: the align.py workflow - 8 frames 4096x4096, one shared 50 000-source scene,
# shifts -> ONE all-gather -> linear-transformation fit per frame
# ----------------------------------------------------------------------------
def config4_truth(f):
    """Injected misalignment of frame f: shift <= 0.5 px,
    rotation <= 1e-4 rad, scale 1 +- 1e-5."""
    import numpy as np
    rng = np.random.default_rng(20244 + 101 * f)
    return (tuple(rng.uniform(-0.5, 0.5, 2)), float(rng.uniform(-1e-4, 1e-4)),
            float(1.0 + rng.uniform(-1e-5, 1e-5)))


def bench_config4(args, rank, world, dev, get_gather):
    """Frames are dealt to the ranks round robin (one frame per GPU at N = 8); every rank
    measures all sources of its frames with the fused gather + CC + fit kernel, writes the
    packed result rows into the persistent send buffer, one all-gather hands every rank the
    full [8 x 50 000] table, rank 0 fits the 8 transforms (align.iter_linear_fit, the
    interface consumed at align.py:720-724) and checks them against the injected truth and
    against the same fit fed with the CPU reference's shifts on a 2048-source subsample."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from subpixal_b200 import align, batch, synthetic as syn
    n_frames, n_src, n, side = 8, max(64, int(50000 * args.sizes_scale)), 64, 4096
    if n_frames % world:
        return None
    sc = syn.make_scene(n_src, side, side, n, seed=20244)            # the shared scene
    boxes_np = syn.make_boxes(sc)
    boxes = torch.from_numpy(boxes_np).to(dev)
    mine = [f for f in range(n_frames) if f % world == rank]
    frames = {f: syn.make_frames(sc, dev, *config4_truth(f)[:2], scale=config4_truth(f)[2])[0]
              for f in mine}
    g = get_gather(len(mine) * n_src)

    def step():
        for j, f in enumerate(mine):
            batch.cc_fit_frames(frames[f], boxes, n, n, "NCC",
                                table=g.local[j * n_src:(j + 1) * n_src])
        return g.gather()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step()
    barrier()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sh, pk, st = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    ms = float(ms.item())
    table = sh.cpu().numpy().reshape(world, len(mine), n_src, 2)     # [rank, slot, source]
    xy = np.stack([sc["x"], sc["y"]], axis=1)
    centre = np.array([(side - 1) / 2.0] * 2)
    worst = {"offset_px": 0.0, "rot_rad": 0.0, "scale": 0.0, "rmse_px": 0.0}
    for f in range(n_frames):
        shifts = table[f % world, f // world]
        fit = align.iter_linear_fit(xy, xy + shifts, fitgeom="rscale", center=centre, nclip=3, sigma=3.0)
        t_shift, t_rot, t_scale = config4_truth(f)
        worst["offset_px"] = max(worst["offset_px"], float(np.max(np.abs(fit["offset"] - t_shift))))
        rot = (np.deg2rad(fit["rot"]) + np.pi) % (2 * np.pi) - np.pi
        worst["rot_rad"] = max(worst["rot_rad"], abs(float(rot) - t_rot))
        worst["scale"] = max(worst["scale"], abs(fit["scale"][0] - t_scale))
        worst["rmse_px"] = max(worst["rmse_px"], float(fit["rmse"]))
    out = {"config": "config 4: 8 frames 4096x4096, %d shared sources, 64x64 cutouts, NCC; frames "
                     "dealt round robin to %d GPU(s); 1 all-gather; rscale fit per frame on rank 0"
                     % (n_src, world),
           "ms_per_pass": ms, "pairs_per_s": n_frames * n_src / (ms / 1e3),
           "fit_vs_truth_max_over_frames": worst,
           "fit_tolerances": {"offset_px": 0.15, "rot_rad": 2e-5, "scale": 2e-5,
                              "note": "offset: the bias of the method itself near half-pixel shifts "
                                      "(5x5 quadratic fit on blended neighbours); the reference has "
                                      "the same bias - see parity.fit_offset_diff_px"},
           "fit_ok": bool(worst["offset_px"] < 0.15 and worst["rot_rad"] < 2e-5 and
                          worst["scale"] < 2e-5)}
    if not args.no_cpu_baseline:
        idx = np.arange(0, n_src, max(1, n_src // 2048))[:2048]
        f0 = mine[0]
        stack = syn.cut_stack(frames[f0].cpu().numpy(), boxes_np[idx])
        arm = CpuArm(stack)
        try:
            _, ref = arm.run(0, len(idx))
        finally:
            arm.close()
        ref = np.array(ref)
        got = table[f0 % world, f0 // world][idx]
        f_cpu = align.iter_linear_fit(xy[idx], xy[idx] + ref, fitgeom="rscale", center=centre)
        f_gpu = align.iter_linear_fit(xy[idx], xy[idx] + got, fitgeom="rscale", center=centre)
        out["parity"] = {"n_compared": int(len(idx)), "against": arm.kind, "frame": f0,
                         "max_abs_shift_diff": float(np.max(np.abs(ref - got))),
                         "fit_offset_diff_px": float(np.max(np.abs(f_cpu["offset"] - f_gpu["offset"]))),
                         "fit_matrix_diff": float(np.max(np.abs(f_cpu["matrix"] - f_gpu["matrix"]))),
                         "ok": bool(np.max(np.abs(ref - got)) < 1e-3 and
                                    np.max(np.abs(f_cpu["offset"] - f_gpu["offset"])) < 1e-5)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--pairs", type=int, default=10000)
    ap.add_argument("--seed", type=int, default=20242)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-pairs-per-core", type=int, default=192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sizes", default="128,32,64,256",
                    help="comma-separated sizes of the `sizes` block ('' = none)")
    ap.add_argument("--sizes-scale", type=float, default=1.0,
                    help="scale the pairs per size (1.0 = the BASELINE.json counts)")
    ap.add_argument("--pool-bytes", type=int, default=8 << 30)
    ap.add_argument("--no-config4", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout must carry exactly ONE JSON line: libraries that print to fd 1
    # (e.g. NCCL's version banner) are sent to stderr until the line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
        args.gpus = world

    import numpy as np
    import torch
    import torch.distributed as dist
    from subpixal_b200 import batch, dist as sdist, synthetic as syn

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, npairs = args.size, args.pairs
    sc = syn.make_scene(npairs, FRAME_H, FRAME_W, n, seed=args.seed + rank)
    frames, truth = syn.make_frames(sc, dev)
    boxes_np = syn.make_boxes(sc)
    boxes = torch.from_numpy(boxes_np).to(dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    gatherers = {}

    def get_gather(per):
        """Persistent send / receive buffers of the ONE collective of the path (the
        all-gather of the per-rank result table), one set per table length."""
        g = gatherers.get(per)
        if g is None:
            g = gatherers[per] = sdist.ResultGather(per, world, dev, depth=2)
        return g

    def step_device(fr=None, bx=None, pipelined=False):
        """One step.  N > 1: the kernel writes its packed result rows into a persistent send
        buffer and ONE all-gather returns the table to every rank; `pipelined`: the
        all-gather is only enqueued (NCCL's stream) and overlaps the next step's kernel -
        two buffer sets, every table complete after get_gather(npairs).wait()."""
        fr, bx = (frames, boxes) if fr is None else (fr, bx)
        if world > 1:
            g = get_gather(npairs)
            table = g.advance()
            batch.cc_fit_frames(fr, bx, n, n, "NCC", table=table)      # packed rows, no glue
            return g.gather(async_op=pipelined)
        out = batch.cc_fit_frames(fr, bx, n, n, "NCC")
        return out[0], out[1], out[2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step_device()
    barrier()
    assert bool(torch.isfinite(out[0]).all()), "non-finite shifts in warm-up"

    sampler = ClockSampler(physical_gpu_index(local_rank)) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    if world == 1:
        for a, b in ev:
            flush.fill_(1)                  # evict L2 (not inside the event window)
            a.record()
            step_device()
            b.record()
        barrier()
        ms_steps = [a.elapsed_time(b) for a, b in ev]
    else:
        # N > 1: the K steps are timed as ONE region (first kernel launch -> last all-gather
        # complete) because the all-gather of step i overlaps the kernel of step i + 1; the
        # five frames (168 MB) exceed the L2, so no flush between the steps
        flush.fill_(1)
        ev[0][0].record()
        for _ in range(args.steps):
            step_device(pipelined=True)
        get_gather(npairs).wait()
        ev[0][1].record()
        barrier()
        ms_steps = [ev[0][0].elapsed_time(ev[0][1]) / args.steps] * args.steps
    clocks = sampler.finish() if sampler else None
    ms_total = torch.tensor([sum(ms_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    value = world * npairs * args.steps / (ms_total / 1e3)

    # ---- end to end: host buffers in, host results out, every step ---------
    # Every step copies ITS five frames + box table from pinned host memory and
    # copies ITS shift table back.  Two device buffer sets: the upload of step
    # i+1 (copy stream) overlaps the kernel of step i (compute stream), as a
    # real pipeline over many exposures would; nothing is reused between steps.
    h_frames = frames.cpu().pin_memory()
    h_boxes = torch.from_numpy(boxes_np).pin_memory()
    n_out = npairs * world if world > 1 else npairs
    bufs = [(torch.empty_like(frames), torch.empty_like(boxes),
             torch.empty((n_out, 2), dtype=torch.float64).pin_memory()) for _ in range(2)]
    s_copy, s_comp = torch.cuda.Stream(), torch.cuda.Stream()
    up_done = [torch.cuda.Event() for _ in range(2)]
    buf_free = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        d_fr, d_bx, _ = bufs[i % 2]
        with torch.cuda.stream(s_copy):
            s_copy.wait_event(buf_free[i % 2])
            d_fr.copy_(h_frames, non_blocking=True)
            d_bx.copy_(h_boxes, non_blocking=True)
            up_done[i % 2].record(s_copy)

    def compute(i):
        d_fr, d_bx, h_out = bufs[i % 2]
        with torch.cuda.stream(s_comp):
            s_comp.wait_event(up_done[i % 2])
            sh = step_device(d_fr, d_bx)[0]
            h_out.copy_(sh, non_blocking=True)
            buf_free[i % 2].record(s_comp)
        return h_out

    def run_e2e(k):
        upload(0)
        for i in range(k):
            if i + 1 < k:
                upload(i + 1)
            out = compute(i)
        s_comp.synchronize()
        return out

    e2e_steps = max(4, min(args.steps, 20))
    for ev_ in buf_free:
        ev_.record()
    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = run_e2e(e2e_steps)
    torch.cuda.current_stream().wait_stream(s_comp)
    e1.record()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    assert bool(torch.isfinite(last).all())
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    e2e_value = world * npairs * e2e_steps / (e2e_ms / 1e3)
    h2d = int(h_frames.numel() * 4 + h_boxes.numel() * 4)
    d2h = int(n_out * 16)
    h2d_gbs_rank = h2d * e2e_steps / (e2e_ms / 1e3) / 1e9
    del h_frames, bufs

    # ---- roofline of the dominant kernel (rank 0's own launches) ----------
    peak, peak_src = measured_peak()
    kern_ms = statistics.mean(ms_steps)
    achieved = npairs * algorithmic_bytes(n) / (kern_ms / 1e3) / 1e9
    traffic, traffic_src = ncu_traffic(n, "frames", npairs)
    sm_mhz = (clocks or {}).get("sm_mhz")
    fp32_peak = FP32_LANES * 2 * (sm_mhz or 1965.0) * 1e6
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "cc_fit_kernel<Cfg%d>" % n, "peak_source": peak_src,
                "bytes_per_pair": algorithmic_bytes(n), "pairs_per_launch": npairs,
                "launch_ms": kern_ms,
                "fp32_flop_frac": (npairs / (kern_ms / 1e3)) * MFLOP_PER_PAIR.get(n, 0.0) * 1e6
                / fp32_peak,
                "note": "HBM fraction as BASELINE.json defines it; the kernel is bound by fp32 "
                        "issue / shared-memory bandwidth (SURVEY.md section 8d), see DESIGN.md"}

    # ---- CPU baseline: the reference on this box's host cores (rank 0, N=1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        want = int(min(npairs, max(8 * cores, 4000)))
        stack = syn.cut_stack(frames.cpu().numpy(), boxes_np[:want])
        arm = CpuArm(stack)
        try:
            dt, res, m, passes = 0.0, None, 0, 0
            while dt < args.cpu_seconds:             # bounded sample: ~10-30 s of CPU work
                d, r = arm.run(0, want)
                dt += d
                m += want
                passes += 1
                res = r
        finally:
            arm.close()
        # the same pairs on the GPU: parity of the benchmarked configuration
        sh = step_device()[0][:want].cpu().numpy()
        ref = np.array(res)
        cpu = {"value": m / dt, "unit": UNIT, "cores": cores, "kind": arm.kind,
               "sample": "first %d of the %d pairs, %d passes = %d pairs in %.1f s wall "
                         "(multiprocessing pool on all cores; %s)"
                         % (want, npairs, passes, m, dt, arm.describe()),
               "max_abs_shift_diff_vs_gpu_px": float(np.max(np.abs(sh - ref)))}

    # ---- the other sizes of the metric: configs 3 and 5 --------------------------
    sizes = {}
    del frames
    torch.cuda.empty_cache()
    for s in [int(v) for v in args.sizes.split(",") if v.strip()]:
        if s not in SIZE_PLAN:
            continue
        entry = bench_size(s, args, rank, world, dev, get_gather, sm_mhz)
        if rank == 0:
            sizes[str(s)] = entry

    config4 = None
    if not args.no_config4:
        config4 = bench_config4(args, rank, world, dev, get_gather)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "wall_ms_per_step": wall_ms / e2e_steps,
                    "h2d_gbs_per_rank": h2d_gbs_rank,
                    "limiter": "host->device copy of the five frames (168 MB per step and rank) "
                               "over PCIe: %.1f GB/s per rank, %.0f GB/s for the %d ranks of this "
                               "host; the kernel needs %.2f ms of the %.2f ms step"
                               % (h2d_gbs_rank, h2d_gbs_rank * world, world, kern_ms,
                                  e2e_ms / e2e_steps),
                    "note": "pinned host -> device copy of the 5 frames + boxes and device -> "
                            "host copy of the shifts every step; upload of step i+1 overlaps "
                            "the kernel of step i (two buffer sets, two streams)"},
            "gpu_launches": args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "sizes": sizes, "config4": config4,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
