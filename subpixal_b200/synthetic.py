"""Deterministic synthetic scenes for the benchmark and the parity tests
(SURVEY.md section 8d).  Input generation only - not part of the hot path.

A scene is a list of elliptical Gaussian sources on an ACS/WFC-like frame.
``make_frames`` renders five frames from it: the noisy "image" frame and four
noise-free "reference" frames displaced by ONE frame-level misalignment
(shift + small rotation) and sampled on grids offset by (0,0), (+1/2,0),
(0,+1/2), (+1/2,+1/2) px - the inputs cc.find_displacement expects for every
source cutout (reference cc.py:21-50, align.py:664-685).
"""
import math

import numpy as np
import torch

PHASES = ((0, 0), (1, 0), (0, 1), (1, 1))


def make_scene(n_src, height, width, cutout, seed):
    rng = np.random.default_rng(seed)
    m = cutout / 2.0 + 4.0
    return {
        "x": rng.uniform(m, width - 1 - m, n_src),
        "y": rng.uniform(m, height - 1 - m, n_src),
        "sigma": rng.uniform(1.5, 4.0, n_src),
        "q": rng.uniform(0.5, 1.0, n_src),
        "pa": rng.uniform(0.0, math.pi, n_src),
        "amp": np.exp(rng.uniform(math.log(50.0), math.log(5000.0), n_src)),
        "height": height, "width": width, "cutout": cutout, "seed": seed,
    }


def misalign(scene, shift=(0.37, -0.21), rot=5e-5):
    """Positions of the sources in the reference frames: rotation about the
    frame centre plus a shift.  Returns (x', y') and the per-source true shift
    of the reference relative to the image."""
    cx, cy = (scene["width"] - 1) / 2.0, (scene["height"] - 1) / 2.0
    dx, dy = scene["x"] - cx, scene["y"] - cy
    c, s = math.cos(rot), math.sin(rot)
    xr = cx + c * dx - s * dy + shift[0]
    yr = cy + s * dx + c * dy + shift[1]
    return xr, yr, np.stack([xr - scene["x"], yr - scene["y"]], axis=1)


def render(scene, xs, ys, device, half=21, chunk=4096):
    """Sum of elliptical Gaussians evaluated in float64, returned as float32."""
    H, W = scene["height"], scene["width"]
    out = torch.zeros(H * W, dtype=torch.float64, device=device)
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64, device=device)
    xs, ys = t(xs), t(ys)
    sig, q, pa, amp = t(scene["sigma"]), t(scene["q"]), t(scene["pa"]), t(scene["amp"])
    off = torch.arange(-half, half + 1, device=device, dtype=torch.float64)
    for a in range(0, xs.numel(), chunk):
        sl = slice(a, a + chunk)
        ix0, iy0 = torch.round(xs[sl]), torch.round(ys[sl])
        px = ix0[:, None, None] + off[None, None, :]
        py = iy0[:, None, None] + off[None, :, None]
        ddx, ddy = px - xs[sl, None, None], py - ys[sl, None, None]
        c, s = torch.cos(pa[sl])[:, None, None], torch.sin(pa[sl])[:, None, None]
        u, v = c * ddx + s * ddy, -s * ddx + c * ddy
        sg = sig[sl, None, None]
        val = amp[sl, None, None] * torch.exp(-(u * u / (2 * sg * sg) +
                                                v * v / (2 * (sg * q[sl, None, None]) ** 2)))
        ok = (px >= 0) & (px < W) & (py >= 0) & (py < H)
        idx = (py.clamp(0, H - 1) * W + px.clamp(0, W - 1)).to(torch.int64)
        out.index_put_((idx[ok],), val[ok], accumulate=True)
    return out.reshape(H, W).to(torch.float32)


def make_frames(scene, device, shift=(0.37, -0.21), rot=5e-5, noise=1.0):
    """[5, H, W] float32: image frame + the four reference-phase frames."""
    H, W = scene["height"], scene["width"]
    xr, yr, truth = misalign(scene, shift, rot)
    frames = torch.empty((5, H, W), dtype=torch.float32, device=device)
    gen = torch.Generator(device="cpu").manual_seed(int(scene["seed"]) + 17)
    frames[0] = render(scene, scene["x"], scene["y"], device)
    if noise:
        frames[0] += (noise * torch.randn((H, W), generator=gen, dtype=torch.float32)).to(device)
    for k, (a, b) in enumerate(PHASES):
        frames[1 + k] = render(scene, xr - a / 2.0, yr - b / 2.0, device)
    return frames, truth


def make_boxes(scene):
    """int32 [n, 4] (x1, y1, x2, y2): x1 = round(cx) - N/2, x2 = x1 + N - 1."""
    n = scene["cutout"]
    x1 = np.round(scene["x"]).astype(np.int64) - n // 2
    y1 = np.round(scene["y"]).astype(np.int64) - n // 2
    return np.stack([x1, y1, x1 + n - 1, y1 + n - 1], axis=1).astype(np.int32)


def cut_stack(frames_np, boxes):
    """Host-side crop of the five frames into a [n, 5, N, N] stack (numpy,
    boxes must lie inside the frame) - input for the CPU oracle."""
    n = boxes.shape[0]
    h, w = boxes[0, 3] - boxes[0, 1] + 1, boxes[0, 2] - boxes[0, 0] + 1
    out = np.empty((n, 5, h, w), dtype=np.float32)
    for k, (x1, y1, x2, y2) in enumerate(boxes):
        out[k] = frames_np[:, y1:y2 + 1, x1:x2 + 1]
    return out
