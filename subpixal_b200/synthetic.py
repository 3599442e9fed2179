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


def misalign(scene, shift=(0.37, -0.21), rot=5e-5, scale=1.0):
    """Positions of the sources in the reference frames: rotation (and isotropic
    scale) about the frame centre plus a shift.  Returns (x', y') and the per-source
    true shift of the reference relative to the image."""
    cx, cy = (scene["width"] - 1) / 2.0, (scene["height"] - 1) / 2.0
    dx, dy = scene["x"] - cx, scene["y"] - cy
    c, s = scale * math.cos(rot), scale * math.sin(rot)
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


def make_frames(scene, device, shift=(0.37, -0.21), rot=5e-5, noise=1.0, scale=1.0):
    """[5, H, W] float32: image frame + the four reference-phase frames."""
    H, W = scene["height"], scene["width"]
    xr, yr, truth = misalign(scene, shift, rot, scale)
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


def make_pair_pool(n_pairs, n, seed, device, disc_mask=False, zero_frac=0.0, chunk=None):
    """``[n_pairs, 5, n, n]`` float32 stack of synthetic cutout pairs generated on the
    device: per pair one elliptical Gaussian
    source near the cutout centre plus a fainter neighbour, unit read noise in the image
    tile (slot 0), noise-free reference tiles at the four half-pixel phases (slots 1..4)
    displaced by a per-pair true shift around (0.37, -0.21) px.

    ``disc_mask``: pixels outside a disc of radius 0.4 n around the source are 0 in the
    four phase tiles (segmentation mask + align.py:661); ``zero_frac``: that fraction of
    random phase pixels is 0 as well (the NaN pixels the gather masks and zeroes,
    cutout.py:755 + align.py:661).  Returns ``(stack, truth [n_pairs, 2])``."""
    gen = torch.Generator(device=device).manual_seed(int(seed))
    stack = torch.empty((n_pairs, 5, n, n), dtype=torch.float32, device=device)
    truth = torch.empty((n_pairs, 2), dtype=torch.float64, device=device)
    if chunk is None:
        chunk = max(1, (1 << 25) // (n * n))
    ax = torch.arange(n, device=device, dtype=torch.float32)
    yy, xx = ax[None, :, None], ax[None, None, :]
    c = (n - 1) / 2.0
    u = lambda m, lo, hi: lo + (hi - lo) * torch.rand((m, 1, 1), generator=gen, device=device)

    def blob(x0, y0, sig, q, pa, amp):
        dx, dy = xx - x0, yy - y0
        cs, sn = torch.cos(pa), torch.sin(pa)
        a, b = cs * dx + sn * dy, -sn * dx + cs * dy
        return amp * torch.exp(-(a * a / (2 * sig * sig) + b * b / (2 * (sig * q) ** 2)))

    for lo in range(0, n_pairs, chunk):
        m = min(chunk, n_pairs - lo)
        x0, y0 = c + u(m, -2.5, 2.5), c + u(m, -2.5, 2.5)
        src = (u(m, 1.5, 4.0), u(m, 0.5, 1.0), u(m, 0.0, math.pi),
               torch.exp(u(m, math.log(50.0), math.log(5000.0))))
        # a fainter neighbour somewhere in the cutout, moving coherently with the source
        x1, y1 = u(m, 0.15 * n, 0.85 * n), u(m, 0.15 * n, 0.85 * n)
        nb = (u(m, 1.5, 3.0), u(m, 0.6, 1.0), u(m, 0.0, math.pi), src[3] * u(m, 0.02, 0.2))
        sx, sy = 0.37 + u(m, -0.1, 0.1), -0.21 + u(m, -0.1, 0.1)
        out = stack[lo:lo + m]
        out[:, 0] = blob(x0, y0, *src) + blob(x1, y1, *nb) + \
            torch.randn((m, n, n), generator=gen, device=device)
        keep = None
        if disc_mask:
            keep = ((xx - x0) ** 2 + (yy - y0) ** 2) <= (0.4 * n) ** 2
        for k, (a, b) in enumerate(PHASES):
            ph = blob(x0 + sx - a / 2.0, y0 + sy - b / 2.0, *src) + \
                blob(x1 + sx - a / 2.0, y1 + sy - b / 2.0, *nb)
            if keep is not None:
                ph = ph * keep
            if zero_frac > 0.0:
                ph = ph * (torch.rand((m, n, n), generator=gen, device=device) >= zero_frac)
            out[:, 1 + k] = ph
        truth[lo:lo + m, 0] = sx.reshape(-1).double()
        truth[lo:lo + m, 1] = sy.reshape(-1).double()
    return stack, truth
