"""Batched, tensor-level entry points over the C-ABI (the throughput path).

All inputs and outputs are ``torch`` tensors on a CUDA device; the calls only
enqueue work on the current stream.  These are what the reference-signature
wrappers in ``cc.py`` / ``centroid.py`` / ``cutout.py`` and ``dist.py`` use.
"""
import ctypes as C

import torch

from . import _lib as L

_CC_TYPES = {"CC": L.CC, "NCC": L.NCC, "ZNCC": L.ZNCC}
_workspaces = {}

_KIND = {torch.float32: L.T_F32, torch.float64: L.T_F64, torch.int8: L.T_I8,
         torch.uint8: L.T_I8, torch.int16: L.T_I16, torch.int32: L.T_I32,
         torch.int64: L.T_I64}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("subpixal_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")


def cc_type_code(cc_type):
    """cc.py:107-108: case-insensitive; anything unknown behaves as 'CC'."""
    return _CC_TYPES.get(str(cc_type).upper(), L.CC)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    nbytes = max(int(nbytes), 1 << 20)
    if ws is None or ws.numel() < nbytes or (ws.numel() > _SCRATCH_SLACK * nbytes and
                                             ws.numel() > (64 << 20)):
        ws = None
        _workspaces.pop(key, None)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_scratch_bufs = {}
_SCRATCH_SLACK = 4        # a cached buffer this many times larger than the request is dropped


def scratch_f32(name, numel, device):
    """float32 scratch tensor per (name, device, stream): internal buffers of the
    multi-kernel pipelines are reused across calls instead of going through the
    allocator (reuse is ordered by the stream the kernels run on).  A cached buffer
    that is too small, or more than ``_SCRATCH_SLACK`` times larger than the request
    (and above 64 MiB), is handed back to torch's caching allocator; ``clear_scratch()``
    releases everything."""
    key = (name, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch_bufs.get(key)
    numel = max(int(numel), 1024)
    if buf is not None and (buf.numel() < numel or
                            (buf.numel() > _SCRATCH_SLACK * numel and buf.numel() > (16 << 20))):
        buf = None
        _scratch_bufs.pop(key, None)
    if buf is None:
        buf = torch.empty(numel, dtype=torch.float32, device=device)
        _scratch_bufs[key] = buf
    return buf


def clear_scratch():
    """Release the cached scratch buffers and kernel workspaces of every device and
    stream (they return to torch's caching allocator; ``torch.cuda.empty_cache()``
    then gives the memory back to the driver)."""
    _scratch_bufs.clear()
    _workspaces.clear()


def _f32c(t, name):
    if not t.is_cuda:
        raise ValueError(name + " must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise TypeError(name + " must be float32")
    return t.contiguous()


def _outputs(n, ny, nx, want_icc, device, table=None):
    """Output tensors of a cc_fit call.  With ``table`` (a caller-owned ``[>= n, 4]``
    float64 CUDA tensor, e.g. the persistent send buffer of ``dist.ResultGather``) the
    kernel writes the packed rows (dx, dy, peak, status) itself and the returned
    shifts / peak / status are float64 VIEWS of its first n rows."""
    if table is not None:
        if (table.dtype != torch.float64 or table.dim() != 2 or table.shape[1] != 4 or
                table.shape[0] < n or not table.is_contiguous() or table.device != device):
            raise ValueError("table must be a contiguous [>= n, 4] float64 tensor on the device")
        shifts, peak, status = table[:n, :2], table[:n, 2], table[:n, 3]
    else:
        shifts = torch.empty((n, 2), dtype=torch.float64, device=device)
        peak = torch.empty((n,), dtype=torch.int32, device=device)
        status = torch.empty((n,), dtype=torch.int32, device=device)
    icc = (torch.empty((n, 2 * ny, 2 * nx), dtype=torch.float32, device=device)
           if want_icc else None)
    return shifts, peak, status, icc


def _out_ptrs(shifts, peak, status, table):
    """(shifts, peak, status) pointers of the C ABI; peak = status = NULL selects the
    packed table form."""
    if table is not None:
        return _ptr(table), C.c_void_p(0), C.c_void_p(0)
    return _ptr(shifts), _ptr(peak), _ptr(status)


def cc_fit_stack(stack, cc_type="NCC", want_icc=False, table=None):
    """Cross-correlate + fit n pairs stored as a ``[n, 5, ny, nx]`` float32
    stack (slot 0 = ref_image, 1..4 = image00, image10, image01, image11).
    Returns ``(shifts [n,2] f64, peak [n] i32, status [n] i32, icc or None)``.
    Batched form of cc.find_displacement (reference cc.py:21-95).  ``table``: see
    ``_outputs`` (packed [n,4] float64 result rows written by the kernel)."""
    require_cuda()
    lib = L.load()
    stack = _f32c(stack, "stack")
    if stack.dim() != 4 or stack.shape[1] != 5:
        raise ValueError("stack must have shape [n, 5, ny, nx]")
    n, _, ny, nx = (int(v) for v in stack.shape)
    dev = stack.device
    with torch.cuda.device(dev):
        shifts, peak, status, icc = _outputs(n, ny, nx, want_icc, dev, table)
        if n == 0:
            return shifts, peak, status, icc
        nb = lib.spx_cc_workspace_bytes(n, ny, nx, int(want_icc))
        ws = _workspace(nb, dev)
        tile = ny * nx
        base = stack.data_ptr()
        strides = (C.c_int64 * 5)(*([5 * tile] * 5))
        L.check(lib.spx_cc_fit_stack(
            *(C.c_void_p(base + 4 * tile * s) for s in range(5)), C.byref(strides),
            n, ny, nx, cc_type_code(cc_type), *_out_ptrs(shifts, peak, status, table),
            _ptr(icc), _ptr(ws), ws.numel(), _stream()))
    return shifts, peak, status, icc


_SIZE_CLASSES = (20, 32, 42, 64, 84, 128, 170, 256)     # pick_cfg() in csrc/spx_cc.cu


def cc_fit_ragged(stack, offsets, shapes, cc_type="NCC", want_icc=False):
    """Cross-correlate + fit n pairs of DIFFERENT shapes: pair p is the contiguous
    block ``[5, ny_p, nx_p]`` at element ``offsets[p]`` of the flat float32 CUDA
    tensor ``stack``, ``shapes[p] = (ny_p, nx_p)`` (ragged cutouts of
    create_cutouts, reference cutout.py:159-175; cc.py:21-95 per pair).  Pairs are
    bucketed by size class (20/32/42/64/84/128/170/256 px) - one launch per class.  Returns
    ``(shifts [n,2] f64, peak [n] i32, status [n] i32, icc, icc_offsets)`` with
    ``icc`` a flat float32 tensor holding surface p (``[2ny_p, 2nx_p]``) at
    ``icc_offsets[p]`` (both None unless ``want_icc``).  A pair outside the supported
    range (an axis under 3 or over 256 px - e.g. one extended source among the segment
    boxes of create_cutouts) does not abort the batch: it gets a NaN shift, peak -1,
    status ``S_UNSUPPORTED_SIZE`` and an all-NaN surface."""
    require_cuda()
    lib = L.load()
    stack = _f32c(stack, "stack").reshape(-1)
    dev = stack.device
    off = torch.as_tensor(offsets).to(dtype=torch.int64, device="cpu").reshape(-1)
    shp = torch.as_tensor(shapes).to(dtype=torch.int32, device="cpu").reshape(-1, 2)
    n = int(shp.shape[0])
    if off.numel() != n:
        raise ValueError("offsets and shapes must describe the same number of pairs")
    big = shp.max(dim=1).values if n else shp.new_zeros(0)
    if n and int(shp.min()) < 1:
        raise ValueError("Cutouts must be 2D images.")
    unsupported = ((shp.min(dim=1).values < L.MIN_CUTOUT) | (big > L.MAX_CUTOUT)) if n \
        else torch.zeros(0, dtype=torch.bool)
    if n and (int(off.min()) < 0 or
              int((off + 5 * shp[:, 0].long() * shp[:, 1].long()).max()) > stack.numel()):
        raise ValueError("pairs exceed the stack buffer")
    with torch.cuda.device(dev):
        shifts = torch.empty((n, 2), dtype=torch.float64, device=dev)
        peak = torch.empty((n,), dtype=torch.int32, device=dev)
        status = torch.empty((n,), dtype=torch.int32, device=dev)
        icc = icc_off = d_icc_off = None
        if want_icc:
            icc_off = torch.zeros(n + 1, dtype=torch.int64)
            torch.cumsum(4 * shp[:, 0].long() * shp[:, 1].long(), 0, out=icc_off[1:])
            icc = torch.empty(int(icc_off[-1]), dtype=torch.float32, device=dev)
            d_icc_off = icc_off[:-1].contiguous().to(dev)
        if n == 0:
            return shifts, peak, status, icc, icc_off
        d_off, d_shp = off.to(dev), shp.contiguous().to(dev)
        if bool(unsupported.any()):
            bad = torch.nonzero(unsupported).reshape(-1)
            d_bad = bad.to(dev)
            shifts[d_bad] = float("nan")
            peak[d_bad] = -1
            status[d_bad] = L.S_UNSUPPORTED_SIZE
            if want_icc:
                for k in bad.tolist():
                    icc[int(icc_off[k]):int(icc_off[k + 1])] = float("nan")
        lo = 0
        for cls in _SIZE_CLASSES:
            idx = torch.nonzero((big > lo) & (big <= cls) & ~unsupported).reshape(-1)
            lo = cls
            m = int(idx.numel())
            if m == 0:
                continue
            my, mx = int(shp[idx, 0].max()), int(shp[idx, 1].max())
            nb = lib.spx_cc_workspace_bytes(m, my, mx, int(want_icc))
            ws = _workspace(nb, dev)
            order = idx.to(torch.int32).to(dev)
            L.check(lib.spx_cc_fit_ragged(
                _ptr(stack), _ptr(d_off), _ptr(d_shp), _ptr(order), m, my, mx,
                cc_type_code(cc_type), _ptr(shifts), _ptr(peak), _ptr(status),
                _ptr(icc), _ptr(d_icc_off), _ptr(ws), ws.numel(), _stream()))
    return shifts, peak, status, icc, icc_off


def cc_fit_frames(frames, boxes, ny, nx, cc_type="NCC", want_icc=False,
                  zero_bad_phases=True, zero_bad_ref=False, table=None):
    """Fused gather + cross-correlation + fit straight from five full frames
    (``frames[0]`` = image frame, ``frames[1..4]`` = reference frames at the
    four half-pixel phases; a ``[5, H, W]`` tensor or a sequence of five
    ``[H, W]`` tensors) and an int32 ``[n, 4]`` box table (x1, y1, x2, y2),
    trc inclusive, every box nx wide and ny high."""
    require_cuda()
    lib = L.load()
    fr = [_f32c(f, "frame") for f in frames]
    if len(fr) != 5 or any(f.dim() != 2 or f.shape != fr[0].shape for f in fr):
        raise ValueError("need five [H, W] frames of identical shape")
    H, W = (int(v) for v in fr[0].shape)
    dev = fr[0].device
    boxes = boxes.to(device=dev, dtype=torch.int32).contiguous()
    if boxes.dim() != 2 or boxes.shape[1] != 4:
        raise ValueError("boxes must have shape [n, 4]")
    n = int(boxes.shape[0])
    flags = (L.F_ZERO_BAD_PHASES if zero_bad_phases else 0) | \
            (L.F_ZERO_BAD_REF if zero_bad_ref else 0)
    with torch.cuda.device(dev):
        shifts, peak, status, icc = _outputs(n, ny, nx, want_icc, dev, table)
        if n == 0:
            return shifts, peak, status, icc
        nb = lib.spx_cc_workspace_bytes(n, ny, nx, int(want_icc))
        ws = _workspace(nb, dev)
        ptrs = (C.c_void_p * 5)(*(f.data_ptr() for f in fr))
        L.check(lib.spx_cc_fit_frames(
            C.byref(ptrs), H, W, W, _ptr(boxes), n, int(ny), int(nx),
            cc_type_code(cc_type), flags, *_out_ptrs(shifts, peak, status, table),
            _ptr(icc), _ptr(ws), ws.numel(), _stream()))
    return shifts, peak, status, icc


def gather_cutouts(frame, boxes, dq=None, weight=None, seg=None, src_ids=None,
                   dq_good_bits=None, fillval=0.0, zero_masked=False):
    """Batched Cutout crop (reference cutout.py:737-779 and the mask updates at
    cutout.py:188-196, 286-292, 416-427).  ``boxes`` is an int ``[n, 4]`` table
    (x1, y1, x2, y2), trc inclusive; sizes may differ (ragged).  Returns a dict
    with flat ``data`` / ``mask`` (/ ``dq`` / ``weight``) tensors, the int64
    ``offsets`` ``[n+1]`` and ``shapes`` ``[n, 2]`` (h, w) (CPU tensors)."""
    require_cuda()
    lib = L.load()
    if not frame.is_cuda or frame.dim() != 2:
        raise ValueError("frame must be a 2-D CUDA tensor")
    if frame.dtype not in (torch.float32, torch.float64):
        raise TypeError("frame dtype %s is not supported (float32/float64)" % frame.dtype)
    if frame.stride(1) != 1:
        frame = frame.contiguous()
    dev = frame.device
    H, W = (int(v) for v in frame.shape)
    bx = torch.as_tensor(boxes).to(dtype=torch.int64, device="cpu").reshape(-1, 4)
    n = int(bx.shape[0])
    hw = torch.stack([bx[:, 3] - bx[:, 1] + 1, bx[:, 2] - bx[:, 0] + 1], dim=1)
    if n and int(hw.min()) < 1:
        raise ValueError("Ill-formed extraction box")
    sizes = hw[:, 0] * hw[:, 1]
    offsets = torch.zeros(n + 1, dtype=torch.int64)
    torch.cumsum(sizes, 0, out=offsets[1:])
    total = int(offsets[-1])

    def aux(t, name, kinds):
        if t is None:
            return None
        if not t.is_cuda or tuple(t.shape) != (H, W):
            raise ValueError("Image's %s array shape must match the shape of image 'data'." % name)
        if t.dtype not in kinds:
            raise TypeError("%s dtype %s is not supported" % (name, t.dtype))
        return t.contiguous()

    ints = (torch.int8, torch.uint8, torch.int16, torch.int32, torch.int64)
    dq = aux(dq, "DQ", ints)
    weight = aux(weight, "weight", (torch.float32, torch.float64))
    seg = aux(seg, "segmentation", (torch.int32, torch.int64))
    with torch.cuda.device(dev):
        out = {
            "data": torch.empty(total, dtype=frame.dtype, device=dev),
            "mask": torch.empty(total, dtype=torch.uint8, device=dev),
            "dq": torch.empty(total, dtype=dq.dtype, device=dev) if dq is not None else None,
            "weight": (torch.empty(total, dtype=weight.dtype, device=dev)
                       if weight is not None else None),
            "offsets": offsets, "shapes": hw,
        }
        if n == 0:
            return out
        d_boxes = bx.to(torch.int32).to(dev)
        d_off = offsets[:-1].contiguous().to(dev)
        d_ids = (torch.as_tensor(src_ids).to(dtype=torch.int64).reshape(-1).to(dev)
                 if (seg is not None and src_ids is not None) else None)
        if d_ids is not None and d_ids.numel() != n:
            raise ValueError("src_ids must have one entry per box")
        a = L.GatherArgs()
        a.frame, a.frame_kind = frame.data_ptr(), _KIND[frame.dtype]
        a.frame_h, a.frame_w, a.frame_pitch = H, W, int(frame.stride(0))
        if dq is not None:
            a.dq, a.dq_kind = dq.data_ptr(), _KIND[dq.dtype]
            a.dq_out = out["dq"].data_ptr()
        if weight is not None:
            a.weight, a.weight_kind = weight.data_ptr(), _KIND[weight.dtype]
            a.weight_out = out["weight"].data_ptr()
        if seg is not None:
            a.seg, a.seg_kind = seg.data_ptr(), _KIND[seg.dtype]
        if d_ids is not None:
            a.src_ids = d_ids.data_ptr()
        a.dq_use_bits = int(dq is not None and dq_good_bits is not None)
        a.dq_good_bits = int(dq_good_bits) if dq_good_bits is not None else 0
        a.boxes, a.out_offset, a.n = d_boxes.data_ptr(), d_off.data_ptr(), n
        a.max_h, a.max_w = int(hw[:, 0].max()), int(hw[:, 1].max())
        a.fillval, a.zero_masked = float(fillval), int(bool(zero_masked))
        a.data_out, a.mask_out = out["data"].data_ptr(), out["mask"].data_ptr()
        L.check(lib.spx_gather_cutouts(C.byref(a), _stream()))
        # keep the argument tensors alive until the stream has consumed them
        out["_keep"] = (d_boxes, d_off, d_ids)
    return out


_BLOT_INTERP = {"nearest": L.BLOT_NEAREST, "linear": L.BLOT_LINEAR, "poly3": L.BLOT_POLY3,
                "poly5": L.BLOT_POLY5}

# align.py:664-679: blot to the image cutout, then with dx -= 0.5, dy -= 0.5 (both),
# dy -= 0.5 alone.  Cutout.dx -= 0.5 moves the cutout grid by +0.5 px in the image
# frame (cutout.py:1112-1145), so pixel i of the displaced blot samples the map of
# the undisplaced cutout at i + 0.5.  Order = find_displacement's argument order.
PHASES_00_10_01_11 = ((0.0, 0.0), (0.5, 0.0), (0.0, 0.5), (0.5, 0.5))


def blot_phases(src, src_offsets, src_shapes, maps, tgt_shapes, phases=PHASES_00_10_01_11,
                scale=None, interp="poly5", misval=0.0, prepare=True, out=None,
                out_offsets=None):
    """Batched blot (reference blot.py:79-155 through cdriz.tblot; the four
    half-pixel phases of align.py:661-679 in one launch).

    ``src``: flat float32 CUDA tensor with the (masked-pixels-zeroed) source
    cutouts, cutout k at ``src_offsets[k]`` with shape ``src_shapes[k] = (h, w)``;
    ``maps``: ``[n, 20]`` float64 biquadratic target->source pixel maps (see
    ``include/subpixal_b200.h``); ``tgt_shapes``: ``[n, 2]`` (ny, nx).
    Returns ``(out, out_offsets)``: flat float32 tensor holding ``[nph, ny, nx]``
    per cutout and the int64 ``[n+1]`` offsets (CPU).  ``prepare``: (poly5) tabulate
    the sources' x-direction central differences once in a scratch buffer shared by
    all phases (16 B per source pixel; bit-identical results, about twice as fast).
    ``out`` / ``out_offsets``: write into an existing flat float32 CUDA tensor, phase 0
    of cutout k at element ``out_offsets[k]`` (e.g. slots 1..4 of a ragged
    ``cc_fit_ragged`` stack) instead of allocating."""
    require_cuda()
    lib = L.load()
    if interp not in _BLOT_INTERP:
        raise ValueError("interp %r is not supported on the GPU (nearest, linear, poly3, poly5)"
                         % (interp,))
    src = _f32c(src, "src").reshape(-1)
    dev = src.device
    cpu = lambda t, dt: torch.as_tensor(t).to(dtype=dt, device="cpu").contiguous()
    s_off = cpu(src_offsets, torch.int64).reshape(-1)
    s_shp = cpu(src_shapes, torch.int32).reshape(-1, 2)
    t_shp = cpu(tgt_shapes, torch.int32).reshape(-1, 2)
    mp = cpu(maps, torch.float64).reshape(-1, 20)
    n = int(t_shp.shape[0])
    if not (s_shp.shape[0] == n and mp.shape[0] == n and s_off.numel() >= n):
        raise ValueError("src_offsets, src_shapes, maps and tgt_shapes must describe n cutouts")
    s_off = s_off[:n].contiguous()
    ph = [(float(a), float(b)) for a, b in phases]
    nph = len(ph)
    if not 1 <= nph <= L.BLOT_MAX_PHASES:
        raise ValueError("between 1 and %d phases" % L.BLOT_MAX_PHASES)
    if n and (int(s_shp.min()) < 1 or int(t_shp.min()) < 1):
        raise ValueError("Cutouts must be 2D images.")
    if n and (int(s_off.min()) < 0 or
              int((s_off + s_shp[:, 0].long() * s_shp[:, 1].long()).max()) > src.numel()):
        raise ValueError("source cutouts exceed the src buffer")
    sizes = t_shp[:, 0].long() * t_shp[:, 1].long() * nph
    if out is None:
        o_off = torch.zeros(n + 1, dtype=torch.int64)
        torch.cumsum(sizes, 0, out=o_off[1:])
        o_start = o_off[:-1].contiguous()
    else:
        out = _f32c(out, "out")
        if out.device != dev or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 tensor on the device of src")
        o_off = o_start = cpu(out_offsets, torch.int64).reshape(-1)
        if o_start.numel() != n or (n and (int(o_start.min()) < 0 or
                                           int((o_start + sizes).max()) > out.numel())):
            raise ValueError("out_offsets must place every blot inside out")
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty(int(o_off[-1]), dtype=torch.float32, device=dev)
        if n == 0:
            return out, o_off
        keep = [s_off.to(dev), s_shp.to(dev), mp.to(dev), t_shp.to(dev), o_start.to(dev)]
        d_scale = None
        if scale is not None:
            d_scale = cpu(scale, torch.float32).reshape(-1)
            if d_scale.numel() != n:
                raise ValueError("scale must have one entry per cutout")
            d_scale = d_scale.to(dev)
        a = L.BlotArgs()
        a.src = src.data_ptr()
        a.src_offset, a.src_shape = keep[0].data_ptr(), keep[1].data_ptr()
        a.map, a.tgt_shape, a.out_offset = keep[2].data_ptr(), keep[3].data_ptr(), keep[4].data_ptr()
        a.scale = d_scale.data_ptr() if d_scale is not None else None
        a.out = out.data_ptr()
        flat = (C.c_double * (2 * nph))(*[v for p2 in ph for v in p2])
        a.phases_host = C.cast(flat, C.POINTER(C.c_double))
        a.n, a.nph, a.interp = n, nph, _BLOT_INTERP[interp]
        a.max_ny, a.max_nx = int(t_shp[:, 0].max()), int(t_shp[:, 1].max())
        a.misval = float(misval)
        if prepare and interp == "poly5":
            # spx_blot_workspace_bytes(h, w) / 16 = (h + 5) * (w + 1) float4 per source
            units = (s_shp[:, 0].long() + 5) * (s_shp[:, 1].long() + 1)
            w_off = torch.zeros(n + 1, dtype=torch.int64)
            torch.cumsum(units, 0, out=w_off[1:])
            assert int(units[0]) * 16 == lib.spx_blot_workspace_bytes(int(s_shp[0, 0]), int(s_shp[0, 1]))
            ws = torch.empty(int(w_off[-1]) * 4, dtype=torch.float32, device=dev)
            keep += [ws, w_off[:-1].contiguous().to(dev)]
            a.workspace, a.ws_offset = ws.data_ptr(), keep[-1].data_ptr()
            a.max_src_h, a.max_src_w = int(s_shp[:, 0].max()), int(s_shp[:, 1].max())
        # the argument tables are freed in stream order by the caching allocator
        L.check(lib.spx_blot_phases(C.byref(a), _stream()))
    return out, o_off


def gather_into(frame, boxes, out, out_offsets, fillval=0.0, zero_masked=False, seg=None,
                src_ids=None, dq=None, dq_good_bits=None):
    """Crop ``boxes`` ([n,4] x1,y1,x2,y2, trc inclusive) out of the 2-D float32 CUDA
    ``frame`` straight into the flat float32 tensor ``out``, cutout k at element
    ``out_offsets[k]`` (``spx_gather_cutouts`` without a mask output): used to place
    image cutouts in slot 0 of a ragged ``cc_fit_ragged`` stack and to build the
    source stack of ``blot_phases`` without a host round trip.  Masked pixels
    (outside the frame, non-finite, ``seg != src_id``, bad DQ bits) are zeroed when
    ``zero_masked`` (align.py:661)."""
    require_cuda()
    lib = L.load()
    frame = _f32c(frame, "frame")
    out = _f32c(out, "out")
    if frame.dim() != 2 or not out.is_contiguous():
        raise ValueError("frame must be 2-D and out contiguous")
    dev = frame.device
    H, W = (int(v) for v in frame.shape)
    bx = torch.as_tensor(boxes).to(dtype=torch.int64, device="cpu").reshape(-1, 4)
    n = int(bx.shape[0])
    if n == 0:
        return
    hw = torch.stack([bx[:, 3] - bx[:, 1] + 1, bx[:, 2] - bx[:, 0] + 1], dim=1)
    off = torch.as_tensor(out_offsets).to(dtype=torch.int64, device="cpu").reshape(-1)
    if int(hw.min()) < 1:
        raise ValueError("Ill-formed extraction box")
    if off.numel() != n or int(off.min()) < 0 or int((off + hw[:, 0] * hw[:, 1]).max()) > out.numel():
        raise ValueError("out_offsets must place every cutout inside out")
    with torch.cuda.device(dev):
        keep = [bx.to(torch.int32).to(dev), off.to(dev)]
        a = L.GatherArgs()
        a.frame, a.frame_kind = frame.data_ptr(), L.T_F32
        a.frame_h, a.frame_w, a.frame_pitch = H, W, int(frame.stride(0))
        if seg is not None and src_ids is not None:
            if tuple(seg.shape) != (H, W) or seg.dtype not in (torch.int32, torch.int64):
                raise ValueError("segmentation image must be int32/int64 with the frame's shape")
            seg = seg.contiguous()
            keep.append(torch.as_tensor(src_ids).to(dtype=torch.int64).reshape(-1).to(dev))
            if keep[-1].numel() != n:
                raise ValueError("src_ids must have one entry per box")
            a.seg, a.seg_kind, a.src_ids = seg.data_ptr(), _KIND[seg.dtype], keep[-1].data_ptr()
        if dq is not None:
            if tuple(dq.shape) != (H, W) or dq.dtype not in _KIND or dq.is_floating_point():
                raise ValueError("DQ image must be an integer image with the frame's shape")
            dq = dq.contiguous()
            a.dq, a.dq_kind = dq.data_ptr(), _KIND[dq.dtype]
            a.dq_use_bits = int(dq_good_bits is not None)
            a.dq_good_bits = int(dq_good_bits) if dq_good_bits is not None else 0
        a.boxes, a.out_offset, a.n = keep[0].data_ptr(), keep[1].data_ptr(), n
        a.max_h, a.max_w = int(hw[:, 0].max()), int(hw[:, 1].max())
        a.fillval, a.zero_masked = float(fillval), int(bool(zero_masked))
        a.data_out = out.data_ptr()
        L.check(lib.spx_gather_cutouts(C.byref(a), _stream()))


def find_peak_batch(surfaces, mask=None, guess=None, search_box=(0, 0),
                    fit_box=(5, 5)):
    """Batched centroid.find_peak (reference centroid.py:18-236) on a
    ``[n, h, w]`` float32/float64 CUDA tensor.  ``mask`` (uint8/bool, 1 = good),
    ``guess`` ``[n, 2]`` float64 (xmax, ymax) or None.  Returns
    ``(xy [n,2] f64, peak [n] i32, status [n] i32)``."""
    require_cuda()
    lib = L.load()
    if not surfaces.is_cuda or surfaces.dim() != 3:
        raise ValueError("surfaces must be a [n, h, w] CUDA tensor")
    if surfaces.dtype not in (torch.float32, torch.float64):
        raise TypeError("surfaces must be float32 or float64")
    surfaces = surfaces.contiguous()
    n, h, w = (int(v) for v in surfaces.shape)
    dev = surfaces.device
    if mask is not None:
        mask = mask.to(device=dev, dtype=torch.uint8).contiguous()
        if tuple(mask.shape) != (n, h, w):
            raise ValueError("mask shape must match surfaces")
    if guess is not None:
        guess = guess.to(device=dev, dtype=torch.float64).contiguous()
        if tuple(guess.shape) != (n, 2):
            raise ValueError("guess must have shape [n, 2]")
    with torch.cuda.device(dev):
        xy = torch.empty((n, 2), dtype=torch.float64, device=dev)
        peak = torch.empty((n,), dtype=torch.int32, device=dev)
        status = torch.empty((n,), dtype=torch.int32, device=dev)
        if n:
            L.check(lib.spx_find_peak(
                _ptr(surfaces), _KIND[surfaces.dtype], n, h, w, _ptr(mask), _ptr(guess),
                int(search_box[0]), int(search_box[1]), int(fit_box[0]), int(fit_box[1]),
                _ptr(xy), _ptr(peak), _ptr(status), _stream()))
    return xy, peak, status


MAX_SEG_LABEL = 1 << 24


def segment_boxes(seg):
    """Bounding box of every label of a 2-D int32/int64 CUDA segmentation image in
    one pass (replaces the per-source scans of reference cutout.py:151-162).
    Returns an int32 ``[max_label+1, 4]`` CUDA tensor (x1, x2, y1, y2); absent
    labels have x2 = -1."""
    require_cuda()
    lib = L.load()
    if not seg.is_cuda or seg.dim() != 2 or seg.dtype not in (torch.int32, torch.int64):
        raise ValueError("seg must be a 2-D int32/int64 CUDA tensor")
    seg = seg.contiguous()
    h, w = (int(v) for v in seg.shape)
    max_label = int(seg.max()) if seg.numel() else 0
    if max_label > MAX_SEG_LABEL:
        raise ValueError("segmentation labels above %d are not supported" % MAX_SEG_LABEL)
    max_label = max(max_label, 0)
    with torch.cuda.device(seg.device):
        out = torch.empty((max_label + 1, 4), dtype=torch.int32, device=seg.device)
        L.check(lib.spx_segment_boxes(_ptr(seg), _KIND[seg.dtype], h, w, w, max_label,
                                      _ptr(out), _stream()))
    return out
