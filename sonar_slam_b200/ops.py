"""Batch (device-resident) operator API over libsonarfe's *_dev entry points.

torch is used for what it is good at here -- owning device memory and streams; every
computation is a hand-written CUDA kernel inside libsonarfe.so.  All functions enqueue
on the torch current stream of the tensor's device and return torch tensors.
"""
import ctypes

import torch

from . import _lib

ALG = {"CA": 0, "SOCA": 1, "GOCA": 2, "OS": 3}
_ctx_cache = {}


def context(device=None):
    """libsonarfe context bound to torch's current stream on `device`."""
    dev = torch.cuda.current_device() if device is None else torch.device(device).index or 0
    stream = torch.cuda.current_stream(dev).cuda_stream
    key = (dev, stream)
    ctx = _ctx_cache.get(key)
    if ctx is None:
        ctx = _ctx_cache[key] = _lib.Context(dev, stream)
    return ctx


def _dp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def cfar(img, alg, train_hs, guard_hs, tau, k=0, gate=None, want_mask=True, want_thr=False, want_bits=False,
         ctx=None):
    """Batched CFAR (include/sonarfe.h: sfe_cfar_dev).

    img: cuda tensor [F, R, B] (or [R, B]) uint8 or float32, contiguous.
    Returns dict(mask=uint8[F,R,B] | None, thr=float32[F,R,B] | None, bits=int32[F,R,ceil(B/32)] | None).
    """
    if not img.is_cuda:
        raise _lib.SonarFEError("ops.cfar: image tensor must live on a CUDA device (no CPU path)")
    if img.dim() == 2:
        img = img.unsqueeze(0)
    if img.dtype not in (torch.uint8, torch.float32):
        raise TypeError("ops.cfar: image must be uint8 or float32")
    img = img.contiguous()
    F, R, B = img.shape
    ctx = ctx or context(img.device)
    mask = torch.empty((F, R, B), dtype=torch.uint8, device=img.device) if want_mask else None
    thr = torch.empty((F, R, B), dtype=torch.float32, device=img.device) if want_thr else None
    bits = torch.empty((F, R, (B + 31) // 32), dtype=torch.int32, device=img.device) if want_bits else None
    a = ALG[alg] if isinstance(alg, str) else int(alg)
    _lib.check(ctx.lib.sfe_cfar_dev(ctx.handle, _dp(img), 0 if img.dtype == torch.uint8 else 1, F, R, B, a,
                                    int(train_hs), int(guard_hs), int(k), float(tau),
                                    0 if gate is None else 1, 0.0 if gate is None else float(gate),
                                    _dp(mask), _dp(thr), _dp(bits)), "sfe_cfar_dev")
    return dict(mask=mask, thr=thr, bits=bits)


def cart_points(maps, mask=None, bits=None, capacity=8192, ctx=None):
    """Cartesian feature pixels / points of polar masks (include/sonarfe.h: sfe_cart_points_dev).

    maps: _lib.Maps.  mask: uint8 [F,R,B] cuda tensor, or bits: int32 [F,R,ceil(B/32)] (from cfar()).
    Returns dict(ij=int32[F,capacity,2], xy=float32[F,capacity,2], count=int32[F]).
    """
    src = mask if mask is not None else bits
    if src is None or not src.is_cuda:
        raise _lib.SonarFEError("ops.cart_points: pass a CUDA mask or bits tensor")
    if src.dim() == 2:
        src = src.unsqueeze(0)
    src = src.contiguous()
    F = src.shape[0]
    ctx = ctx or context(src.device)
    ij = torch.empty((F, capacity, 2), dtype=torch.int32, device=src.device)
    xy = torch.empty((F, capacity, 2), dtype=torch.float32, device=src.device)
    count = torch.empty((F,), dtype=torch.int32, device=src.device)
    _lib.check(ctx.lib.sfe_cart_points_dev(ctx.handle, maps.handle, _dp(src) if mask is not None else None,
                                           _dp(src) if mask is None else None, F, int(capacity), _dp(ij), _dp(xy),
                                           _dp(count)), "sfe_cart_points_dev")
    return dict(ij=ij, xy=xy, count=count)


def _offsets(counts, device):
    off = torch.zeros(len(counts) + 1, dtype=torch.int32, device=device)
    off[1:] = torch.cumsum(torch.as_tensor(counts, device=device), 0)
    return off


def downsample(pts, off, n_max, resolution, ctx=None):
    """Batched pcl.downsample on packed clouds (pts [total,2] f32, off [n+1] i32, cuda)."""
    ctx = ctx or context(pts.device)
    n = off.numel() - 1
    out = torch.empty_like(pts)
    idx = torch.empty((pts.shape[0],), dtype=torch.int32, device=pts.device)
    cnt = torch.empty((n,), dtype=torch.int32, device=pts.device)
    _lib.check(ctx.lib.sfe_downsample_dev(ctx.handle, _dp(pts), _dp(off), n, int(pts.shape[1]), int(n_max),
                                          float(resolution), _dp(out), _dp(idx), _dp(cnt)), "sfe_downsample_dev")
    return dict(pts=out, idx=idx, count=cnt)


def remove_outlier(pts, off, n_max, radius, min_points, ctx=None):
    ctx = ctx or context(pts.device)
    n = off.numel() - 1
    out = torch.empty_like(pts)
    idx = torch.empty((pts.shape[0],), dtype=torch.int32, device=pts.device)
    cnt = torch.empty((n,), dtype=torch.int32, device=pts.device)
    _lib.check(ctx.lib.sfe_remove_outlier_dev(ctx.handle, _dp(pts), _dp(off), n, int(pts.shape[1]), int(n_max),
                                              float(radius), int(min_points), _dp(out), _dp(idx), _dp(cnt)),
               "sfe_remove_outlier_dev")
    return dict(pts=out, idx=idx, count=cnt)


def match(ref_pts, ref_off, in_pts, in_off, n_ref_max, max_dist, ctx=None):
    ctx = ctx or context(ref_pts.device)
    n = ref_off.numel() - 1
    ids = torch.empty((in_pts.shape[0],), dtype=torch.int32, device=in_pts.device)
    d = torch.empty((in_pts.shape[0],), dtype=torch.float32, device=in_pts.device)
    _lib.check(ctx.lib.sfe_match_dev(ctx.handle, _dp(ref_pts), _dp(ref_off), _dp(in_pts), _dp(in_off), n,
                                     int(n_ref_max), float(max_dist), _dp(ids), _dp(d)), "sfe_match_dev")
    return dict(ids=ids, dists=d)


def icp(src_pts, src_off, tgt_pts, tgt_off, guess, ns_max, nt_max, params=None, src_id=None, tgt_id=None, ctx=None):
    """Batched ICP (include/sonarfe.h: sfe_icp_dev).  guess: [P,3,3] f32 cuda.  Returns dict of cuda tensors."""
    ctx = ctx or context(src_pts.device)
    prm = params or _lib.IcpParams()
    guess = guess.contiguous()
    P = guess.shape[0]
    T = torch.empty((P, 3, 3), dtype=torch.float32, device=guess.device)
    iters = torch.empty((P,), dtype=torch.int32, device=guess.device)
    inl = torch.empty_like(iters)
    st = torch.empty_like(iters)
    _lib.check(ctx.lib.sfe_icp_dev(ctx.handle, ctypes.byref(prm), _dp(src_pts), _dp(src_off), _dp(tgt_pts),
                                   _dp(tgt_off), _dp(src_id), _dp(tgt_id), P, int(ns_max), int(nt_max), _dp(guess),
                                   _dp(T), _dp(iters), _dp(inl), _dp(st)), "sfe_icp_dev")
    return dict(T=T, iterations=iters, inliers=inl, status=st)
