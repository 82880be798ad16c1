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
