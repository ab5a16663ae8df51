"""Drop-ins for the evaluation metrics of the reference's `functions/metrics.py` (what `test_multiple_rollouts` calls
per rollout, src/models/simulators.py:508-527) on the CUDA path: one launch for all frames instead of a Python loop
over frames with one Sinkhorn solve / Gram matrix each."""
import numpy as np
import torch

from . import _lib as L
from .features import Pedestrians


MAX_AGENTS = 1024            # PIML_METRICS_MAX_AGENTS (include/piml_b200.h)


def collision_count(position, threshold, real_position=None, reduction=None):
    """metrics.py:16-26."""
    collisions = Pedestrians.collision_detection(position, threshold, real_position)
    if reduction == 'sum':
        return torch.sum(collisions).item()
    if reduction == 'mean':
        return torch.mean(collisions).item()
    if reduction is None:
        return collisions
    raise NotImplementedError


def _frames(p, q, mask, want_ot, want_mmd, eps=0.1, max_iter=100, kernel_mul=2.0, kernel_num=5):
    dev, origin, (p, q, mask) = L.stage(p, q, mask)
    if p.dim() != 3 or p.shape[-1] != 2 or q.shape != p.shape or mask.shape != p.shape[:2]:
        raise NotImplementedError("metrics on the CUDA path take p, q (T,N,2) and mask (T,N)")
    p, q = L.f32c(p), L.f32c(q)
    m8 = (mask == 1).to(torch.uint8).contiguous()
    T, N = p.shape[0], p.shape[1]
    mae = torch.empty(T, device=dev)
    ot = torch.empty(T, device=dev) if want_ot else None
    mmd = torch.empty(T, device=dev) if want_mmd else None
    cnt = torch.empty(T, dtype=torch.int32, device=dev)
    L.check(L.load().piml_metrics_frames_f32(L.ptr(p), L.ptr(q), L.ptr(m8), T, N, float(eps), int(max_iter),
                                             float(kernel_mul), int(kernel_num), L.ptr(mae), L.ptr(ot), L.ptr(mmd),
                                             L.ptr(cnt), L.stream_ptr(dev)), "piml_metrics_frames_f32")
    if (want_ot or want_mmd) and T and int(cnt.max()) > MAX_AGENTS:
        # the per-frame Sinkhorn / Gram kernel keeps a frame's points in shared memory; never return its NaN silently
        raise NotImplementedError(f"ot / mmd with more than {MAX_AGENTS} masked agents in a frame "
                                  f"(got {int(cnt.max())}); the reference has no such limit")
    return mae, ot, mmd, cnt


def _reduce(values, reduction):
    out = [float(x) for x in values]                  # the reference collects python floats and reduces with numpy
    if reduction == 'sum':
        return np.sum(out)
    if reduction == 'mean':
        return np.mean(out)
    return out


def mae_with_time_mask(p, q, mask, reduction=None):
    """metrics.py:29-42 (p, q (T,N,2), mask (T,N))."""
    mae, _, _, cnt = _frames(p, q, mask, False, False)
    total = float(mae.double().sum())
    if reduction == 'sum':
        return total
    if reduction == 'mean':
        return total / max(int(cnt.sum()), 1)
    raise NotImplementedError


def ot_with_time_mask(p, q, mask, eps=0.1, max_iter=100, reduction=None, dvs='cpu'):
    """metrics.py:45-67: one Sinkhorn distance per frame with more than one masked agent."""
    _, ot, _, cnt = _frames(p, q, mask, True, False, eps=eps, max_iter=max_iter)
    return _reduce(ot[cnt > 1].cpu().tolist(), reduction)


def mmd_with_time_mask(p, q, mask, kernel_mul=2.0, kernel_num=5, fix_sigma=None, reduction=None):
    """metrics.py:70-91."""
    if fix_sigma:
        raise NotImplementedError("fix_sigma")
    _, _, mmd, cnt = _frames(p, q, mask, False, True, kernel_mul=kernel_mul, kernel_num=kernel_num)
    return _reduce(mmd[cnt > 1].cpu().tolist(), reduction)
