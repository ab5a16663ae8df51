"""Host-side mirror of the reference's `Pedestrians` feature helpers (src/data/data.py:343-535) on the CUDA path.

Same method names, argument order and return shapes as the reference class, so `BaseSimulator`/`TimeIndexedPedData`
can inherit from (or be patched with) this class unchanged.  All arithmetic happens in libpiml_b200.so.
"""
import math

import torch

from . import _lib as L


def cos_threshold(angle_threshold):
    """data.py:442-443: `math.cos(3.14 * angle / 180)` (3.14, not pi), compared in fp32."""
    return float(torch.tensor(math.cos(3.14 * angle_threshold / 180), dtype=torch.float32))


class Pedestrians(object):
    """Drop-in for reference `data.data.Pedestrians` (data.py:343)."""

    def __init__(self):
        super(Pedestrians, self).__init__()

    # ---- data.py:351-395 -------------------------------------------------------------------------------------
    @staticmethod
    def get_heading_direction(velocity):
        """velocity (*c, t, N, 2) -> heading_direction, same shape (zero-speed frames filled, then normalised)."""
        if velocity.dim() not in (3, 4):
            raise ValueError("get_heading_direction expects (t,N,2) or (c,t,N,2)")
        _, origin, (velocity,) = L.stage(velocity)
        v = L.f32c(velocity)
        Cc = v.shape[0] if v.dim() == 4 else 1
        T, N = v.shape[-3], v.shape[-2]
        out = torch.empty_like(v)
        L.check(L.load().piml_heading_f32(L.ptr(v), Cc, T, N, L.ptr(out), L.stream_ptr(v.device)),
                "piml_heading_f32")
        return out.to(origin)

    # ---- data.py:398-414 -------------------------------------------------------------------------------------
    @staticmethod
    def get_relative_quantity(A, B):
        """relative_A[..., n, m, :] = B[..., m, :] - A[..., n, :] as a dense (..., N, M, dim) tensor.  Kept for the
        reference's polar / symbolic-regression callers (out of the hot path); the fused kernels never materialise it."""
        _, origin, (A, B) = L.stage(A, B)
        a, b = L.f32c(A), L.f32c(B)
        if a.shape[:-2] != b.shape[:-2] or a.shape[-1] != b.shape[-1]:
            raise ValueError("get_relative_quantity: A (..., N, dim) and B (..., M, dim) must share the leading dims")
        lead, N, Mo, d = a.shape[:-2], a.shape[-2], b.shape[-2], a.shape[-1]
        out = torch.empty(*lead, N, Mo, d, dtype=torch.float32, device=a.device)
        frames = 1
        for s_ in lead:
            frames *= s_
        L.check(L.load().piml_relative_quantity_f32(L.ptr(a), L.ptr(b), frames, N, Mo, d, L.ptr(out),
                                                    L.stream_ptr(a.device)), "piml_relative_quantity_f32")
        return out.to(origin)

    # ---- data.py:416-447 -------------------------------------------------------------------------------------
    def get_nearby_obj_in_sight(self, position, objects, heading_direction, k, angle_threshold):
        """Returns (sorted_dist[..., :k], indices[..., :k]); objects outside the field of view carry inf."""
        _, origin, (position, objects, heading_direction) = L.stage(position, objects, heading_direction)
        pos, obj, head = L.f32c(position), L.f32c(objects), L.f32c(heading_direction)
        lead = pos.shape[:-2]
        N, M = pos.shape[-2], obj.shape[-2]
        B = 1
        for s in lead:
            B *= s
        if obj.shape[:-2] != lead:
            obj = obj.expand(*lead, M, 2).contiguous()
        kk = min(k, M)
        dist = torch.empty(*lead, N, kk, dtype=torch.float32, device=pos.device)
        idx = torch.empty(*lead, N, kk, dtype=torch.int64, device=pos.device)
        lib = L.load()
        done = 0
        while done < B:                       # grid.y limit: 65535 frames per launch
            nb = min(B - done, 65535)
            L.check(lib.piml_select_neighbors_f32(
                L.C.c_void_p(pos.data_ptr() + done * N * 8), L.C.c_void_p(obj.data_ptr() + done * M * 8), M * 2,
                L.C.c_void_p(head.data_ptr() + done * N * 8), nb, N, M, k, cos_threshold(angle_threshold),
                L.C.c_void_p(dist.data_ptr() + done * N * kk * 4), L.C.c_void_p(idx.data_ptr() + done * N * kk * 8),
                L.stream_ptr(pos.device)), "piml_select_neighbors_f32")
            done += nb
        return dist.to(origin), idx.to(origin)

    # ---- data.py:449-464 -------------------------------------------------------------------------------------
    def get_filtered_features(self, features, nearby_idx, nearby_dist, dist_threshold):
        """gather the k selected columns of (..., N, M, dim) and zero slots farther than dist_threshold.
        Only used by out-of-scope reference callers; the hot path uses the fused get_relative_features."""
        _, origin, (features, nearby_idx, nearby_dist) = L.stage(features, nearby_idx, nearby_dist)
        f, dist = L.f32c(features), L.f32c(nearby_dist)
        idx = nearby_idx.to(torch.int64).contiguous()
        Mo, d, k = f.shape[-2], f.shape[-1], idx.shape[-1]
        rows = idx.numel() // max(k, 1)
        out = torch.empty(*idx.shape, d, dtype=torch.float32, device=f.device)
        L.check(L.load().piml_filtered_features_f32(L.ptr(f), L.ptr(idx), L.ptr(dist), rows, Mo, k, d,
                                                    float(dist_threshold), L.ptr(out), L.stream_ptr(f.device)),
                "piml_filtered_features_f32")
        return out.to(origin)

    # ---- data.py:466-512 -------------------------------------------------------------------------------------
    def get_relative_features(self, position, velocity, acceleration, destination, obstacles, topk_ped,
                              sight_angle_ped, dist_threshold_ped, topk_obs, sight_angle_obs, dist_threshold_obs,
                              return_selection=False):
        """position/velocity/acceleration/destination (*c, t, N, 2); obstacles (M, 2) or (c, M, 2).
        Returns ped_features (*c,t,N,k1,6), obs_features (*c,t,N,k2,6), dest_features (*c,t,N,2).
        Like the reference it zeroes NaNs IN PLACE in the caller's velocity and acceleration tensors.
        When gradients are being recorded for position / velocity / acceleration / destination (the differentiable
        rollout, simulators.py:772-776) the call goes through autograd.RelativeFeaturesFunction, whose backward is
        the CUDA scatter kernel; velocity / acceleration must then be NaN-free (they are: simulators.py:745)."""
        if torch.is_grad_enabled() and not return_selection and any(
                t.requires_grad for t in (position, velocity, acceleration, destination)):
            from .autograd import RelativeFeaturesFunction
            if not position.is_cuda:
                raise RuntimeError("piml_b200: the differentiable feature path needs CUDA tensors; no CPU fallback")
            return RelativeFeaturesFunction.apply(self, position, velocity, acceleration, destination, obstacles,
                                                  topk_ped, sight_angle_ped, dist_threshold_ped, topk_obs,
                                                  sight_angle_obs, dist_threshold_obs)
        return self._relative_features_raw(position, velocity, acceleration, destination, obstacles, topk_ped,
                                           sight_angle_ped, dist_threshold_ped, topk_obs, sight_angle_obs,
                                           dist_threshold_obs, return_selection)

    def _relative_features_raw(self, position, velocity, acceleration, destination, obstacles, topk_ped,
                               sight_angle_ped, dist_threshold_ped, topk_obs, sight_angle_obs, dist_threshold_obs,
                               return_selection=False):
        if position.dim() not in (3, 4):
            raise ValueError("get_relative_features expects (t,N,2) or (c,t,N,2) inputs")
        _, origin, (pos, vel, acc, dest, obs) = L.stage(position, velocity, acceleration, destination, obstacles)
        pos, dest = L.f32c(pos), L.f32c(dest)
        vel, acc = L.f32c(vel), L.f32c(acc)
        obs = L.f32c(obs)
        lead = pos.shape[:-2]
        Cc = pos.shape[0] if pos.dim() == 4 else 1
        T, N = pos.shape[-3], pos.shape[-2]
        M = obs.shape[-2] if obs.numel() > 0 else 0
        per_channel = 1 if (obs.dim() == 3 and M > 0) else 0
        if per_channel and obs.shape[0] != Cc:
            raise ValueError("per-channel obstacles must have the same channel count as position")
        kp, ko = min(topk_ped, N), (min(topk_obs, M) if M else 0)
        dev = pos.device
        ped_f = torch.empty(*lead, N, kp, 6, dtype=torch.float32, device=dev)
        obs_f = torch.empty(*lead, N, ko, 6, dtype=torch.float32, device=dev) if M else \
            torch.tensor([[] for _ in range(T)], device=dev)                   # data.py:499
        dest_f = torch.empty(*lead, N, 2, dtype=torch.float32, device=dev)
        sel = None
        if return_selection:
            sel = (torch.empty(*lead, N, kp, dtype=torch.int64, device=dev),
                   torch.empty(*lead, N, kp, dtype=torch.float32, device=dev),
                   torch.empty(*lead, N, ko, dtype=torch.int64, device=dev),
                   torch.empty(*lead, N, ko, dtype=torch.float32, device=dev))
        lib = L.load()
        head = None
        if T > 1:
            # heading needs the fill over time (data.py:362-389); the kernel sanitises NaN velocities on read
            head = torch.empty_like(vel)
            L.check(lib.piml_heading_f32(L.ptr(vel), Cc, T, N, L.ptr(head), L.stream_ptr(dev)), "piml_heading_f32")
        if Cc * T > 65535:
            raise ValueError("more than 65535 frames in one get_relative_features call; split the batch")
        selp = [L.ptr(s) for s in sel] if sel else [None] * 4
        L.check(lib.piml_relative_features_f32(
            L.ptr(pos), L.ptr(vel), L.ptr(acc), L.ptr(dest), L.ptr(head), L.ptr(obs) if M else None, per_channel,
            Cc, T, N, M, topk_ped, cos_threshold(sight_angle_ped), float(dist_threshold_ped), topk_obs,
            cos_threshold(sight_angle_obs), float(dist_threshold_obs), L.ptr(ped_f), L.ptr(obs_f) if M else None,
            L.ptr(dest_f), *selp, L.stream_ptr(dev)), "piml_relative_features_f32")
        # in-place NaN->0 side effect on the caller's tensors (data.py:483-484) when we had to copy them
        if vel.data_ptr() != velocity.data_ptr():
            velocity.detach().copy_(vel)
        if acc.data_ptr() != acceleration.data_ptr():
            acceleration.detach().copy_(acc)
        if origin != dev:
            ped_f, obs_f, dest_f = ped_f.to(origin), obs_f.to(origin), dest_f.to(origin)
            sel = tuple(s_.to(origin) for s_ in sel) if sel else None
        if return_selection:
            return ped_f, obs_f, dest_f, sel
        return ped_f, obs_f, dest_f

    # ---- data.py:515-535 -------------------------------------------------------------------------------------
    @staticmethod
    def calculate_collision_label(ped_features):
        """ped_features (..., k, 6) -> collisions (..., k) in {0, 1}."""
        _, origin, (ped_features,) = L.stage(ped_features)
        f = L.f32c(ped_features)
        out = torch.empty(f.shape[:-1], dtype=torch.float32, device=f.device)
        L.check(L.load().piml_collision_label_f32(L.ptr(f), out.numel(), L.ptr(out), L.stream_ptr(f.device)),
                "piml_collision_label_f32")
        return out.to(origin)

    # ---- data.py:538-601 -------------------------------------------------------------------------------------
    @staticmethod
    def collision_detection(position, threshold, real_position=None, rowsum_only=False):
        """position (t,N,2) or (c,t,N,2) with NaN for absent pedestrians -> collisions (...,N,N) in {0,1}: pairs
        closer than `threshold`, self pairs and "friends" removed (3-d input: pairs touching in more than 25 frames of
        position / real_position; 4-d input: pairs touching within the first 4 frames of a channel).
        rowsum_only=True returns collisions.sum(-1) without materialising the N x N tensor -- the only use the
        training rollout makes of it (simulators.py:707-724)."""
        if position.dim() not in (3, 4):
            raise ValueError("collision_detection expects (t,N,2) or (c,t,N,2)")
        if real_position is not None and (real_position.dim() != 3 or position.dim() != 3):
            raise AssertionError('Value Error: real_position only supports 3 dimensional inputs (t,N,2)')
        _, origin, (position, real_position) = L.stage(position, real_position)
        pos = L.f32c(position.detach())
        real = L.f32c(real_position.detach()) if real_position is not None else None
        if real is not None and real.shape[-2] != pos.shape[-2]:
            raise ValueError("real_position must have the same number of pedestrians as position")
        mode = pos.dim()
        Cc = pos.shape[0] if mode == 4 else 1
        T, N = pos.shape[-3], pos.shape[-2]
        if real is not None and real.shape[0] != T:
            raise NotImplementedError("real_position with a different number of frames")
        dev = pos.device
        full = None if rowsum_only else torch.empty(*pos.shape[:-1], N, device=dev)
        rows = torch.empty(*pos.shape[:-1], device=dev) if rowsum_only else None
        L.check(L.load().piml_collision_detection_f32(L.ptr(pos), L.ptr(real), Cc, T, N, float(threshold), mode,
                                                      L.ptr(full), L.ptr(rows), L.stream_ptr(dev)),
                "piml_collision_detection_f32")
        return (rows if rowsum_only else full).to(origin)
