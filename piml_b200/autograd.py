"""torch.autograd bindings of the CUDA forward/backward kernels used by the training paths.

The reference trains through eager autograd: `loss.backward()` (src/models/simulators.py:359) walks the graph built
by `model(*state_features)` (:330, :701), the Euler chain (:741-743), the entry overwrite (:762-769) and
`get_relative_features` (:772-776 -> src/data/data.py:466-512).  Here each of those stages is one
`torch.autograd.Function` whose forward AND backward are kernels of libpiml_b200.so; torch only routes the gradients.
"""
import torch

from . import _lib as L


def _floats(fn_name, desc, has_obs, R, kp, ko):
    n = int(getattr(L.load(), fn_name)(L.C.byref(desc), 1 if has_obs else 0, R, kp, ko))
    if n < 0:
        raise RuntimeError(f"{fn_name} failed: {L.last_error()}")
    return n


class PinnsfFunction(torch.autograd.Function):
    """PINNSF-family forward with activation stash (piml_pinnsf_forward_train_f32) and its backward
    (piml_pinnsf_backward_f32).  `params` are the module's Linear weights/biases in `models.linear_keys` order; they
    only route gradients -- the kernels read the packed copies `packed_fwd` / `packed_bwd`."""

    @staticmethod
    def forward(ctx, spec, packed_fwd, packed_bwd, ped, obs, slf, drop_ped, drop_obs, *params):
        dev = L.require_cuda(packed_fwd, ped, slf)
        ped, slf = L.f32c(ped), L.f32c(slf)
        lead = slf.shape[:-1]
        R = 1
        for s in lead:
            R *= s
        kp = ped.shape[-2]
        if slf.dim() == 2:
            group = 0
        elif slf.dim() == 3:
            group = slf.shape[1]          # torch.norm(..., dim=1) reduces over the agents of a channel (model.py:1206)
        else:
            raise NotImplementedError("self_features must be (N,7) or (C,N,7)")
        ko = 0
        if spec.has_obs:
            obs = L.f32c(obs)
            ko = obs.shape[-2]
        else:
            obs = None
        desc = spec.desc()
        mw = spec.msg_width
        acc = torch.empty(*lead, 2, device=dev)
        pm = torch.empty(*lead, kp, mw, device=dev)
        om = torch.empty(*lead, ko, mw, device=dev) if spec.has_obs else None
        coll = torch.empty(*lead, kp, 1, device=dev) if spec.coll_dims else None
        stash = torch.empty(_floats("piml_pinnsf_stash_floats", desc, spec.has_obs, R, kp, ko), device=dev)
        dp = L.f32c(drop_ped) if drop_ped is not None else None
        do = L.f32c(drop_obs) if (drop_obs is not None and spec.has_obs) else None
        L.check(L.load().piml_pinnsf_forward_train_f32(
            L.C.byref(desc), L.ptr(packed_fwd), 1 if spec.has_obs else 0, spec.tau, L.ptr(ped), L.ptr(obs),
            L.ptr(slf), R, kp, ko, group, L.ptr(dp), L.ptr(do), L.ptr(acc), L.ptr(pm), L.ptr(om), L.ptr(coll),
            L.ptr(stash), L.stream_ptr(dev)), "piml_pinnsf_forward_train_f32")
        ctx.spec, ctx.dims = spec, (R, kp, ko, group)
        ctx.param_shapes = [p.shape for p in params]
        ctx.save_for_backward(packed_bwd, ped, obs, slf, dp, do, stash)
        outs = [acc, pm]
        if spec.has_obs:
            outs.append(om)
        if spec.coll_dims:
            outs.append(coll)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        packed_bwd, ped, obs, slf, dp, do, stash = ctx.saved_tensors
        spec = ctx.spec
        R, kp, ko, group = ctx.dims
        dev = ped.device
        grads = list(grads)
        g_acc = grads.pop(0)
        g_pm = grads.pop(0)
        g_om = grads.pop(0) if spec.has_obs else None
        g_coll = grads.pop(0) if spec.coll_dims else None
        g_acc = L.f32c(g_acc) if g_acc is not None else torch.zeros(R, 2, device=dev)
        g_pm = L.f32c(g_pm) if g_pm is not None else None
        g_om = L.f32c(g_om) if g_om is not None else None
        g_coll = L.f32c(g_coll) if g_coll is not None else None
        desc = spec.desc()
        n_par = sum(int(torch.Size(s).numel()) for s in ctx.param_shapes)
        g_params = torch.empty(n_par, device=dev)
        g_ped = torch.empty_like(ped) if ctx.needs_input_grad[3] else None
        g_obs = torch.empty_like(obs) if (obs is not None and ctx.needs_input_grad[4]) else None
        g_self = torch.empty_like(slf) if ctx.needs_input_grad[5] else None
        ws = torch.empty(_floats("piml_pinnsf_backward_workspace_floats", desc, spec.has_obs, R, kp, ko), device=dev)
        L.check(L.load().piml_pinnsf_backward_f32(
            L.C.byref(desc), L.ptr(packed_bwd), 1 if spec.has_obs else 0, spec.tau, L.ptr(ped), L.ptr(obs),
            L.ptr(slf), R, kp, ko, group, L.ptr(dp), L.ptr(do), L.ptr(stash), L.ptr(g_acc), L.ptr(g_pm), L.ptr(g_om),
            L.ptr(g_coll), L.ptr(g_params), L.ptr(g_ped), L.ptr(g_obs), L.ptr(g_self), L.ptr(ws), L.stream_ptr(dev)),
            "piml_pinnsf_backward_f32")
        pg, off = [], 0
        for shp in ctx.param_shapes:
            n = int(torch.Size(shp).numel())
            pg.append(g_params[off:off + n].view(shp))
            off += n
        return (None, None, None, g_ped, g_obs, g_self, None, None, *pg)


class RelativeFeaturesFunction(torch.autograd.Function):
    """`Pedestrians.get_relative_features` (data.py:466-512) with a CUDA backward: the forward kernel reports which
    agent / obstacle filled every slot; the backward scatters the feature gradients back to p, v, a and dest."""

    @staticmethod
    def forward(ctx, peds, position, velocity, acceleration, destination, obstacles, topk_ped, sight_angle_ped,
                dist_threshold_ped, topk_obs, sight_angle_obs, dist_threshold_obs):
        ped_f, obs_f, dest_f, sel = peds._relative_features_raw(
            position, velocity, acceleration, destination, obstacles, topk_ped, sight_angle_ped, dist_threshold_ped,
            topk_obs, sight_angle_obs, dist_threshold_obs, return_selection=True)
        ctx.save_for_backward(L.f32c(position), L.f32c(destination), sel[0], sel[2])
        return ped_f, obs_f, dest_f

    @staticmethod
    def backward(ctx, g_ped, g_obs, g_dest):
        pos, dest, pidx, oidx = ctx.saved_tensors
        dev = pos.device
        N = pos.shape[-2]
        B = pos.numel() // (2 * N)
        kp, ko = pidx.shape[-1], oidx.shape[-1]
        g_ped = L.f32c(g_ped) if g_ped is not None else torch.zeros(*pidx.shape, 6, device=dev)
        if ko:
            g_obs = L.f32c(g_obs) if g_obs is not None else torch.zeros(*oidx.shape, 6, device=dev)
        else:
            g_obs = None
        g_dest = L.f32c(g_dest) if g_dest is not None else None
        outs = [torch.empty_like(pos) for _ in range(4)]
        L.check(L.load().piml_relative_features_backward_f32(
            L.ptr(pos), L.ptr(dest), L.ptr(pidx), L.ptr(oidx) if ko else None, B, N, kp, ko, L.ptr(g_ped),
            L.ptr(g_obs), L.ptr(g_dest), *[L.ptr(o) for o in outs], L.stream_ptr(dev)),
            "piml_relative_features_backward_f32")
        g_pos, g_vel, g_acc, g_dst = outs
        return (None, g_pos, g_vel, g_acc, g_dst, None, None, None, None, None, None, None)


class IntegrateTrainFunction(torch.autograd.Function):
    """State update of the differentiable rollout (simulators.py:741-769): lagged explicit Euler, waypoint switch
    WITHOUT removal on arrival, teacher-forced entry.  Returns new tensors (p, v, a, dest, dest_idx); the last two
    carry no gradient."""

    @staticmethod
    def forward(ctx, p, v, a, a_next, dest, dest_idx, dest_num, waypoints, dt, entry, p_gt, v_gt, a_gt, dest_gt,
                dest_idx_gt):
        from .rollout import integrate_step
        p2, v2, a2 = L.f32c(p).clone(), L.f32c(v).clone(), L.f32c(a).clone()
        dest2, didx2 = L.f32c(dest).clone(), dest_idx.contiguous().clone()
        integrate_step(p2, v2, a2, a_next, dest2, didx2, dest_num, waypoints, dt, False, entry, p_gt, v_gt, a_gt,
                       dest_gt, dest_idx_gt)
        ctx.dt = float(dt)
        ctx.save_for_backward(entry.contiguous() if entry is not None else None)
        ctx.mark_non_differentiable(dest2, didx2)
        return p2, v2, a2, dest2, didx2

    @staticmethod
    def backward(ctx, g_p2, g_v2, g_a2, _gd, _gi):
        (entry,) = ctx.saved_tensors
        ref = next(g for g in (g_p2, g_v2, g_a2) if g is not None)
        dev = ref.device
        z = None
        gs = []
        for g in (g_p2, g_v2, g_a2):
            if g is None:
                z = z if z is not None else torch.zeros_like(ref)
                g = z
            gs.append(L.f32c(g))
        outs = [torch.empty_like(gs[0]) for _ in range(4)]
        L.check(L.load().piml_integrate_step_backward_f32(
            L.ptr(entry), gs[0].numel() // 2, ctx.dt, *[L.ptr(g) for g in gs], *[L.ptr(o) for o in outs],
            L.stream_ptr(dev)), "piml_integrate_step_backward_f32")
        return (outs[0], outs[1], outs[2], outs[3]) + (None,) * 11


class RolloutLossesFunction(torch.autograd.Function):
    """The 'sum'-reduced rollout losses of test_multiple_rollouts_for_training (simulators.py:172-249, called at
    :795-824) in one pass: returns a (3,) tensor [mse, collision, hard collision] (piml_rollout_losses_f32); the
    backward is one kernel (piml_rollout_losses_backward_f32).  `labels` may be a strided view whose last dimension is
    a slice of a wider tensor (data.labels[..., :2], data.labels[..., 4:6]): it is read in place."""

    @staticmethod
    def forward(ctx, pred, labels, time_decay, reverse, collisions, hard_collisions, abnormal_mask):
        dev = L.require_cuda(pred, labels)
        pred = L.f32c(pred)
        C, T, N = pred.shape[0], pred.shape[1], pred.shape[2]
        if labels.dtype != torch.float32 or labels.stride(-1) != 1 or labels.shape[:3] != pred.shape[:3] or \
                labels.stride(1) != N * labels.stride(2) or labels.stride(0) != T * N * labels.stride(2):
            labels = labels.float().contiguous()
        stride = labels.stride(2)
        coll = L.f32c(collisions) if collisions is not None else None
        hard = L.f32c(hard_collisions) if hard_collisions is not None else None
        ab = L.f32c(abnormal_mask).reshape(-1) if abnormal_mask is not None else None
        out = torch.empty(3, device=dev)
        ws = torch.empty(int(L.load().piml_rollout_losses_workspace_floats(C, N)), device=dev)
        L.check(L.load().piml_rollout_losses_f32(L.ptr(pred), L.ptr(labels), stride, C, T, N, float(time_decay),
                                                 1 if reverse else 0, L.ptr(coll), L.ptr(hard), L.ptr(ab), L.ptr(out),
                                                 L.ptr(ws), L.stream_ptr(dev)), "piml_rollout_losses_f32")
        ctx.save_for_backward(pred, labels, coll, hard, ab)
        ctx.cfg = (stride, C, T, N, float(time_decay), 1 if reverse else 0)
        return out

    @staticmethod
    def backward(ctx, g_out):
        pred, labels, coll, hard, ab = ctx.saved_tensors
        stride, C, T, N, time_decay, reverse = ctx.cfg
        g_out = L.f32c(g_out)
        g_pred = torch.empty_like(pred)
        L.check(L.load().piml_rollout_losses_backward_f32(
            L.ptr(pred), L.ptr(labels), stride, C, T, N, time_decay, reverse, L.ptr(coll), L.ptr(hard), L.ptr(ab),
            L.ptr(g_out), L.ptr(g_pred), L.stream_ptr(pred.device)), "piml_rollout_losses_backward_f32")
        return g_pred, None, None, None, None, None, None


class L1SumFunction(torch.autograd.Function):
    """l1_reg_loss(x, weight, 'sum') (simulators.py:169-170): weight * sum |x| -> scalar tensor."""

    @staticmethod
    def forward(ctx, x, weight):
        dev = L.require_cuda(x)
        x = L.f32c(x)
        out = torch.empty(1, device=dev)
        L.check(L.load().piml_l1_sum_f32(L.ptr(x), x.numel(), float(weight), L.ptr(out), L.stream_ptr(dev)),
                "piml_l1_sum_f32")
        ctx.save_for_backward(x)
        ctx.weight = float(weight)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        gx = torch.empty_like(x)
        g = L.f32c(g).reshape(1)
        L.check(L.load().piml_l1_sum_backward_f32(L.ptr(x), x.numel(), ctx.weight, L.ptr(g), L.ptr(gx),
                                                  L.stream_ptr(x.device)), "piml_l1_sum_backward_f32")
        return gx, None


class BceSumFunction(torch.autograd.Function):
    """F.binary_cross_entropy(pred, target, reduction='sum') and the number of correct rounded predictions
    (simulators.py:826-830) in one launch: returns (loss, hits) scalars; gradient to `pred` only."""

    @staticmethod
    def forward(ctx, pred, target):
        dev = L.require_cuda(pred, target)
        pred, target = L.f32c(pred), L.f32c(target)
        out = torch.empty(2, device=dev)
        L.check(L.load().piml_bce_sum_f32(L.ptr(pred), L.ptr(target), pred.numel(), L.ptr(out), L.stream_ptr(dev)),
                "piml_bce_sum_f32")
        ctx.save_for_backward(pred, target)
        loss, hits = out[0].clone(), out[1].clone()
        ctx.mark_non_differentiable(hits)
        return loss, hits

    @staticmethod
    def backward(ctx, g, _):
        pred, target = ctx.saved_tensors
        gp = torch.empty_like(pred)
        g = L.f32c(g).reshape(1)
        L.check(L.load().piml_bce_sum_backward_f32(L.ptr(pred), L.ptr(target), pred.numel(), L.ptr(g), L.ptr(gp),
                                                   L.stream_ptr(pred.device)), "piml_bce_sum_backward_f32")
        return gp, None

