"""Differentiable training rollout on the CUDA path: drop-in for
`BaseSimulator.test_multiple_rollouts_for_training` (reference src/models/simulators.py:659-832) and host mirrors of
the rollout losses it calls (:169-249).

Per step the reference runs the model, four N x N `collision_detection` passes, the Euler / waypoint / entry
bookkeeping and a differentiable `get_relative_features`, all as eager ops recorded by autograd.  Here every one of
those stages is a kernel of libpiml_b200.so wrapped in a `torch.autograd.Function` (piml_b200/autograd.py), so
`loss.backward()` (:359) also runs in the library: network backward (dX chain, dW, db), feature scatter, Euler chain.
The 'sum'-reduced position / collision / teacher losses (:795-824) are one fused kernel each way
(`RolloutLossesFunction`), the L1 regulariser (:735-737) and the collision-prediction BCE with its
accuracy (:826-830) one small kernel each (`L1SumFunction`, `BceSumFunction`).  Plain-torch restatements of these
losses live in tests/torch_ref.py as the reference of the kernels' tests -- nothing here computes a loss in eager torch.
"""
import torch

from . import _lib as L
from .autograd import BceSumFunction, IntegrateTrainFunction, L1SumFunction, RolloutLossesFunction
from .features import Pedestrians

_PEDS = Pedestrians()


# ---- the rollout ------------------------------------------------------------------------------------------------
def test_multiple_rollouts_for_training(simulator, data, t_start=0, _sync_free=False):
    """Drop-in body for `BaseSimulator.test_multiple_rollouts_for_training(self, data, t_start=0)`.

    `_sync_free` (used by piml_b200.train_graph, which captures the whole training step in a CUDA graph): no device ->
    host reads inside the call -- the per-frame `torch.sum(mask) > 0` test of :705 is taken as true (the caller checks
    it for the whole batch before replaying), the NaN assert of :745 and the two collision counters become device
    tensors in `simulator._deferred = (nan_flag, collision_sum, hard_collision_sum)` for the caller to read afterwards.

    `simulator` provides .args, .model (a reference PINNSF module patched by piml_b200.patch, or a piml_b200.models
    mirror) and the counters .collision_count / .hard_collision_count.  `data` is a channelled clip
    (`ChanneledTimeIndexedPedData`, src/data/data.py:1046-1160) whose tensors live on a CUDA device.
    Returns (loss, mse_loss, collision_loss, hard_collision_loss, collision_pred_loss, collision_pred_acc, reg_loss)
    exactly like the reference; `loss.backward()` then runs the library's backward kernels.
    """
    args = simulator.args
    model = simulator.model
    dev = L.require_cuda(data.position, data.velocity, data.acceleration)
    if data.position.dim() != 4:
        raise NotImplementedError("training rollouts take channelled data (c,t,N,2)")
    C, _, N = data.position.shape[:3]
    if C == 1 or N == 1:
        raise NotImplementedError("the reference's .squeeze() calls collapse c == 1 / N == 1 (SURVEY.md B-11)")

    waypoints = L.f32c(data.waypoints)
    obstacles = data.obstacles
    mask_p_ = data.mask_p_pred.clone().long()                                    # c, t, n  (:676)
    state_features = [data.ped_features[..., t_start, :, :, :], data.obs_features[..., t_start, :, :, :],
                      data.self_features[..., t_start, :, :]]
    desired_speed = state_features[-1][..., -1].unsqueeze(-1)                     # c, n, 1  (:680)
    a_cur = data.acceleration[..., t_start, :, :]
    v_cur = data.velocity[..., t_start, :, :]
    p_cur = data.position[..., t_start, :, :]
    dest_cur = data.destination[..., t_start, :, :]
    dest_idx_cur = data.dest_idx[..., t_start, :]
    dest_num = data.dest_num
    new_peds_flag = (data.mask_p - data.mask_p_pred).long()                      # c, t, n  (:688)

    loss = torch.zeros((), device=dev, requires_grad=True)      # (no host scalar: the step may be captured in a graph)
    p_res = torch.zeros(data.position.shape, device=dev)
    collisions = torch.zeros(mask_p_.shape, device=dev)
    hard_collisions = torch.zeros(mask_p_.shape, device=dev)
    label_collisions = torch.zeros(mask_p_.shape, device=dev)
    label_hard_collisions = torch.zeros(mask_p_.shape, device=dev)
    a_res = torch.zeros(data.acceleration.shape, device=dev)
    pred_collisions = torch.zeros(data.ped_features[..., 0].shape, device=dev)
    true_collision = torch.zeros(data.ped_features[..., 0].shape, device=dev)
    reg_loss = torch.zeros((), device=dev)
    thr = args.collision_threshold
    T = data.num_frames
    nan_flag = torch.zeros((), dtype=torch.bool, device=dev)
    for t in range(t_start, T):
        predictions = model(*state_features)                                      # :701  CUDA fwd (+ stash)
        p_msg = predictions[1]
        mask = mask_p_[:, t, :]
        if _sync_free or torch.sum(mask) > 0:                                     # :705
            p_det = p_cur.clone().detach()
            lab = data.labels[:, t, :, :2]
            collisions[:, t, :] = _PEDS.collision_detection(p_det, thr, rowsum_only=True)            # :707-709
            hard_collisions[:, t, :] = _PEDS.collision_detection(p_det, thr / 2, rowsum_only=True)   # :711-714
            label_collisions[:, t, :] = _PEDS.collision_detection(lab, thr, rowsum_only=True)        # :716-719
            label_hard_collisions[:, t, :] = _PEDS.collision_detection(lab, thr / 2, rowsum_only=True)
            p_res[:, t, ...] = p_cur                                              # :728-729
            a_res[:, t, ...] = a_cur
            if args.collision_pred_weight > 0 and args.model == 'pinnsf_bm':      # :731-733
                pred_collisions[:, t, ...] = predictions[-1]
                true_collision[:, t, ...] = Pedestrians.calculate_collision_label(state_features[0])
            if args.reg_weight > 0:                                               # :735-737 (running sum, as is)
                reg_loss = reg_loss + L1SumFunction.apply(p_msg, args.reg_weight)
                loss = loss + reg_loss
        a_next = predictions[0]
        if _sync_free:
            nan_flag = nan_flag | a_next.isnan().any()
        else:
            assert ~a_next.isnan().any(), print('find nan in epoch :', getattr(simulator, 'epoch', None),
                                                getattr(simulator, 'batch_idx', None))          # :745
        # Euler with the old a and v, waypoint switch without removal, teacher-forced entry   (:741-769)
        last = t >= T - 1
        entry = None if last else new_peds_flag[..., t + 1, :]
        p_cur, v_cur, a_cur, dest_cur, dest_idx_new = IntegrateTrainFunction.apply(
            p_cur, v_cur, a_cur, a_next, dest_cur, dest_idx_cur, dest_num, waypoints, data.time_unit, entry,
            None if last else data.position[..., t + 1, :, :], None if last else data.velocity[..., t + 1, :, :],
            None if last else data.acceleration[..., t + 1, :, :],
            None if last else data.destination[..., t + 1, :, :], None if last else data.dest_idx[..., t + 1, :])
        dest_idx_cur.copy_(dest_idx_new)                       # the reference updates this view of the data in place
        # features of the new state (:772-778)
        ped_features, obs_features, dest_features = _PEDS.get_relative_features(
            p_cur.unsqueeze(-3), v_cur.unsqueeze(-3), a_cur.unsqueeze(-3), dest_cur.unsqueeze(-3), obstacles,
            args.topk_ped, args.sight_angle_ped, args.dist_threshold_ped, args.topk_obs, args.sight_angle_obs,
            args.dist_threshold_obs)
        self_features = torch.cat((dest_features.squeeze(), v_cur, a_cur, desired_speed), dim=-1)
        state_features = [ped_features.squeeze(), obs_features.squeeze(), self_features]

    if args.new_collision_loss_flag:                                              # :781-787
        label_collisions = torch.sum(label_collisions, dim=-2, keepdim=True).repeat(1, collisions.shape[1], 1)
        label_hard_collisions = torch.sum(label_hard_collisions, dim=-2, keepdim=True).repeat(
            1, hard_collisions.shape[1], 1)
        collisions = collisions.masked_fill(label_collisions > 0, 0.)
        hard_collisions = hard_collisions.masked_fill(label_hard_collisions > 0, 0.)
    if _sync_free:
        simulator._deferred = (nan_flag, torch.sum(collisions), torch.sum(hard_collisions))
    else:
        simulator.collision_count = getattr(simulator, 'collision_count', 0) + torch.sum(collisions).item()
        simulator.hard_collision_count = getattr(simulator, 'hard_collision_count', 0) + \
            torch.sum(hard_collisions).item()

    p_res.masked_fill_((mask_p_ == 0).unsqueeze(-1), 0.)                          # :792-793
    data.labels.masked_fill_((mask_p_ == 0).unsqueeze(-1), 0.)
    labels_p = data.labels[:, :, :, :2]
    # the three 'sum'-reduced position losses of :795-813 in ONE fused pass (piml_rollout_losses_f32); the collision
    # weights are only needed when their loss is switched on
    zero = torch.zeros((), device=dev)
    collision_loss, hard_collision_loss, collision_pred_loss, collision_pred_acc = zero, zero, zero, zero
    want_coll = args.collision_loss_weight > 0 and args.collision_loss_version in ('v0', 'v2')
    am = data.abnormal_mask if (want_coll and args.collision_loss_version == 'v2') else None
    fused = RolloutLossesFunction.apply(p_res, labels_p, args.time_decay, False, collisions if want_coll else None,
                                        hard_collisions if want_coll else None, am)
    mse_loss = fused[0]
    loss = loss + mse_loss
    if args.collision_loss_weight > 0:                                            # :799-819
        if want_coll:
            collision_loss, hard_collision_loss = fused[1], fused[2]
        collision_loss = collision_loss * args.collision_loss_weight
        hard_collision_loss = hard_collision_loss * args.collision_loss_weight * args.hard_collision_penalty
        loss = loss + collision_loss + hard_collision_loss
    if args.teacher_weight > 0:                                                   # :821-824
        a_mse_loss = RolloutLossesFunction.apply(a_res, data.labels[..., 4:6], args.time_decay, True, None, None,
                                                 None)[0]
        loss = loss + a_mse_loss * args.teacher_weight
    if args.collision_pred_weight > 0:                                            # :826-830
        bce, hits = BceSumFunction.apply(pred_collisions, true_collision)
        collision_pred_loss = bce * args.collision_pred_weight
        collision_pred_acc = hits / true_collision.numel()
        loss = loss + collision_pred_loss
    return loss, mse_loss, collision_loss, hard_collision_loss, collision_pred_loss, collision_pred_acc, reg_loss
