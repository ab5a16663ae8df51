"""GPU parity of the training path: network backward, feature backward, Euler-chain backward,
collision_detection and the whole differentiable rollout -- through the host adapters -> ctypes -> C ABI.

References: (i) golden vectors produced by the UNMODIFIED reference (training_step.npz: module forward + autograd
backward; training_rollout.npz: BaseSimulator.test_multiple_rollouts_for_training + loss.backward()), (ii) the plain
PyTorch fp32 restatements in tests/torch_ref.py (pinned to (i) by the CPU suite) on seeded random inputs.
Tolerance (written here): 1e-5 relative for outputs/losses; gradients 2e-5 of the tensor's largest entry (fp32 sums over
up to ~10^4 rows in a different order than MKL's)."""
import argparse

import numpy as np
import pytest
import torch

from . import torch_ref as TR
from .golden_args import base_args, model_args
from .test_training_ref import CASES, drop_masks, load_case, reference_weights
from .util import golden, group

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda")


def max_rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-12))


def mirror(kind, dsn, sd=None, train=False, **over):
    from piml_b200 import models as M
    args = base_args(model=kind, dataset_name=dsn, **over)
    torch.manual_seed(666)
    net = M.CLASSES[kind](args)
    if sd is not None:
        net.load_state_dict({k: v for k, v in sd.items()}, strict=True)
    net = net.to(dev())
    net.train(train)
    return net


@pytest.mark.parametrize("case,kind,inp,train", CASES)
def test_pinnsf_backward_golden(case, kind, inp, train):
    from piml_b200 import models as M
    z = golden("training_step")
    g, ref_grads = load_case(z, case)
    sd, _ = reference_weights(kind, str(g["dataset_name"]))
    net = mirror(kind, str(g["dataset_name"]), sd, train)
    sfx = "_c" if inp == "ped_c" else ""
    ped, obs, slf = [torch.from_numpy(z[n + sfx]).to(dev()).requires_grad_(True) for n in ("ped", "obs", "self")]
    if train:
        dp, do = [m.to(dev()) for m in drop_masks(g)]
        outs = M.pinnsf_forward_autograd(net, net.spec, net._train_cache, ped, obs, slf, dp, do)
    else:
        outs = net(ped, obs, slf)
    loss = sum((o * torch.from_numpy(g[f"w{i}"]).to(dev())).sum() for i, o in enumerate(outs))
    loss.backward()
    for i, o in enumerate(outs):
        assert max_rel(o.detach().cpu().numpy(), g[f"out{i}"]) < 1e-5, f"out{i}"
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5 * max(1.0, abs(float(g["loss"]))) * 20
    for name, t in (("g_ped", ped), ("g_obs", obs), ("g_self", slf)):
        assert max_rel(t.grad.cpu().numpy(), g[name]) < 2e-5, name
    named = dict(net.named_parameters())
    for k, ref in ref_grads.items():
        assert named[k].grad is not None, k
        assert max_rel(named[k].grad.cpu().numpy(), ref.numpy()) < 2e-5, k
    for k in g["dead"].tolist():                      # ResDNN block 0: grad stays None like the reference
        assert named[k].grad is None, k


@pytest.mark.parametrize("kind,R,kp,ko,small,has_obs,chan", [
    ("pinnsf_bm", 301, 6, 10, False, True, 0), ("pinnsf_m", 77, 6, 10, False, True, 0),
    ("pinnsf_bm", 64, 3, 2, True, True, 4), ("pinnsf_m", 45, 5, 0, True, False, 0),
    ("pinnsf_bottleneck", 130, 6, 10, True, True, 0), ("pinnsf", 50, 4, 7, False, True, 5),
    ("pinnsf_bm", 140, 6, 10, "single", True, 0), ("pinnsf_m", 90, 6, 4, "single", True, 0)])
def test_pinnsf_backward_random(kind, R, kp, ko, small, has_obs, chan):
    """Random inputs / shapes vs torch autograd of the plain-torch restatement (tile tails, no obstacle branch,
    narrow nets, channelled destination norm, train-mode dropout multipliers)."""
    over = {}
    if small:
        over = dict(encoder_hidden_size=32, processor_hidden_size=32, decoder_hidden_size=16, encoder_hidden_layers=2,
                    processor_hidden_layers=4, decoder_hidden_layers=1)
    if small == "single":                                   # one processor block: relu(W e + b) + e instead of 2e
        over["processor_hidden_layers"] = 1
    if not has_obs:
        over["obs_feature_dim"] = 0
    from piml_b200 import models as M
    net = mirror(kind, "gc1560", None, True, **over)
    g = torch.Generator().manual_seed(R * 7 + kp)
    lead = (chan, R // chan) if chan else (R,)
    ped = torch.randn(*lead, kp, 6, generator=g)
    ped[..., -1, :] = 0                                                      # a zero-padded slot (f(0) path)
    obs = torch.randn(*lead, max(ko, 1), 6, generator=g)[..., :ko, :]
    slf = torch.randn(*lead, 7, generator=g)
    slf[..., 0, :2] = 0                                                      # zero destination vector (norm guard)
    dp = (torch.rand(*lead, kp, net.spec.pw, generator=g) > 0.5).float() * 2
    do = (torch.rand(*lead, ko, net.spec.pw, generator=g) > 0.5).float() * 2
    sd_cpu = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    ins_cpu = [x.clone().requires_grad_(True) for x in (ped, obs, slf)]
    ref = TR.pinnsf_forward_ref(sd_cpu, kind, net.spec.tau, *ins_cpu, has_obs, dp, do if has_obs else None,
                                single_block=net.spec.proc_mode == 1)
    ws = [torch.randn(o.shape, generator=g) for o in ref]
    sum((o * w).sum() for o, w in zip(ref, ws)).backward()
    ins = [x.to(dev()).requires_grad_(True) for x in (ped, obs, slf)]
    outs = M.pinnsf_forward_autograd(net, net.spec, net._train_cache, ins[0], ins[1] if has_obs else None, ins[2],
                                     dp.to(dev()), do.to(dev()) if has_obs else None)
    sum((o * w.to(dev())).sum() for o, w in zip(outs, ws)).backward()
    for o, r in zip(outs, ref):
        assert max_rel(o.detach().cpu().numpy(), r.detach().numpy()) < 1e-5
    for i, (a, b) in enumerate(zip(ins, ins_cpu)):
        if i == 1 and not has_obs:
            continue
        assert max_rel(a.grad.cpu().numpy(), b.grad.numpy()) < 2e-5, i
    named = dict(net.named_parameters())
    for k in M.linear_keys(net.spec):
        if k.startswith("obs") and not has_obs:
            continue
        for s in (".weight", ".bias"):
            assert max_rel(named[k + s].grad.cpu().numpy(), sd_cpu[k + s].grad.numpy()) < 2e-5, k + s


@pytest.mark.parametrize("B,N,M,kp,ko", [(1, 300, 50, 6, 10), (7, 40, 0, 6, 10), (3, 5, 4, 6, 10)])
def test_relative_features_backward(B, N, M, kp, ko):
    """The autograd path of get_relative_features vs torch autograd of the gather restatement on the same selection."""
    import piml_b200 as P
    g = torch.Generator().manual_seed(B * 100 + N)
    L = max((N / 0.5) ** 0.5, 3.0)
    p = torch.rand(B, N, 2, generator=g) * L
    d = torch.rand(B, N, 2, generator=g) * L
    p[:, ::11] = float("nan"); d[:, ::11] = float("nan")
    v = torch.randn(B, N, 2, generator=g); a = torch.randn(B, N, 2, generator=g)
    obs = torch.rand(M, 2, generator=g) * L
    peds = P.Pedestrians()
    ins = [x.to(dev()).requires_grad_(True) for x in (p, v, a, d)]
    # (c = B, t = 1, N, 2): the shape the differentiable rollout passes (simulators.py:772-776)
    pf, of, df = peds.get_relative_features(ins[0][:, None], ins[1][:, None], ins[2][:, None], ins[3][:, None],
                                            obs.to(dev()), kp, 90, 4, ko, 90, 4)
    _, _, _, sel = peds.get_relative_features(*[x.detach().clone()[:, None] for x in ins], obs.to(dev()), kp, 90, 4,
                                              ko, 90, 4, return_selection=True)
    ws = [torch.randn(x.shape, generator=g) for x in (pf, of, df)]
    sum((x * w.to(dev())).sum() for x, w in zip((pf, of, df), ws)).backward()
    cin = [x.clone().requires_grad_(True) for x in (p, v, a, d)]
    oidx = sel[2][:, 0].cpu() if M else torch.zeros(B, N, 0, dtype=torch.int64)
    rp, ro, rd = TR.gathered_features_ref(*cin, obs, sel[0][:, 0].cpu(), oidx)
    assert torch.equal(rp, pf[:, 0].detach().cpu()) and torch.equal(rd, df[:, 0].detach().cpu())
    loss = (rp * ws[0][:, 0]).sum() + (rd * ws[2][:, 0]).sum()
    if M:
        assert torch.equal(ro, of[:, 0].detach().cpu())
        loss = loss + (ro * ws[1][:, 0]).sum()
    loss.backward()
    for x, c in zip(ins, cin):
        assert max_rel(x.grad.cpu().numpy(), c.grad.numpy()) < 1e-5


def test_integrate_backward():
    from piml_b200.autograd import IntegrateTrainFunction
    g = torch.Generator().manual_seed(5)
    S, N, D, dt = 3, 17, 2, 0.08
    p, v, a, an = [torch.randn(S, N, 2, generator=g) for _ in range(4)]
    dest = torch.randn(S, N, 2, generator=g); wp = torch.randn(S, D, N, 2, generator=g)
    didx = torch.zeros(S, N, dtype=torch.int64); dnum = torch.full((N,), 2, dtype=torch.int64)
    entry = (torch.rand(S, N, generator=g) < 0.3).long()
    gt = [torch.randn(S, N, 2, generator=g) for _ in range(4)]
    ins = [x.to(dev()).requires_grad_(True) for x in (p, v, a, an)]
    out = IntegrateTrainFunction.apply(*ins, dest.to(dev()), didx.to(dev()), dnum.to(dev()), wp.to(dev()), dt,
                                       entry.to(dev()), *[x.to(dev()) for x in gt], didx.to(dev()))
    ws = [torch.randn(S, N, 2, generator=g) for _ in range(3)]
    sum((o * w.to(dev())).sum() for o, w in zip(out[:3], ws)).backward()
    cin = [x.clone().requires_grad_(True) for x in (p, v, a, an)]
    keep = (entry == 0).unsqueeze(-1)
    v2 = torch.where(keep, cin[1] + cin[2] * dt, gt[1])
    p2 = torch.where(keep, cin[0] + cin[1] * dt, gt[0])
    a2 = torch.where(keep, cin[3], gt[2])
    assert torch.equal(out[0].detach().cpu(), p2.detach()) and torch.equal(out[1].detach().cpu(), v2.detach())
    ((p2 * ws[0]).sum() + (v2 * ws[1]).sum() + (a2 * ws[2]).sum()).backward()
    for x, c in zip(ins, cin):
        assert max_rel(x.grad.cpu().numpy(), c.grad.numpy()) < 1e-6


@pytest.mark.parametrize("shape,thr,real", [((32, 60, 2), 0.5, False), ((40, 25, 2), 0.8, True),
                                            ((5, 8, 30, 2), 0.5, False), ((3, 2, 9, 2), 0.25, False)])
def test_collision_detection(shape, thr, real):
    """Pedestrians.collision_detection (data.py:538-601): 0/1 matrix bit-exact, friends rule for 3-d and 4-d inputs."""
    import piml_b200 as P
    g = torch.Generator().manual_seed(sum(shape))
    p = torch.rand(*shape, generator=g) * 3
    p[..., 1, :] = p[..., 0, :] + 0.05                       # a permanent pair: "friends"
    p.view(-1, shape[-2], 2)[::3, 4] = float("nan")
    rp = (p + 0.3 * torch.randn(*shape, generator=g)) if real else None
    ref = TR.collision_detection_ref(p, thr, rp)
    got = P.Pedestrians.collision_detection(p.to(dev()), thr, rp.to(dev()) if real else None)
    assert torch.equal(got.cpu(), ref)
    rows = P.Pedestrians.collision_detection(p.to(dev()), thr, rp.to(dev()) if real else None, rowsum_only=True)
    assert torch.equal(rows.cpu(), ref.sum(-1))


class _Batch(object):
    pass


def _batch_from_golden(g):
    b = _Batch()
    for k, v in g.items():
        if k.startswith("in/") and v.dtype.kind in "fi" and v.ndim > 0:
            setattr(b, k[3:], torch.from_numpy(v).to(dev()))
    b.time_unit = float(g["in/time_unit"])
    b.num_frames = int(g["in/num_frames"])
    return b


@pytest.mark.parametrize("case", ["ucy_bm", "gc_bm_full"])
def test_training_rollout_golden(case):
    """test_multiple_rollouts_for_training + loss.backward() vs the reference's own run on the same batch."""
    from piml_b200 import train_rollout as TRO
    g = group(golden("training_rollout"), case)
    kind, dsn = str(g["in/model"]), str(g["in/dataset_name"])
    a = g["in/args"]
    args = base_args(model=kind, dataset_name=dsn, reg_weight=float(a[0]), collision_threshold=float(a[1]),
                     collision_loss_weight=float(a[2]), hard_collision_penalty=float(a[3]), teacher_weight=float(a[4]),
                     collision_pred_weight=float(a[5]), collision_focus_weight=float(a[6]),
                     new_collision_loss_flag=int(a[7]), time_decay=float(a[8]),
                     collision_loss_version=str(g["in/collision_loss_version"]))
    sim = argparse.Namespace(args=args, model=mirror(kind, dsn, None, False), collision_count=0,
                             hard_collision_count=0, epoch=0, batch_idx=0)
    batch = _batch_from_golden(g)
    res = TRO.test_multiple_rollouts_for_training(sim, batch)
    res[0].backward()
    for i, r in enumerate(res):
        ref = float(g[f"out{i}"])
        assert abs(float(r.detach()) - ref) <= 2e-5 * max(abs(ref), 1e-3), (i, float(r.detach()), ref)
    assert sim.collision_count == float(g["collision_count"])
    assert sim.hard_collision_count == float(g["hard_collision_count"])
    assert np.array_equal(batch.dest_idx.cpu().numpy(), g["dest_idx_after"])
    named = dict(sim.model.named_parameters())
    worst = 0.0
    for k, v in g.items():
        if k.startswith("grad/"):
            worst = max(worst, max_rel(named[k[5:]].grad.cpu().numpy(), v))
    assert worst < 5e-5, worst


@pytest.mark.parametrize("C,T,N,reverse,with_coll,with_mask", [(32, 10, 144, False, True, False),
                                                               (6, 5, 122, False, True, True),
                                                               (4, 7, 33, True, False, False),
                                                               (1, 1, 5, False, True, False)])
def test_fused_rollout_losses_match_the_torch_restatement(C, T, N, reverse, with_coll, with_mask):
    """piml_rollout_losses_f32 / _backward_f32 (simulators.py:172-249 with reduction 'sum', fused) against the plain
    torch restatement of the same three losses (tests/torch_ref.py multiple_rollout_*), values and d/d pred, with
    labels read in place from a wider (.., 6 + k) tensor like data.labels[..., :2]."""
    from tests import torch_ref as TR
    from piml_b200.autograd import RolloutLossesFunction
    g = torch.Generator().manual_seed(C * 100 + T)
    pred = (torch.randn(C, T, N, 2, generator=g) * 3).cuda().requires_grad_(True)
    wide = (torch.randn(C, T, N, 12, generator=g) * 3).cuda()
    labels = wide[..., 4:6] if reverse else wide[..., :2]
    coll = hard = am = None
    if with_coll:
        coll = (torch.rand(C, T, N, generator=g) < 0.1).float().cuda() * 2
        hard = (torch.rand(C, T, N, generator=g) < 0.03).float().cuda()
    if with_mask:
        am = (torch.rand(N, generator=g) < 0.7).float().cuda()
    decay = 0.9
    out = RolloutLossesFunction.apply(pred, labels, decay, reverse, coll, hard, am)
    gw = torch.tensor([1.0, 10.0, 100.0]).cuda()
    (out * gw).sum().backward()
    got_g = pred.grad.clone()
    pred.grad = None
    want0 = TR.multiple_rollout_mse_loss(pred, labels, decay, 'sum', reverse=reverse)
    want = [want0, torch.zeros((), device="cuda"), torch.zeros((), device="cuda")]
    if with_coll:
        want[1] = TR.multiple_rollout_collision_loss(pred, labels, decay, 10, coll.clone(), 'sum', am)
        want[2] = TR.multiple_rollout_collision_loss(pred, labels, decay, 10, hard.clone(), 'sum', am)
    (want[0] * gw[0] + want[1] * gw[1] + want[2] * gw[2]).backward()
    for q in range(3):
        ref = float(want[q])
        assert abs(float(out[q]) - ref) <= 2e-5 * max(abs(ref), 1.0), (q, float(out[q]), ref)
    scale = float(pred.grad.abs().max())
    assert float((got_g - pred.grad).abs().max()) <= 2e-5 * max(scale, 1.0)


@pytest.mark.parametrize("case", ["ucy", "gc", "tiny"])
def test_fused_rollout_losses_match_reference_methods(case):
    """f-3 pinned to the reference: the fused loss kernels against the values and gradients of the reference's own
    BaseSimulator.multiple_rollout_* methods (tests/golden/make_golden.py losses)."""
    from piml_b200.autograd import RolloutLossesFunction
    g = group(golden("losses"), case)
    wide = torch.from_numpy(g["wide"]).to(dev())
    pred = torch.from_numpy(g["pred"]).to(dev()).requires_grad_(True)
    a_pred = torch.from_numpy(g["a_pred"]).to(dev()).requires_grad_(True)
    am = torch.from_numpy(g["abnormal_mask"]).to(dev()) if "abnormal_mask" in g else None
    decay = float(g["decay"])
    out = RolloutLossesFunction.apply(pred, wide[..., :2], decay, False, torch.from_numpy(g["coll"]).to(dev()),
                                      torch.from_numpy(g["hard"]).to(dev()), am)
    (out * torch.tensor([1.0, 10.0, 100.0], device=dev())).sum().backward()
    a_out = RolloutLossesFunction.apply(a_pred, wide[..., 4:6], decay, True, None, None, None)
    a_out[0].backward()
    for got, key in ((out[0], "mse"), (out[1], "collision"), (out[2], "hard_collision"), (a_out[0], "a_mse")):
        ref = float(g[key])
        assert abs(float(got.detach()) - ref) <= 2e-5 * max(abs(ref), 1.0), (key, float(got.detach()), ref)
    scale = float(np.abs(g["g_pred"]).max())
    assert float(np.abs(pred.grad.cpu().numpy() - g["g_pred"]).max()) <= 2e-5 * max(scale, 1.0)
    assert np.allclose(a_pred.grad.cpu().numpy(), g["g_a_pred"], rtol=2e-5, atol=2e-6)


def test_l1_and_bce_sum_kernels_match_torch():
    """L1SumFunction (simulators.py:169-170) and BceSumFunction (:826-830) against torch's own ops, values and
    gradients, incl. saturated predictions (log clamp at -100) and exact zeros."""
    from piml_b200.autograd import BceSumFunction, L1SumFunction
    g = torch.Generator().manual_seed(3)
    x = torch.randn(6, 5, 144, 2, generator=g).cuda()
    x[0, 0, :7] = 0
    x.requires_grad_(True)
    out = L1SumFunction.apply(x, 1e-3)
    (out * 3).backward()
    xr = x.detach().clone().requires_grad_(True)
    ref = TR.l1_reg_loss(xr, 1e-3, 'sum')
    (ref * 3).backward()
    assert abs(float(out) - float(ref)) <= 2e-6 * abs(float(ref))
    assert torch.equal(x.grad, xr.grad)
    p = torch.rand(6, 5, 144, 6, generator=g).cuda()
    p.view(-1)[:4] = torch.tensor([0.0, 1.0, 1e-30, 0.5]).cuda()
    t = (torch.rand(6, 5, 144, 6, generator=g) < 0.2).float().cuda()
    p.requires_grad_(True)
    loss, hits = BceSumFunction.apply(p, t)
    (loss * 10).backward()
    pr = p.detach().clone().requires_grad_(True)
    want = torch.nn.functional.binary_cross_entropy(pr, t, reduction='sum')
    (want * 10).backward()
    assert abs(float(loss) - float(want)) <= 2e-6 * abs(float(want))
    assert float(hits) == float(torch.sum(torch.round(pr.detach()) == t))
    assert torch.allclose(p.grad, pr.grad, rtol=1e-6, atol=0)


def test_graphed_training_step_equals_the_eager_step():
    """piml_b200.train_graph: rollout + backward + Adam captured in ONE CUDA graph must do what the eager step does: same
    losses, same collision counters, same parameters after several steps (eval-mode network: no dropout RNG)."""
    from piml_b200 import train_rollout as TRO
    from piml_b200.train_graph import GraphedRolloutTraining
    g = group(golden("training_rollout"), "ucy_bm")
    kind, dsn = str(g["in/model"]), str(g["in/dataset_name"])
    a = g["in/args"]
    args = base_args(model=kind, dataset_name=dsn, reg_weight=float(a[0]), collision_threshold=float(a[1]),
                     collision_loss_weight=float(a[2]), hard_collision_penalty=float(a[3]), teacher_weight=float(a[4]),
                     collision_pred_weight=float(a[5]), collision_focus_weight=float(a[6]),
                     new_collision_loss_flag=int(a[7]), time_decay=float(a[8]),
                     collision_loss_version=str(g["in/collision_loss_version"]))

    def make():
        net = mirror(kind, dsn, None, False)
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, capturable=True)
        sim = argparse.Namespace(args=args, model=net, collision_count=0, hard_collision_count=0, epoch=0, batch_idx=0)
        return net, opt, sim
    fresh = lambda: _batch_from_golden(g)
    net_e, opt_e, sim_e = make()
    net_g, opt_g, sim_g = make()
    WARM = 3
    graphed = GraphedRolloutTraining(sim_g, opt_g, fresh(), warmup=WARM)
    sim_g.collision_count = sim_g.hard_collision_count = 0
    for _ in range(WARM):                                         # the eager twin takes the same number of steps
        opt_e.zero_grad(set_to_none=True)
        TRO.test_multiple_rollouts_for_training(sim_e, fresh())[0].backward()
        opt_e.step()
    sim_e.collision_count = sim_e.hard_collision_count = 0
    for it in range(3):
        opt_e.zero_grad(set_to_none=True)
        b = fresh()
        res_e = TRO.test_multiple_rollouts_for_training(sim_e, b)
        res_e[0].backward()
        opt_e.step()
        bg = fresh()
        res_g = graphed.step(bg)
        for i, (x, y) in enumerate(zip(res_e, res_g)):
            x, y = float(x.detach()), float(y.detach())
            assert abs(x - y) <= 1e-5 * max(abs(x), 1e-3), (it, i, x, y)
    assert sim_e.collision_count == sim_g.collision_count
    assert sim_e.hard_collision_count == sim_g.hard_collision_count
    for (k, p), (_, q) in zip(net_e.named_parameters(), net_g.named_parameters()):
        assert max_rel(q.detach().cpu().numpy(), p.detach().cpu().numpy()) < 1e-5, k
