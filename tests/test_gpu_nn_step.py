"""The fused NN rollout step (piml_nn_step_f32, csrc/nn_step.cu) against the three library calls it replaces
(state_features -> pinnsf_forward (tensor cores) -> integrate_step), which the other GPU tests pin to the oracle and to
the reference's golden vectors: every state tensor must be BIT-identical after every step, and the optional dense
features must equal the separate feature call's.  Reference: src/models/simulators.py:595-652.
"""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NET_ARGS = dict(model='pinnsf_bm', dataset_name='gc1560', dropout=0.5, encoder_hidden_size=128,
                processor_hidden_size=128, decoder_hidden_size=64, encoder_hidden_layers=3, processor_hidden_layers=16,
                decoder_hidden_layers=2, ped_feature_dim=6, obs_feature_dim=6, self_feature_dim=7)
FEAT = (6, 90, 4, 10, 90, 4)
DT = 0.08


def make_net(seed=666, cls="PINNSF_bottleneck_multitask", **over):
    from piml_b200 import models as M
    torch.manual_seed(seed)
    a = dict(NET_ARGS)
    a.update(over)
    net = getattr(M, cls)(argparse.Namespace(**a)).cuda().eval()
    return net, M.pack_device(net.state_dict(), net.spec), M.pack_device_tc(net.state_dict(), net.spec)


def crowd(S, N, M, seed, rho=0.5, absent=0.0, per_scene_obs=False):
    g = torch.Generator().manual_seed(seed)
    side = (N / rho) ** 0.5
    p = torch.rand(S, N, 2, generator=g) * side
    v = torch.randn(S, N, 2, generator=g) * 0.8
    a = torch.randn(S, N, 2, generator=g) * 0.3
    dest = torch.rand(S, N, 2, generator=g) * side
    ds = 1.0 + 0.3 * torch.rand(S, N, generator=g)
    obs = torch.rand(*((S, M, 2) if per_scene_obs else (M, 2)), generator=g) * side
    if absent > 0:
        gone = torch.rand(S, N, generator=g) < absent
        p[gone] = float('nan'); v[gone] = float('nan'); a[gone] = float('nan')
    # a few stationary agents (heading fallback), a few standing on their destination (arrival / removal), twins
    v[:, 3::97] = 0.0
    dest[:, 5::53] = p[:, 5::53] + 0.1
    if N > 40:
        p[:, 11] = p[:, 10]
    D = 3
    wp = torch.rand(S, D, N, 2, generator=g) * side
    wp[:, 0] = dest
    dnum = torch.randint(1, D + 1, (S, N), generator=g)
    return [x.cuda().contiguous() for x in (p, v, a, dest, ds, obs, wp, dnum)]


def three_calls(net, packed, packed_tc, st, obs, ds, wp, dnum, entry=None, gt=None, rec=None, remove=True):
    from piml_b200 import models as M
    from piml_b200.rollout import integrate_step, state_features
    p, v, a, dest, didx, hist = st
    S, N = p.shape[:2]
    pf, of, sf = state_features(p, v, a, dest, obs, hist, ds, *FEAT)
    kp, ko = pf.shape[2], of.shape[2]
    with torch.no_grad():
        acc = M.pinnsf_forward(net.spec, packed, pf.view(S * N, kp, 6), of.view(S * N, ko, 6), sf.view(S * N, 7),
                               need_msgs=False, packed_tc=packed_tc)[0].view(S, N, 2)
    kw = {}
    if entry is not None:
        kw = dict(entry=entry, p_gt=gt[0], v_gt=gt[1], a_gt=gt[2], dest_gt=gt[3], dest_idx_gt=gt[4])
    if rec is not None:
        kw.update(rec_p=rec[0], rec_v=rec[1], rec_a=rec[2], rec_mask=rec[3])
    integrate_step(p, v, a, acc, dest, didx, dnum, wp, DT, remove, hist_v=hist, **kw)
    return pf, of, sf, acc


def same(x, y):
    return np.array_equal(x.cpu().numpy(), y.cpu().numpy(), equal_nan=True)


@pytest.mark.parametrize("S,N,M,absent,per_scene", [(1, 20000, 2000, 0.02, False), (1, 5000, 50, 0.0, False),
                                                    (3, 700, 300, 0.3, True), (1, 100000, 2000, 0.0, False),
                                                    (2, 9, 40, 0.2, False)])
def test_fused_step_is_bit_identical_to_the_three_calls(S, N, M, absent, per_scene):
    from piml_b200.rollout import NNStep
    net, packed, packed_tc = make_net()
    p, v, a, dest, ds, obs, wp, dnum = crowd(S, N, M, seed=N + M, absent=absent, per_scene_obs=per_scene)
    didx = torch.zeros(S, N, dtype=torch.int64, device='cuda')
    dest = wp[:, 0].clone().contiguous()
    hist = torch.where(torch.isnan(v), torch.zeros_like(v), v).contiguous()
    A = [x.clone() for x in (p, v, a, dest, didx, hist)]              # three calls
    B = [x.clone() for x in (p, v, a, dest, didx, hist)]              # fused
    kp, ko = min(6, N), min(10, M)
    dense = (torch.empty(S, N, kp, 6, device='cuda'), torch.empty(S, N, ko, 6, device='cuda'),
             torch.empty(S, N, 7, device='cuda'), torch.empty(S, N, 2, device='cuda'))
    a_next = torch.empty(S, N, 2, device='cuda')
    step = NNStep(net.spec, packed_tc, *B, dnum, wp, ds, obs, DT, *FEAT, a_next=a_next, dense=dense)
    g = torch.Generator().manual_seed(1)
    for it in range(4):
        entry = gt = recA = recB = None
        if it >= 1:                                                   # teacher-forced entries + recording
            entry = (torch.rand(S, N, generator=g) < 0.05).long().cuda()
            gt = [(torch.rand(S, N, 2, generator=g) * 50).cuda() for _ in range(4)] + \
                 [torch.zeros(S, N, dtype=torch.int64, device='cuda')]
            recA = [torch.zeros(S, N, 2, device='cuda') for _ in range(3)] + [torch.zeros(S, N, device='cuda')]
            recB = [torch.zeros(S, N, 2, device='cuda') for _ in range(3)] + [torch.zeros(S, N, device='cuda')]
        pf, of, sf, acc = three_calls(net, packed, packed_tc, A, obs, ds, wp, dnum, entry, gt, recA)
        step.step(entry, gt, recB)
        torch.cuda.synchronize()
        assert same(dense[0], pf) and same(dense[2], sf), f"step {it}: dense features differ"
        if M:
            assert same(dense[1], of)
        assert same(a_next, acc), f"step {it}: model output differs"
        for name, x, y in zip(("p", "v", "a", "dest", "dest_idx", "hist_v"), A, B):
            assert same(x, y), f"step {it}: {name} differs"
        if recA is not None:
            for x, y in zip(recA, recB):
                assert same(x, y)
    assert torch.isfinite(B[0]).any()


def test_fused_step_without_dense_outputs_and_unsupported_networks():
    from piml_b200 import models as M
    from piml_b200.rollout import NNStep
    net, packed, packed_tc = make_net()
    S, N, Mo = 1, 30000, 2000
    p, v, a, dest, ds, obs, wp, dnum = crowd(S, N, Mo, seed=5)
    didx = torch.zeros(S, N, dtype=torch.int64, device='cuda')
    dest = wp[:, 0].clone().contiguous()
    hist = v.clone()
    A = [x.clone() for x in (p, v, a, dest, didx, hist)]
    B = [x.clone() for x in (p, v, a, dest, didx, hist)]
    step = NNStep(net.spec, packed_tc, *B, dnum, wp, ds, obs, DT, *FEAT)
    for it in range(3):
        three_calls(net, packed, packed_tc, A, obs, ds, wp, dnum)
        step.step()
    torch.cuda.synchronize()
    for x, y in zip(A, B):
        assert same(x, y)
    # the summed-embedding network is not a per-slot decoder: the fused step declines it
    net_m, _, ptc_m = make_net(cls="PINNSF_multitask", model='pinnsf_m')
    with pytest.raises(NotImplementedError):
        NNStep(net_m.spec, ptc_m, *B, dnum, wp, ds, obs, DT, *FEAT)
