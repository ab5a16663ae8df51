"""Small invocations of every kernel family of libpiml_b200.so, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck / initcheck); sizes are tiny because the tools slow kernels down 10-1000x.

    compute-sanitizer --tool memcheck python scripts/sanitize_workload.py [--part all|smoke|rollout|train|misc]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import piml_b200 as P  # noqa: E402
from piml_b200 import models as M  # noqa: E402
from tests.util import golden, group  # noqa: E402


def bm_args(model='pinnsf_bm'):
    return argparse.Namespace(model=model, dataset_name='gc1560', dropout=0.5, encoder_hidden_size=128,
                              processor_hidden_size=128, decoder_hidden_size=64, encoder_hidden_layers=3,
                              processor_hidden_layers=16, decoder_hidden_layers=2, ped_feature_dim=6,
                              obs_feature_dim=6, self_feature_dim=7, topk_ped=6, topk_obs=10, sight_angle_ped=90,
                              sight_angle_obs=90, dist_threshold_ped=4, dist_threshold_obs=4, time_unit=0.08)


def cu(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x), dtype=dtype).cuda()


def part_smoke():
    import __graft_entry__ as G
    G.smoke()


def part_rollout():
    """Whole get_multiple_rollouts loops (C side): NN (tensor-core and FP32-pipe forward) and the persistent SFM kernel."""
    from piml_b200.rollout import rollout_scenes
    for name, frames in (("rollout_gc_bm", 40), ("rollout_syn_sfm", 60)):
        g = golden(name)
        t0 = int(g["in/t_start"])
        sc = {k: cu(g["in/" + k])[None] for k in ("position", "velocity", "acceleration", "destination", "waypoints",
                                                   "mask_p", "mask_p_pred")}
        sc["dest_idx"] = cu(g["in/dest_idx"], torch.int64)[None]
        sc["dest_num"] = cu(g["in/dest_num"], torch.int64)[None]
        sc["obstacles"] = cu(g["in/obstacles"])
        sc["desired_speed"] = cu(g["in/desired_speed"])[None]
        for k in ("ped_features0", "obs_features0", "self_features0"):
            sc[k] = cu(g["in/" + k])[None]
        args = bm_args()
        args.time_unit = float(g["in/time_unit"])
        if str(g["in/model"]) == "sfm":
            spec, packed, packed_tc = P.SocialForce("gc1560").spec, None, None
            res = rollout_scenes(spec, packed, args, sc, t0, t0 + frames)
            assert torch.isfinite(res[3]).all()
        else:
            torch.manual_seed(666)
            net = M.PINNSF_bottleneck_multitask(args).cuda().eval()
            sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd/")}
            net.load_state_dict(sd)
            packed = M.pack_device(net.state_dict(), net.spec)
            for tc in (M.pack_device_tc(net.state_dict(), net.spec), None):
                res = rollout_scenes(net.spec, packed, args, sc, t0, t0 + frames, packed_tc=tc)
                assert torch.isfinite(res[3]).all()
    torch.cuda.synchronize()


def part_train():
    """Training forward / backward of both network kinds, feature + Euler backward, fused losses."""
    from piml_b200.autograd import BceSumFunction, L1SumFunction, RolloutLossesFunction
    rng = np.random.default_rng(0)
    N = 300
    for kind, cls in (("pinnsf_bm", M.PINNSF_bottleneck_multitask), ("pinnsf_m", M.PINNSF_multitask)):
        torch.manual_seed(1)
        net = cls(bm_args(kind)).cuda().train()
        ped = cu(rng.normal(0, 1, (N, 6, 6))).requires_grad_(True)
        obs = cu(rng.normal(0, 1, (N, 10, 6)))
        slf = cu(rng.normal(0, 1, (N, 7)))
        out = net(ped, obs, slf)
        (out[0].square().sum() + out[1].abs().sum() + out[-1].sum()).backward()
    p = cu(rng.random((4, 3, 50, 2)) * 8).requires_grad_(True)
    v, a = cu(rng.normal(0, 1, (4, 3, 50, 2))), cu(rng.normal(0, 1, (4, 3, 50, 2)))
    d = cu(rng.random((4, 3, 50, 2)) * 8)
    obs = cu(rng.random((20, 2)) * 8)
    f = P.Pedestrians().get_relative_features(p, v, a, d, obs, 6, 90, 4, 10, 90, 4)
    (f[0].sum() + f[1].sum() + f[2].sum()).backward()
    pred = cu(rng.normal(0, 1, (4, 5, 50, 2))).requires_grad_(True)
    lab = cu(rng.normal(0, 1, (4, 5, 50, 12)))
    coll = (cu(rng.random((4, 5, 50))) < 0.1).float()
    RolloutLossesFunction.apply(pred, lab[..., :2], 0.9, False, coll, coll, None).sum().backward()
    x = cu(rng.normal(0, 1, (4, 5, 50, 2))).requires_grad_(True)
    L1SumFunction.apply(x, 1e-3).backward()
    pr = cu(rng.random((4, 5, 50, 6))).requires_grad_(True)
    BceSumFunction.apply(pr, (cu(rng.random((4, 5, 50, 6))) < 0.3).float())[0].backward()
    P.Pedestrians.collision_detection(cu(rng.random((4, 5, 50, 2)) * 6), 0.5)
    P.Pedestrians.collision_detection(cu(rng.random((30, 50, 2)) * 6), 0.5, rowsum_only=True)
    torch.cuda.synchronize()


def part_misc():
    """Metrics, heading with time fill, selection helpers, calc_acceleration v0-v2, dense helpers, desired speed,
    large-k selection, the IEEE MLAPM kernel and row ranges."""
    from piml_b200 import metrics as MT
    from piml_b200.dataset import desired_speed
    rng = np.random.default_rng(2)
    T, N = 12, 70
    p = cu(rng.normal(0, 4, (T, N, 2)))
    q = p + cu(rng.normal(0, 0.2, (T, N, 2)))
    mask = cu(rng.random((T, N)) < 0.7, torch.int64)
    MT.mae_with_time_mask(p, q, mask, reduction='sum')
    MT.ot_with_time_mask(p, q, mask, reduction='sum')
    MT.mmd_with_time_mask(p, q, mask, reduction='sum')
    v = cu(rng.normal(0, 1, (T, N, 2)))
    v[3:6, ::4] = 0
    ped = P.Pedestrians()
    head = ped.get_heading_direction(v)
    ped.get_nearby_obj_in_sight(p, p, head, 20, 120)
    desired_speed(v, 5)
    rq = ped.get_relative_quantity(p, q)
    dist, idx = ped.get_nearby_obj_in_sight(p, q, head, 4, 90)
    ped.get_filtered_features(rq, idx, dist, 4.0)
    for ver, dsn in (("v0", "gc1560"), ("v1", "ucy"), ("v2", "gc2344")):
        P.calc_acceleration(cu(rng.normal(0, 1, (40, 6, 6))), ver, dsn)
    kw = dict(version='GC', tau=0.5, A=7.55, B=-3.0, C=0.2, D=-0.3, theta=56)
    for n in (1, 37, 700, 1300):
        pp, vv = cu(rng.random((n, 2)) * 30), cu(rng.normal(0, 1, (n, 2)))
        dd, ds = cu(rng.random((n, 2)) * 30), cu(np.full((n, 1), 1.3))
        P.MLAPM(**kw).advance(pp, vv, ds, dd, 0.08, 0.3)
        P.MLAPM(**dict(kw, exact_math=True)).step(pp, vv, ds, dd, 0.08)
        P.MLAPM(**dict(kw, version='raw')).step(pp, vv, ds, dd, 0.08, rows=(n // 3, n))
    torch.cuda.synchronize()


def part_nnstep():
    """The fused NN step (cell-list build, sorted-order feature kernel with its shared-memory pools, tensor-core forward
    on compact rows, finish + integrate), the cell-list feature call (dense outputs) and the fused loop inside
    piml_rollout_f32, on small crowds with absent / stationary agents and several scenes."""
    from piml_b200.rollout import NNStep, rollout_scenes, state_features
    args = bm_args()
    torch.manual_seed(666)
    net = M.PINNSF_bottleneck_multitask(args).cuda().eval()
    packed_tc = M.pack_device_tc(net.state_dict(), net.spec)
    g = torch.Generator().manual_seed(3)
    P._lib.check(P._lib.load().piml_set_feature_algorithm(2), "algo")
    for S, N, Mo in ((1, 1500, 200), (3, 200, 60)):
        side = (N / 0.5) ** 0.5
        p = (torch.rand(S, N, 2, generator=g) * side).cuda()
        v = (torch.randn(S, N, 2, generator=g)).cuda()
        a = (0.3 * torch.randn(S, N, 2, generator=g)).cuda()
        p[:, ::7] = float('nan'); v[:, 5::11] = 0.0
        dest = (torch.rand(S, N, 2, generator=g) * side).cuda()
        obs = (torch.rand(S, Mo, 2, generator=g) * side).cuda() if S > 1 else (torch.rand(Mo, 2, generator=g) * side).cuda()
        ds = torch.full((S, N), 1.3).cuda()
        hist = torch.where(torch.isnan(v), torch.zeros_like(v), v).contiguous()
        didx = torch.zeros(S, N, dtype=torch.int64).cuda()
        dnum = torch.ones(S, N, dtype=torch.int64).cuda()
        wp = dest[:, None].contiguous()
        state_features(p, v.clone(), a.clone(), dest, obs, hist, ds, 6, 90, 4, 10, 90, 4)
        step = NNStep(net.spec, packed_tc, p, v, a, dest, didx, hist, dnum, wp, ds, obs, 0.08, 6, 90, 4, 10, 90, 4)
        for _ in range(3):
            step.step()
        torch.cuda.synchronize()
        assert torch.isfinite(p).any()
    P._lib.check(P._lib.load().piml_set_feature_algorithm(0), "algo")
    os.environ["PIML_ROLLOUT_FUSED"] = "1"
    try:
        part_rollout()
    finally:
        os.environ.pop("PIML_ROLLOUT_FUSED", None)


PARTS = {"smoke": part_smoke, "rollout": part_rollout, "train": part_train, "misc": part_misc, "nnstep": part_nnstep}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--part", default="all")
    a = ap.parse_args()
    for name, fn in PARTS.items():
        if a.part in ("all", name):
            fn()
            print(f"part {name}: ok", flush=True)
