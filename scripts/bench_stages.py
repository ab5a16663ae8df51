"""Per-stage device timings of the hot path at the BASELINE shapes (not the headline bench; evidence for DESIGN.md).

    python scripts/bench_stages.py [--agents 100000] [--scenes 4096]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import piml_b200 as P  # noqa: E402
from piml_b200 import models as M  # noqa: E402
from piml_b200.rollout import integrate_step, state_features  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bm_args():
    return argparse.Namespace(model='pinnsf_bm', dataset_name='gc1560', dropout=0.5, encoder_hidden_size=128,
                              processor_hidden_size=128, decoder_hidden_size=64, encoder_hidden_layers=3,
                              processor_hidden_layers=16, decoder_hidden_layers=2, ped_feature_dim=6,
                              obs_feature_dim=6, self_feature_dim=7)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=100000)
    ap.add_argument("--scenes", type=int, default=4096)
    a = ap.parse_args()
    dev = torch.device("cuda")
    out = {}
    N = a.agents
    p, v, ds, dest, obs = [x.to(dev) for x in bench.synthetic_crowd(N)]
    acc = torch.zeros_like(v)
    ped = P.Pedestrians()
    fargs = (6, 90, 4, 10, 90, 4)
    feats = ped.get_relative_features(p[None], v[None], acc[None], dest[None], obs, *fargs)
    from piml_b200 import _lib as L
    for algo, name in ((1, "allpairs"), (2, "cells")):
        L.check(L.load().piml_set_feature_algorithm(algo), "piml_set_feature_algorithm")
        ms = timeit(lambda: ped.get_relative_features(p[None], v[None], acc[None], dest[None], obs, *fargs))
        out["features_%s_N%d_M%d" % (name, N, obs.shape[0])] = {
            "ms": ms, "agent_steps_per_s": N / ms * 1e3, "pairs_per_s": N * (N + obs.shape[0]) / ms * 1e3}
    L.check(L.load().piml_set_feature_algorithm(0), "piml_set_feature_algorithm")
    torch.manual_seed(666)
    net = M.PINNSF_bottleneck_multitask(bm_args()).to(dev).eval()
    packed = M.pack_device(net.state_dict(), net.spec, dev)
    packed_tc = M.pack_device_tc(net.state_dict(), net.spec, dev)
    slf = torch.cat([feats[2][0], v, acc, ds], -1)
    ms = timeit(lambda: M.pinnsf_forward(net.spec, packed, feats[0][0], feats[1][0], slf, need_msgs=False))
    out["pinnsf_bm_forward_fp32pipe_N%d" % N] = {"ms": ms, "agent_steps_per_s": N / ms * 1e3,
                                                 "tflops": 1.52e6 * N / ms * 1e3 / 1e12}
    ms = timeit(lambda: M.pinnsf_forward(net.spec, packed, feats[0][0], feats[1][0], slf, need_msgs=False,
                                         packed_tc=packed_tc))
    out["pinnsf_bm_forward_tcgen05_N%d" % N] = {"ms": ms, "agent_steps_per_s": N / ms * 1e3,
                                                "tflops_algorithmic": 1.52e6 * N / ms * 1e3 / 1e12}
    # GC-shaped scenes, one NN rollout step (forward + integrate + features) for S scenes of 122 slots
    S, Ns, Mo = a.scenes, 122, 100
    g = torch.Generator().manual_seed(1)
    pos = (torch.rand(S, Ns, 2, generator=g) * 20).to(dev)
    pos[:, 30:] = float('nan')                                  # ~25% of the slots active, like the GC clips
    vel = torch.randn(S, Ns, 2, generator=g).to(dev)
    ac = torch.zeros(S, Ns, 2, device=dev)
    dst = (torch.rand(S, Ns, 2, generator=g) * 20).to(dev)
    ob = (torch.rand(Mo, 2, generator=g) * 20).to(dev)
    dsp = torch.full((S, Ns), 1.3, device=dev)
    hist = vel.clone()
    didx = torch.zeros(S, Ns, dtype=torch.int64, device=dev)
    dnum = torch.ones(S, Ns, dtype=torch.int64, device=dev)
    wp = dst[:, None].contiguous()
    pf, of, sf = state_features(pos, vel, ac, dst, ob, hist, dsp, *fargs)
    bufs = (pf, of, sf, torch.empty(S, Ns, 2, device=dev))

    def fwd():
        return M.pinnsf_forward(net.spec, packed, pf.view(S * Ns, 6, 6), of.view(S * Ns, 10, 6), sf.view(S * Ns, 7),
                                need_msgs=False, packed_tc=packed_tc)[0].view(S, Ns, 2)

    def nn_step():
        a_next = fwd()
        integrate_step(pos, vel, ac, a_next, dst, didx, dnum, wp, 0.0, False, hist_v=hist)   # dt=0: state stays put
        state_features(pos, vel, ac, dst, ob, hist, dsp, *fargs, out=bufs)
    key = "nn_rollout_step_S%d_N%d" % (S, Ns)
    ms = timeit(nn_step)
    out[key] = {"ms": ms, "agent_steps_per_s": S * Ns / ms * 1e3}
    out[key]["forward_ms"] = timeit(fwd)
    out[key]["features_ms"] = timeit(lambda: state_features(pos, vel, ac, dst, ob, hist, dsp, *fargs, out=bufs))
    # single GC scene: latency of one step (launch-bound regime)
    S1 = 1
    one = [x[:S1].contiguous() for x in (pos, vel, ac, dst, hist, dsp, didx, dnum, wp)]
    pf1, of1, sf1 = state_features(one[0], one[1], one[2], one[3], ob, one[4], one[5], *fargs)

    def one_step():
        a_next = M.pinnsf_forward(net.spec, packed, pf1[0], of1[0], sf1[0], need_msgs=False,
                                  packed_tc=packed_tc)[0][None]
        integrate_step(one[0], one[1], one[2], a_next, one[3], one[6], one[7], one[8], 0.0, False, hist_v=one[4])
        state_features(one[0], one[1], one[2], one[3], ob, one[4], one[5], *fargs)
    ms = timeit(one_step, iters=50)
    out["nn_rollout_step_S1_N122"] = {"ms": ms, "agent_steps_per_s": Ns / ms * 1e3}
    # whole rollouts through the C-side loop (piml_rollout_f32): S scenes x 122 slots x 300 frames
    from piml_b200.rollout import rollout_scenes
    rargs = argparse.Namespace(time_unit=0.08, topk_ped=6, sight_angle_ped=90, dist_threshold_ped=4, topk_obs=10,
                               sight_angle_obs=90, dist_threshold_obs=4)
    for S2 in (1, 64):
        T2 = 300
        g = torch.Generator().manual_seed(3)
        P0 = (torch.rand(S2, T2, Ns, 2, generator=g) * 20).to(dev)
        P0[:, :, 30:] = float('nan')
        scene = {"position": P0, "velocity": torch.randn(S2, T2, Ns, 2, generator=g).to(dev) * 0.5,
                 "acceleration": torch.zeros(S2, T2, Ns, 2, device=dev),
                 "destination": (torch.rand(S2, T2, Ns, 2, generator=g) * 20).to(dev),
                 "dest_idx": torch.zeros(S2, T2, Ns, dtype=torch.int64, device=dev),
                 "waypoints": (torch.rand(S2, 1, Ns, 2, generator=g) * 20).to(dev),
                 "dest_num": torch.ones(S2, Ns, dtype=torch.int64, device=dev), "obstacles": ob,
                 "mask_p": torch.ones(S2, T2, Ns, device=dev), "mask_p_pred": torch.ones(S2, T2, Ns, device=dev),
                 "desired_speed": torch.full((S2, Ns), 1.3, device=dev)}
        f0 = state_features(P0[:, 0].contiguous(), scene["velocity"][:, 0].contiguous(),
                            scene["acceleration"][:, 0].contiguous(), scene["destination"][:, 0].contiguous(), ob,
                            scene["velocity"][:, 0].contiguous(), scene["desired_speed"], *fargs)
        scene["ped_features0"], scene["obs_features0"], scene["self_features0"] = f0
        for tc in (packed_tc, None):
            ms = timeit(lambda: rollout_scenes(net.spec, packed, rargs, scene, 0, T2, packed_tc=tc), iters=3, warm=1)
            out["rollout_S%d_N%d_T%d_%s" % (S2, Ns, T2, "tcgen05" if tc is not None else "fp32pipe")] = {
                "ms_per_step": ms / T2, "agent_steps_per_s": S2 * Ns * T2 / ms * 1e3}
        if S2 in (1, 64):                                   # BASELINE config 2: the pure social-force model
            sf_spec = P.SocialForce("gc1560").spec
            ms = timeit(lambda: rollout_scenes(sf_spec, None, rargs, scene, 0, T2), iters=3, warm=1)
            out["rollout_S%d_N%d_T%d_socialforce" % (S2, Ns, T2)] = {
                "ms_per_step": ms / T2, "agent_steps_per_s": S2 * Ns * T2 / ms * 1e3}
    # f-1: the feature build of make_dataset over a whole GC-sized clip (data.py:766-771; 7.2 s in the reference)
    T3 = 750
    g = torch.Generator().manual_seed(5)
    Pc = (torch.rand(T3, Ns, 2, generator=g) * 20).to(dev)
    Pc[:, 30:] = float('nan')
    Vc, Ac = torch.randn(T3, Ns, 2, generator=g).to(dev), torch.zeros(T3, Ns, 2, device=dev)
    Dc = (torch.rand(T3, Ns, 2, generator=g) * 20).to(dev)
    ms = timeit(lambda: ped.get_relative_features(Pc, Vc, Ac, Dc, ob, *fargs), iters=5)
    out["make_dataset_features_T%d_N%d_M%d" % (T3, Ns, Mo)] = {"ms": ms, "agent_steps_per_s": T3 * Ns / ms * 1e3}
    # training: forward + backward of one batch (single-step training, simulators.py:327-360; batch 128 as main.py,
    # and the 32 x 144 agent rows of one UCY rollout-training step)
    for kind, cls in (("pinnsf_m", M.PINNSF_multitask), ("pinnsf_bm", M.PINNSF_bottleneck_multitask)):
        targs = bm_args()
        targs.model = kind
        torch.manual_seed(666)
        tnet = cls(targs).to(dev).train()
        for Bt in (128, 4608):
            g = torch.Generator().manual_seed(Bt)
            tp = torch.randn(Bt, 6, 6, generator=g).to(dev)
            to = torch.randn(Bt, 10, 6, generator=g).to(dev)
            ts = torch.randn(Bt, 7, generator=g).to(dev)

            def train_step():
                tnet.zero_grad(set_to_none=True)
                res = tnet(tp, to, ts)
                (res[0].square().sum() + res[-1].sum()).backward()
            ms = timeit(train_step, iters=10)
            out["train_fwd_bwd_%s_B%d" % (kind, Bt)] = {"ms": ms, "rows_per_s": Bt / ms * 1e3}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
