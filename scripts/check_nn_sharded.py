"""torchrun --nproc-per-node G scripts/check_nn_sharded.py : the agent-sharded NN rollout step (ShardedNNCrowd; pinnsf_bm:
the fused step with the peer-memory state push, social force: own-row features + forward, all-gather of the
accelerations, replicated state update) must be bit-identical to the unsharded step sequence on every rank; also times
both."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import piml_b200 as P
from piml_b200 import models as M
from piml_b200.rollout import integrate_step, state_features
from piml_b200.sharded import ShardedNNCrowd
from scripts.bench_stages import bm_args

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world, rank = dist.get_world_size(), dist.get_rank()
N = int(os.environ.get("PIML_CHECK_AGENTS", 100000)) // world * world
steps = 8
p, v, ds, dest, obs = [x.to(dev) for x in bench.synthetic_crowd(N)]
args = bm_args()
args.topk_ped, args.topk_obs, args.sight_angle_ped, args.sight_angle_obs = 6, 10, 90, 90
args.dist_threshold_ped, args.dist_threshold_obs, args.time_unit = 4, 4, 0.08
ok_all = True
for label, net in (("pinnsf_bm", None), ("social force", P.SocialForce("gc1560"))):
    if net is None:
        torch.manual_seed(666)
        net = M.PINNSF_bottleneck_multitask(args).to(dev).eval()
    # unsharded reference sequence (same kernels, all rows)
    pu, vu, au, du = p[None].clone(), v[None].clone(), torch.zeros(1, N, 2, device=dev), dest[None].clone()
    hist, dsp = vu.clone(), ds.reshape(1, N).contiguous()
    didx, dnum = torch.zeros(1, N, dtype=torch.int64, device=dev), torch.ones(1, N, dtype=torch.int64, device=dev)
    wp = du[:, None].contiguous()
    fargs = (6, 90, 4, 10, 90, 4)
    P._lib.check(P._lib.load().piml_set_feature_algorithm(2), "algo")
    if not isinstance(net, P.SocialForce):
        packed = M.pack_device(net.state_dict(), net.spec, dev)
        packed_tc = M.pack_device_tc(net.state_dict(), net.spec, dev)
    with torch.no_grad():
        pf, of, sf = state_features(pu, vu, au, du, obs, hist, dsp, *fargs)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            a_next = net(pf[0], of[0], sf[0])[0].view(1, N, 2) if isinstance(net, P.SocialForce) else \
                M.pinnsf_forward(net.spec, packed, pf[0], of[0], sf[0], need_msgs=False,
                                 packed_tc=packed_tc)[0].view(1, N, 2)
            integrate_step(pu, vu, au, a_next, du, didx, dnum, wp, 0.08, True, hist_v=hist)
            pf, of, sf = state_features(pu, vu, au, du, obs, hist, dsp, *fargs)
        t1.record(); torch.cuda.synchronize()
        ms_u = t0.elapsed_time(t1) / steps
        crowd = ShardedNNCrowd(net, args, N, obs, device=dev)
        zero = torch.zeros(N, 2, device=dev)
        crowd.load(p, v, zero, dest, torch.zeros(N, dtype=torch.int64), torch.ones(N, dtype=torch.int64), dest[None], ds)
        dist.barrier(); torch.cuda.synchronize()
        t0.record()
        for _ in range(steps):
            crowd.step()
        t1.record(); torch.cuda.synchronize()
        ms_s = t0.elapsed_time(t1) / steps
    P._lib.check(P._lib.load().piml_set_feature_algorithm(0), "algo")
    crowd.gather_state()
    eq = lambda x, y: bool(((x == y) | (x.isnan() & y.isnan())).all())       # arrived agents are NaN in both
    same = eq(crowd.p, pu) and eq(crowd.v, vu) and eq(crowd.a, au) and eq(crowd.dest, du)
    arrived = int(pu.isnan().any(-1).sum())
    t = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok_all = ok_all and bool(int(t))
    if rank == 0:
        print(f"{label}: agent-sharded NN step on {world} ranks == unsharded after {steps} steps: {bool(int(t))};  "
              f"N={N}: sharded {ms_s:.3f} ms/step "
              f"({N / ms_s * 1e3 / 1e6:.1f} M agent-steps/s); {arrived} agents arrived")
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
