"""One launch of every hot kernel at the BASELINE shapes, for ncu (never a bench number).

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o gpurun_out/prof \
        python scripts/profile_workloads.py [--agents 100000]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import piml_b200 as P  # noqa: E402
from piml_b200 import _lib as L  # noqa: E402
from piml_b200 import models as M  # noqa: E402
from scripts.bench_stages import bm_args  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=100000)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda")
    N = a.agents
    p, v, ds, dest, obs = [x.to(dev) for x in bench.synthetic_crowd(N)]
    acc = torch.zeros_like(v)
    ped = P.Pedestrians()
    fargs = (6, 90, 4, 10, 90, 4)
    model = P.MLAPM(**bench.MLAPM_KW)
    torch.manual_seed(666)
    net = M.PINNSF_bottleneck_multitask(bm_args()).to(dev).train()
    packed = M.pack_device(net.state_dict(), net.spec, dev)
    packed_tc = M.pack_device_tc(net.state_dict(), net.spec, dev)
    sfm = P.SocialForce("gc1560")
    for _ in range(a.reps):
        model.advance(p, v, ds, dest, bench.DT, bench.RADIUS)                       # mlapm_sym_kernel
        for algo in (1, 2):                                                         # relative_features / cells
            L.check(L.load().piml_set_feature_algorithm(algo), "algo")
            feats = ped.get_relative_features(p[None], v[None], acc[None], dest[None], obs, *fargs)
        L.check(L.load().piml_set_feature_algorithm(0), "algo")
        slf = torch.cat([feats[2][0], v, acc, ds], -1)
        with torch.no_grad():
            net.eval()
            net(feats[0][0], feats[1][0], slf)                                      # pinnsf_tile_kernel (inference)
            M.pinnsf_forward(net.spec, packed, feats[0][0], feats[1][0], slf, need_msgs=False,
                             packed_tc=packed_tc)                                   # pinnsf_tc_kernel (compact mode)
            sfm(feats[0][0], feats[1][0], slf)                                      # sfm_forward_kernel
        net.train()
        out = net(feats[0][0], feats[1][0], slf)                                    # pinnsf_tile_kernel (+ stash)
        (out[0].sum() + out[1].sum() + out[3].sum()).backward()                     # pinnsf_bwd_tile / dw kernels
        net.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    print("profile workloads done")


if __name__ == "__main__":
    main()
