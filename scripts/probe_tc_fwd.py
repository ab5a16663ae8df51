"""Bring-up check of the tensor-core forward vs the FP32-pipe kernel and timing at N=100k."""
import argparse, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piml_b200 as P
from piml_b200 import models as M
from scripts.bench_stages import bm_args, timeit

dev = torch.device("cuda")
def check(kind, R, kp, ko, chan=0, has_obs=True):
    a = bm_args(); a.model = kind
    if not has_obs: a.obs_feature_dim = 0
    torch.manual_seed(666)
    net = M.CLASSES[kind](a).to(dev).eval()
    g = torch.Generator().manual_seed(R)
    lead = (chan, R // chan) if chan else (R,)
    ped = torch.randn(*lead, kp, 6, generator=g).to(dev); ped[..., -1, :] = 0
    obs = torch.randn(*lead, max(ko, 1), 6, generator=g)[..., :ko, :].to(dev)
    slf = torch.randn(*lead, 7, generator=g).to(dev)
    packed = M.pack_device(net.state_dict(), net.spec, dev)
    ptc = M.pack_device_tc(net.state_dict(), net.spec, dev)
    assert ptc is not None
    need = net.spec.kind == 0 and not net.spec.coll_dims
    ref = M.pinnsf_forward(net.spec, packed, ped, obs, slf, need_msgs=need)
    got = M.pinnsf_forward(net.spec, packed, ped, obs, slf, need_msgs=need, packed_tc=ptc)
    torch.cuda.synchronize()
    e = (got[0] - ref[0]).norm(dim=-1) / ref[0].norm(dim=-1).clamp_min(1e-3)
    msg = f"{kind:18s} R={R:6d} kp={kp} ko={ko} chan={chan}: acc max rel err {float(e.max()):.2e}"
    if need:
        msg += f"  msgs max abs {float((got[1]-ref[1]).abs().max()):.2e} / {float(ref[1].abs().max()):.2e}"
    print(msg, flush=True)

check("pinnsf_bm", 21, 6, 10)
check("pinnsf_bm", 300, 6, 10)
check("pinnsf_bottleneck", 1000, 6, 10)
check("pinnsf_bottleneck", 999, 5, 3)
check("pinnsf_m", 257, 6, 10)
check("pinnsf", 640, 6, 0, has_obs=False)
check("pinnsf_bm", 640, 6, 2, chan=5)
check("pinnsf_bm", 100000, 6, 10)
# timing
a = bm_args(); torch.manual_seed(666)
net = M.PINNSF_bottleneck_multitask(a).to(dev).eval()
N = 100000
g = torch.Generator().manual_seed(1)
ped = torch.randn(N, 6, 6, generator=g).to(dev); obs = torch.randn(N, 10, 6, generator=g).to(dev); slf = torch.randn(N, 7, generator=g).to(dev)
packed = M.pack_device(net.state_dict(), net.spec, dev); ptc = M.pack_device_tc(net.state_dict(), net.spec, dev)
t32 = timeit(lambda: M.pinnsf_forward(net.spec, packed, ped, obs, slf, need_msgs=False))
ttc = timeit(lambda: M.pinnsf_forward(net.spec, packed, ped, obs, slf, need_msgs=False, packed_tc=ptc))
print(f"N={N}: FP32-pipe {t32:.3f} ms ({1.52e6*N/t32*1e3/1e12:.1f} TFLOP/s)   tcgen05 3xTF32 {ttc:.3f} ms ({1.52e6*N/ttc*1e3/1e12:.1f} TFLOP/s algorithmic)")
