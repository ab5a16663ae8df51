"""In-kernel cycle counters of the tensor-core forward (PIML_TC_PROF=1) at N = 100k on the bench crowd's features:
dense vs compact mode.  Prints the per-tile breakdown of CTA 0 (stderr of the library)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--no-prof" not in sys.argv:
    os.environ["PIML_TC_PROF"] = "1"
import bench
import piml_b200 as P
from piml_b200 import models as M
from scripts.bench_stages import bm_args

dev = torch.device("cuda")
N = 100000
p, v, ds, dest, obs = [x.to(dev) for x in bench.synthetic_crowd(N)]
acc = torch.zeros_like(v)
feats = P.Pedestrians().get_relative_features(p[None], v[None], acc[None], dest[None], obs, 6, 90, 4, 10, 90, 4)
slf = torch.cat([feats[2][0], v, acc, ds], -1)
torch.manual_seed(666)
net = M.PINNSF_bottleneck_multitask(bm_args()).to(dev).eval()
packed = M.pack_device(net.state_dict(), net.spec, dev)
ptc = M.pack_device_tc(net.state_dict(), net.spec, dev)
for _ in range(3):
    M.pinnsf_forward(net.spec, packed, feats[0][0], feats[1][0], slf, need_msgs=False, packed_tc=ptc)
torch.cuda.synchronize()
nz_p = int((feats[0][0].abs().sum(-1) > 0).sum()); nz_o = int((feats[1][0].abs().sum(-1) > 0).sum())
print(f"non-zero slot rows: ped {nz_p} of {N * 6}, obs {nz_o} of {N * 10}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
os.environ.pop("PIML_TC_PROF", None)
for _ in range(2):
    M.pinnsf_forward(net.spec, packed, feats[0][0], feats[1][0], slf, need_msgs=False, packed_tc=ptc)
e0.record()
for _ in range(10):
    M.pinnsf_forward(net.spec, packed, feats[0][0], feats[1][0], slf, need_msgs=False, packed_tc=ptc)
e1.record()
torch.cuda.synchronize()
print(f"forward (compact + tc16 + finish), no profiling: {e0.elapsed_time(e1) / 10:.4f} ms")
