import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piml_b200 import models as M
from scripts.bench_stages import bm_args, timeit
dev = torch.device("cuda")
a = bm_args(); torch.manual_seed(666)
net = M.PINNSF_bottleneck_multitask(a).to(dev).eval()
N = int(os.environ.get("NAG", "100000"))
g = torch.Generator().manual_seed(1)
ped = torch.randn(N, 6, 6, generator=g).to(dev); obs = torch.randn(N, 10, 6, generator=g).to(dev); slf = torch.randn(N, 7, generator=g).to(dev)
packed = M.pack_device(net.state_dict(), net.spec, dev); ptc = M.pack_device_tc(net.state_dict(), net.spec, dev)
ttc = timeit(lambda: M.pinnsf_forward(net.spec, packed, ped, obs, slf, need_msgs=False, packed_tc=ptc))
print(f"dbg={os.environ.get('PIML_TC_DEBUG','0')} N={N}: tcgen05 forward {ttc:.3f} ms")
