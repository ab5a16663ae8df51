"""Where does a small-batch training step (forward + backward, B = 128) spend its time?  torch.profiler table."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piml_b200 as P
from piml_b200 import models as M
from scripts.bench_stages import bm_args

dev = torch.device("cuda")
args = bm_args(); args.model = "pinnsf_m"
torch.manual_seed(666)
net = M.PINNSF_multitask(args).to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator().manual_seed(1)
tp, to, ts = torch.randn(B, 6, 6, generator=g).to(dev), torch.randn(B, 10, 6, generator=g).to(dev), torch.randn(B, 7, generator=g).to(dev)
lab = torch.randn(B, 2, generator=g).to(dev)

def step():
    opt.zero_grad(set_to_none=True)
    res = net(tp, to, ts)
    loss = torch.nn.functional.mse_loss(res[0], lab, reduction='sum')
    loss.backward()
    opt.step()

for _ in range(5): step()
torch.cuda.synchronize()
l0 = P._lib.launch_count(); t0 = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize()
print(f"B={B}: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per train step (fwd+bwd+Adam), {(P._lib.launch_count() - l0) / 50:.1f} library launches per step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(10): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
