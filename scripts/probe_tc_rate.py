import os, sys, subprocess
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piml_b200 import _lib as L
dev = torch.device("cuda")
K, N = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn(128, K).to(dev); w = torch.randn(N, K).to(dev)
y = torch.zeros(128, N, device=dev)
L.check(L.load().piml_tc_selftest_f32(L.ptr(x), L.ptr(w), K, N, 3, L.ptr(y), L.stream_ptr(dev)), "tc")
torch.cuda.synchronize()
print(f"K={K} N={N} variant={os.environ.get('PIML_TC_SBO')}: issue {float(y[0,0]):.1f} cyc/MMA, issue+drain {float(y[0,1]):.1f} cyc/MMA over {int(y[0,2])} MMAs")
