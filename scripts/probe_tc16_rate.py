"""tcgen05 kind::f16 MMA rate (A from TMEM, B from shared memory, no swizzle): cycles per MMA for several layer shapes
and issue patterns (one accumulator / two accumulators alternating per chain / + commits / alternating per K step)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piml_b200 import _lib as L
dev = torch.device("cuda")
NAMES = {0: "one accumulator", 1: "D0/D1 per chain", 2: "D0/D1 per chain + commit", 3: "D0/D1 per K step", 4: "D0/D1 per chain, first MMA overwrites"}
for (K, N) in ((128, 128), (128, 64), (64, 64), (16, 128)):
    for variant in (0, 1, 2, 3, 4):
        for terms in (3, 1):
            x = torch.randn(128, K, device=dev); w = torch.randn(N, K, device=dev)
            y = torch.zeros(128, N, device=dev)
            L.check(L.load().piml_tc16_selftest_f32(L.ptr(x), L.ptr(w), K, N, terms, -(variant * 1000 + 64), L.ptr(y), L.stream_ptr(dev)), "tc16")
            torch.cuda.synchronize()
            r = y.flatten()[:3].tolist()
            print(f"f16 K={K:3d} N={N:3d} terms={terms} {NAMES[variant]:38s}: issue {r[0]:6.1f} cycles/MMA, issue+drain {r[1]:6.1f}, {int(r[2])} MMAs")
