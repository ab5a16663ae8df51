"""tcgen05 kind::f16 MMA rate (A from TMEM, B from shared memory, no swizzle): cycles per MMA for N = 128 / 64 / 32."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piml_b200 import _lib as L
dev = torch.device("cuda")
for (K, N) in ((128, 128), (128, 64), (64, 64), (128, 32), (16, 128)):
    for terms in (3, 1):
        x = torch.randn(128, K, device=dev); w = torch.randn(N, K, device=dev)
        y = torch.zeros(128, N, device=dev)
        L.check(L.load().piml_tc16_selftest_f32(L.ptr(x), L.ptr(w), K, N, terms, -64, L.ptr(y), L.stream_ptr(dev)), "tc16")
        torch.cuda.synchronize()
        r = y.flatten()[:3].tolist()
        print(f"f16  K={K:3d} N={N:3d} terms={terms}: issue {r[0]:.1f} cycles/MMA, issue+drain {r[1]:.1f}, {int(r[2])} MMAs")
