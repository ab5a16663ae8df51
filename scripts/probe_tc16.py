"""Bring-up check of the 16-bit tcgen05 path (kind::f16, fp16 hi/lo split): one 128 x K x N layer vs fp64."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piml_b200 import _lib as L

dev = torch.device("cuda")
for (K, N) in ((16, 16), (16, 128), (32, 64), (128, 128), (64, 64), (128, 64)):
    g = torch.Generator().manual_seed(K * 1000 + N)
    x = torch.randn(128, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5
    ref = (x.double() @ w.double().T).numpy()
    f32 = (x @ w.T).numpy()
    for swap in (0, 1):
        for terms in (1, 3):
            y = torch.full((128, N), 7.0, device=dev)
            xd, wd = x.to(dev), w.to(dev)
            L.check(L.load().piml_tc16_selftest_f32(L.ptr(xd), L.ptr(wd), K, N, terms, swap, L.ptr(y), L.stream_ptr(dev)), "tc16")
            torch.cuda.synchronize()
            got = y.cpu().numpy()
            err = np.abs(got - ref).max() / np.abs(ref).max()
            print(f"K={K:3d} N={N:3d} swap={swap} terms={terms}: max err vs fp64 {err:.3e}   (torch fp32: "
                  f"{np.abs(f32-ref).max()/np.abs(ref).max():.3e})  nan={np.isnan(got).any()}", flush=True)
# tf32 path for comparison
for (K, N) in ((128, 128),):
    g = torch.Generator().manual_seed(K * 1000 + N)
    x = torch.randn(128, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5
    ref = (x.double() @ w.double().T).numpy()
    y = torch.full((128, N), 7.0, device=dev)
    xd, wd = x.to(dev), w.to(dev)
    L.check(L.load().piml_tc_selftest_f32(L.ptr(xd), L.ptr(wd), K, N, 3, L.ptr(y), L.stream_ptr(dev)), "tc")
    torch.cuda.synchronize()
    print("3xTF32 K=128 N=128:", np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max())
