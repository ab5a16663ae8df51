"""Times the fused NN step (piml_nn_step_f32) against the three-call route on the bench crowd (N = 100k, M = 2000),
L2 flushed between steps.  Usage: python scripts/time_nn_step.py [N] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    from piml_b200 import _lib as L
    dev = torch.device("cuda", 0)
    _, _, _, _, obs_h = bench.synthetic_crowd(N)
    flush = torch.empty(bench.FLUSH_MB << 20, dtype=torch.uint8, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for name in ("three_calls", "fused"):
        crowd = bench.NNCrowd(torch, dev, N, obs_h)
        fn = crowd.step if name == "three_calls" else crowd.step_fused
        for _ in range(3):
            flush.zero_(); fn()
        torch.cuda.synchronize()
        l0 = L.launch_count()
        e0, e1 = ev(), ev()
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        print(f"{name}: {tot / steps:.4f} ms per step, {(L.launch_count() - l0) / steps:.1f} launches per step", flush=True)


if __name__ == "__main__":
    main()
