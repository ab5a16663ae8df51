"""Times the pipe probes of libpiml_b200.so (piml_pipe_probe) on the current GPU: FP32 FFMA, MUFU, packed FFMA2 and
the FFMA2+MUFU+ALU co-issue mixes.  Prints one JSON line; used to fix the roofline denominators in DESIGN.md."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from piml_b200 import _lib as L   # noqa: E402

dev = torch.device("cuda", 0)
sms = torch.cuda.get_device_properties(dev).multi_processor_count
ctas, iters = sms * 8, 4096
out = torch.empty(ctas * 256, device=dev)
names = {0: "ffma", 1: "mufu_ex2", 2: "ffma2", 3: "ffma2x8+mufu2", 4: "ffma2x8+mufu4", 5: "ffma2x8+alu4x2",
         6: "ffma2x8+mufu4+alu4x2", 7: "fma2(x,y,z) distinct", 8: "mul2(x,y) distinct", 9: "add2(x,z) distinct",
         10: "fma2(z,z,x)", 11: "scalar fma(x,y,z) distinct", 12: "fma2(y,z',x) accumulate"}
res = {"sms": sms}
for which, name in names.items():
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.load().piml_pipe_probe(which, ctas, iters, L.ptr(out), L.stream_ptr(dev)), "probe")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    units = ctas * 256 * iters * 4            # units per launch (per thread: iters*4)
    res[name] = {"ms": best, "cycles_per_unit_per_smsp_at_1.9GHz": best * 1e-3 * 1.9e9 / (units / 32 / (sms * 4))}
print(json.dumps(res, indent=1))
