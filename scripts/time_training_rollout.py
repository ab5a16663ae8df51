"""Timing of one rollout-training step (BASELINE config 3: UCY clip, pinnsf_bm, channelled windows): the golden UCY
batch (6 channels x 5 steps x 144 slots) tiled to C = 32 channels, test_multiple_rollouts_for_training + loss.backward()
+ Adam step, all kernels from libpiml_b200.so.  (19.2 s forward + 0.48 s backward per
C = 32, T = 10 batch in the reference on CPU.)"""
import argparse, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piml_b200 as P
from piml_b200 import train_rollout as TRO
from tests.golden_args import base_args
from tests.test_gpu_training import _batch_from_golden, mirror
from tests.util import golden, group

g = group(golden("training_rollout"), "ucy_bm")
kind, dsn = str(g["in/model"]), str(g["in/dataset_name"])
a = g["in/args"]
args = base_args(model=kind, dataset_name=dsn, reg_weight=float(a[0]), collision_threshold=float(a[1]),
                 collision_loss_weight=float(a[2]), hard_collision_penalty=float(a[3]), teacher_weight=float(a[4]),
                 collision_pred_weight=float(a[5]), collision_focus_weight=float(a[6]),
                 new_collision_loss_flag=int(a[7]), time_decay=float(a[8]),
                 collision_loss_version=str(g["in/collision_loss_version"]))
net = mirror(kind, dsn, None, True)
opt = torch.optim.Adam(net.parameters(), lr=4e-6)
sim = argparse.Namespace(args=args, model=net, collision_count=0, hard_collision_count=0, epoch=0, batch_idx=0)
base = _batch_from_golden(g)
C0 = base.position.shape[0]
rep = (32 + C0 - 1) // C0


def make_batch():
    b = type(base)()
    for k, v in base.__dict__.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == C0 and k not in ("obstacles", "dest_num"):
            v = v.repeat(rep, *([1] * (v.dim() - 1)))[:32].clone()
        elif torch.is_tensor(v):
            v = v.clone()
        setattr(b, k, v)
    return b


def step():
    opt.zero_grad(set_to_none=True)
    res = TRO.test_multiple_rollouts_for_training(sim, make_batch())
    res[0].backward()
    opt.step()
    return res


for _ in range(3):
    res = step()
torch.cuda.synchronize()
l0 = P._lib.launch_count(); t0 = time.perf_counter()
n = 10
for _ in range(n):
    step()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / n * 1e3
b = make_batch()
print(f"rollout-training step, {kind}/{dsn}, C={b.position.shape[0]} T={b.position.shape[1]} N={b.position.shape[2]}: "
      f"{ms:.2f} ms per step (forward rollout + losses + backward + Adam), {(P._lib.launch_count() - l0) / n:.0f} library "
      f"launches per step, loss {float(res[0]):.4f}")
