"""torchrun --nproc-per-node G scripts/check_push_exchange.py : the fused peer-memory exchange (ShardedCrowd) must give
bit-identical crowd state to the NCCL all-gather path, on every rank, and both must equal the unsharded step."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import piml_b200 as P
from piml_b200.sharded import ShardedCrowd

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world, rank = dist.get_world_size(), dist.get_rank()
N, steps = 8192 * world, 6
p, v, ds, dest, _ = [x.to(dev) for x in bench.synthetic_crowd(N)]
model = P.MLAPM(**bench.MLAPM_KW)
# reference: unsharded
pu, vu = p.clone(), v.clone()
for _ in range(steps):
    act, pn, arr = model.advance(pu, vu, ds, dest, bench.DT, bench.RADIUS)
    pu, vu = pn, act
# NCCL all-gather path
shard = N // world
r0, r1 = rank * shard, (rank + 1) * shard
pa, va = p.clone(), v.clone()
for _ in range(steps):
    act, pn, arr = model.advance(pa, va, ds, dest, bench.DT, bench.RADIUS, rows=(r0, r1))
    pn2, vn2 = torch.empty_like(pa), torch.empty_like(va)
    dist.all_gather_into_tensor(pn2, pn); dist.all_gather_into_tensor(vn2, act)
    pa, va = pn2, vn2
# fused push path
crowd = ShardedCrowd(N, device=dev)
crowd.load(p, v)
for _ in range(steps):
    crowd.step(model, ds, dest, bench.DT, bench.RADIUS)
torch.cuda.synchronize()
ok = (torch.equal(crowd.position, pa) and torch.equal(crowd.velocity, va) and torch.equal(pa, pu) and torch.equal(va, vu))
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"push exchange == NCCL all-gather == unsharded on all {world} ranks after {steps} steps: {bool(int(t))}"
          f"  (peer buffers: {len(crowd.ptrs)})")
dist.destroy_process_group()
sys.exit(0 if int(t) else 1)
