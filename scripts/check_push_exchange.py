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
from piml_b200 import _lib as L
L.check(L.load().piml_set_mlapm_algorithm(1), "piml_set_mlapm_algorithm")      # ordered pairs: bit-identity checks
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
crowd = ShardedCrowd(N, device=dev, symmetric=False)
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
# symmetric (unordered-pair) evaluation, agent-sharded with the two-stage fused exchange: every rank must hold the
# SAME state bit for bit, and it must agree with the unsharded symmetric and the ordered results to fp32 rounding
L.check(L.load().piml_set_mlapm_algorithm(2), "piml_set_mlapm_algorithm")
ps, vs = p.clone(), v.clone()
for _ in range(steps):
    act, pn, arr = model.advance(ps, vs, ds, dest, bench.DT, bench.RADIUS)
    ps, vs = pn, act
L.check(L.load().piml_set_mlapm_algorithm(0), "piml_set_mlapm_algorithm")
sym = ShardedCrowd(N, device=dev, symmetric=True)
sym.load(p, v)
for _ in range(steps):
    sym.step(model, ds, dest, bench.DT, bench.RADIUS)
torch.cuda.synchronize()
gathered = [torch.empty_like(sym.position) for _ in range(world)]
dist.all_gather(gathered, sym.position.contiguous())
same = all(torch.equal(g_, gathered[0]) for g_ in gathered)
e_sym = float((sym.position - ps).norm(dim=-1).max())
e_ord = float((sym.position - pu).norm(dim=-1).max())
e_v = float(((sym.velocity - vu).norm(dim=-1) / vu.norm(dim=-1).clamp_min(1e-3)).max())
ok2 = same and e_sym < 1e-4 and e_ord < 1e-4 and bool(torch.isfinite(sym.position).all())
t2 = torch.tensor([1 if ok2 else 0], device=dev)
dist.all_reduce(t2, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"symmetric sharded: identical on all ranks {same}; max |dp| vs unsharded symmetric {e_sym:.2e} m, vs "
          f"ordered {e_ord:.2e} m, max rel dv {e_v:.2e} after {steps} steps (rows of rank 0: {sym.rows}): {bool(int(t2))}")
dist.destroy_process_group()
sys.exit(0 if int(t) and int(t2) else 1)
