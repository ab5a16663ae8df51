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
# ---- the benchmarked configuration against the ORACLE, sharded: N = 100 000, one step, force level (dt = 2^20) ----------
import numpy as np
from oracle import oracle as O
N2 = 100000
p2, v2, ds2, dest2, _ = bench.synthetic_crowd(N2)
big = ShardedCrowd(N2, device=dev)                      # symmetric (>= 16 384 agents)
big.load(p2.to(dev), v2.to(dev))
r0, r1 = big.rows
# host-buffer entry: every rank hands over only its rows; afterwards all ranks must hold the full state
big.scatter_rows(p2[r0:r1].pin_memory(), v2[r0:r1].pin_memory(), dest2[r0:r1].pin_memory())
torch.cuda.synchronize()
ok_sc = bool(torch.equal(big.position.cpu(), p2) and torch.equal(big.velocity.cpu(), v2) and torch.equal(big.dest_buf.cpu(), dest2))
BIG_DT = float(2 ** 20)
big.step(model, ds2.to(dev), big.dest_buf, BIG_DT, bench.RADIUS)
torch.cuda.synchronize()
act_big = big.velocity.cpu().numpy().astype(np.float64)
rows = [(r0, min(r0 + 300, r1)), (max(r1 - 300, r0), r1), ((r0 + r1) // 2 + 37, min((r0 + r1) // 2 + 337, r1))]
worst, worst_k, nrows_chk, okf = 0.0, 0.0, 0, True
for (a_, b_) in rows:
    _, force, opsum = O.mlapm_step_diag(p2.numpy(), v2.numpy(), ds2.numpy(), dest2.numpy(), bench.DT, "GC", rows=(a_, b_))
    F = (act_big[a_:b_] - v2.numpy()[a_:b_].astype(np.float64)) / BIG_DT
    e = np.linalg.norm(F - force, axis=-1); nF = np.linalg.norm(force, axis=-1)
    okf = okf and bool((e <= 1e-5 * np.maximum(nF, opsum / 16.0)).all())
    i = int(np.argmax(e / nF))
    if (e / nF)[i] > worst:
        worst, worst_k = float((e / nF)[i]), float(opsum[i] / nF[i])
    nrows_chk += b_ - a_
t3 = torch.tensor([1 if (okf and ok_sc) else 0], device=dev)
dist.all_reduce(t3, op=dist.ReduceOp.MIN)
print(f"rank {rank}: N={N2} sharded symmetric step vs oracle on {nrows_chk} of its rows [{r0},{r1}): strict force error max "
      f"{worst:.2e} (kappa {worst_k:.0f}), gate ok {okf}; scatter_rows rebuilt the full state: {ok_sc}", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t) and int(t2) and int(t3) else 1)
