"""World-size-2 gloo test (CPU) of the agent-sharded path's host logic: the block partition of the rows
and the exchange layout.  The per-rank compute is the oracle's row-range MLAPM step (there is no GPU here); the
sharded crowd after the exchange must be bit-identical to the unsharded step on every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, steps, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from piml_b200.sharded import allgather_state, shard_rows
        O.set_num_threads(1)
        rng = np.random.default_rng(7)
        L = np.sqrt(N / 0.5)
        p = (rng.random((N, 2)) * L).astype(np.float32)
        d = (rng.random((N, 2)) * L).astype(np.float32)
        v = rng.normal(0, 1, (N, 2)).astype(np.float32)
        ds = np.full((N, 1), 1.3, np.float32)
        r0, r1 = shard_rows(N, world, rank)
        assert (r1 - r0) * world == N and r0 == rank * (N // world)
        pos, vel = torch.from_numpy(p.copy()), torch.from_numpy(v.copy())
        pu, vu = p.copy(), v.copy()
        for _ in range(steps):
            act = O.mlapm_step(pos.numpy(), vel.numpy(), ds, d, 0.08, "GC", rows=(r0, r1))      # this rank's rows
            pnew = pos.numpy()[r0:r1] + act * np.float32(0.08)                                    # main_mlapm.py:26
            pos_next, vel_next = torch.empty_like(pos), torch.empty_like(vel)
            allgather_state(pos_next, vel_next, torch.from_numpy(pnew), torch.from_numpy(act))
            pos, vel = pos_next, vel_next
            au = O.mlapm_step(pu, vu, ds, d, 0.08, "GC")                                          # unsharded
            pu, vu = pu + au * np.float32(0.08), au
        ok = np.array_equal(pos.numpy(), pu) and np.array_equal(vel.numpy(), vu)
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


def test_sharded_rows_and_exchange_match_unsharded_world2():
    world, N, steps = 2, 96, 3
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), N, steps, out), nprocs=world, join=True)
        assert dict(out) == {0: 1, 1: 1}


def test_shard_rows_rejects_uneven_split():
    import pytest
    from piml_b200.sharded import shard_rows
    assert shard_rows(100, 4, 3) == (75, 100)
    with pytest.raises(ValueError):
        shard_rows(10, 3, 0)


# ---- agent-sharded NN / social-force step: exchange of the accelerations ---------------------------------------------
def _nn_worker(rank, world, port, N, steps, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from piml_b200.sharded import allgather_rows, shard_rows
        rng = np.random.default_rng(11)
        side = np.sqrt(N / 0.5)
        p = (rng.random((N, 2)) * side).astype(np.float32)
        d = (rng.random((N, 2)) * side).astype(np.float32)
        v = rng.normal(0, 1, (N, 2)).astype(np.float32)
        a = np.zeros((N, 2), np.float32)
        obs = (rng.random((40, 2)) * side).astype(np.float32)
        v0 = np.full((N,), 1.3, np.float32)
        didx, dnum = np.zeros(N, np.int64), np.ones(N, np.int64)
        wp = d[None].copy()
        r0, r1 = shard_rows(N, world, rank)

        def features(pp, vv, aa, dd):
            pf, of, df = O.relative_features(pp[None], vv[None].copy(), aa[None].copy(), dd[None], obs)
            return pf[0], of[0], np.concatenate([df[0], vv, aa, v0[:, None]], -1)

        # sharded: every rank keeps the whole state; own-row model evaluation, all-gather of the accelerations,
        # replicated state update (what ShardedNNCrowd.step does with the CUDA kernels)
        ps, vs, as_, ds_, di = p.copy(), v.copy(), a.copy(), d.copy(), didx.copy()
        pu, vu, au, du, diu = p.copy(), v.copy(), a.copy(), d.copy(), didx.copy()
        for _ in range(steps):
            pf, of, sf = features(ps, vs, as_, ds_)
            a_own = O.sfm_forward(pf[r0:r1], of[r0:r1], sf[r0:r1], "gc1560")[0]
            a_next = torch.empty(N, 2)
            allgather_rows(a_next, torch.from_numpy(a_own))
            ps, vs, as_, ds_, di, _ = O.integrate_step(ps, vs, as_, a_next.numpy(), ds_, di, dnum, wp, 0.08, True)
            pf, of, sf = features(pu, vu, au, du)
            a_full = O.sfm_forward(pf, of, sf, "gc1560")[0]
            pu, vu, au, du, diu, _ = O.integrate_step(pu, vu, au, a_full, du, diu, dnum, wp, 0.08, True)
        same = all(np.array_equal(x, y, equal_nan=True) for x, y in ((ps, pu), (vs, vu), (as_, au), (ds_, du)))
        out[rank] = 1 if same else 0
    finally:
        dist.destroy_process_group()


def test_sharded_nn_step_exchange_matches_unsharded_world2():
    """World-size-2 gloo test of ShardedNNCrowd's host logic: own-row model evaluation + all-gather of the accelerations
    (allgather_rows) + replicated state update == the unsharded step sequence, bit for bit on every rank.  The per-rank
    compute is the oracle (no GPU here)."""
    world, N, steps = 2, 64, 4
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_nn_worker, args=(world, _free_port(), N, steps, out), nprocs=world, join=True)
        assert dict(out) == {0: 1, 1: 1}
