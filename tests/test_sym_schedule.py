"""CPU tests of the host logic of the symmetric (unordered-pair) MLAPM evaluation (piml_b200/csrc/mlapm.cu):
the circulant block-pair schedule, the block ownership of an agent-sharded crowd (through the C ABI, no GPU needed)
and -- world size 2 on gloo -- the two-stage exchange layout (per-rank column-direction shares delivered to the
owners' inboxes, added in rank order)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

BLOCK = 512


def block_pairs(T):
    """(I, J, d) of every block pair the kernel evaluates: CTA row block I, d = 0..L-1, J = I + d mod T."""
    D = T // 2
    out = []
    for I in range(T):
        L = D + 1
        if T % 2 == 0 and 2 * I >= T:
            L = D
        for d in range(L):
            out.append((I, (I + d) % T, d))
    return out


def finalize_sources(T, J):
    """The (I, d) whose column-direction sums the finalize kernel adds for an agent of block J."""
    D = T // 2
    src = []
    for d in range(1, D + 1):
        I = (J - d) % T
        if d == D and T % 2 == 0 and 2 * I >= T:
            continue
        src.append((I, d))
    return src


@pytest.mark.parametrize("T", list(range(1, 42)) + [195, 196])
def test_every_unordered_block_pair_exactly_once(T):
    pairs = block_pairs(T)
    seen = {}
    for I, J, d in pairs:
        key = (min(I, J), max(I, J))
        assert key not in seen, f"block pair {key} evaluated twice"
        seen[key] = (I, d)
        assert (d == 0) == (I == J)
    assert len(seen) == T * (T + 1) // 2
    per_block = [sum(1 for I, _, _ in pairs if I == b) for b in range(T)]
    assert max(per_block) - min(per_block) <= 1                   # balanced: every row block the same work
    for J in range(0, T, max(1, T // 7)):
        want = sorted((I, d) for I, JJ, d in pairs if JJ == J and d > 0)
        assert sorted(finalize_sources(T, J)) == want


@pytest.mark.parametrize("N,world", [(100000, 1), (100000, 2), (100000, 8), (1000000, 8), (16384, 4), (8192 * 2, 2),
                                     (5 * BLOCK + 100, 2)])
def test_sym_shard_rows_cover_the_crowd(N, world):
    from piml_b200 import _lib
    lib = _lib.load()
    T = (N + BLOCK - 1) // BLOCK
    edges = []
    for g in range(world):
        r0, r1 = C.c_int64(), C.c_int64()
        assert lib.piml_mlapm_sym_shard_rows(N, world, g, C.byref(r0), C.byref(r1)) == 0
        assert r0.value % BLOCK == 0 and (r1.value % BLOCK == 0 or r1.value == N)
        edges.append((r0.value, r1.value))
    assert edges[0][0] == 0 and edges[-1][1] == N
    assert all(edges[g][1] == edges[g + 1][0] for g in range(world - 1))
    blocks = [-(-(b - a) // BLOCK) for a, b in edges]
    assert sum(blocks) == T and max(blocks) - min(blocks) <= 1
    stride = -(-T // world) * BLOCK
    assert lib.piml_mlapm_sym_inbox_bytes(N, world) == stride * world * 16
    assert lib.piml_mlapm_sym_shard_workspace_bytes(N, world) > 0
    assert lib.piml_mlapm_workspace_bytes_sym(N) >= lib.piml_mlapm_workspace_bytes(N)


def test_sym_shard_rows_rejects_more_ranks_than_blocks():
    from piml_b200 import _lib
    r0, r1 = C.c_int64(), C.c_int64()
    assert _lib.load().piml_mlapm_sym_shard_rows(1000, 4, 0, C.byref(r0), C.byref(r1)) != 0


# ---- world size 2 on gloo: the exchange layout -------------------------------------------------------------------------
def _pair_term(p, n, m):
    """An asymmetric pair term f(n <- m) with a gate of the row's own state (stands in for view * w * direc)."""
    r = p[m] - p[n]
    d2 = (r * r).sum(-1) + 1.0
    gate = (r[..., 0] * np.cos(n) + r[..., 1] * np.sin(n)) > 0
    return np.where(gate[..., None], r / d2[..., None], 0.0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from piml_b200 import _lib
        lib = _lib.load()
        T = (N + BLOCK - 1) // BLOCK
        rows = []
        for g in range(world):
            r0, r1 = C.c_int64(), C.c_int64()
            assert lib.piml_mlapm_sym_shard_rows(N, world, g, C.byref(r0), C.byref(r1)) == 0
            rows.append((r0.value, r1.value))
        Ib = [r[0] // BLOCK for r in rows] + [T]
        stride = -(-T // world) * BLOCK
        rng = np.random.default_rng(3)
        p = rng.normal(0, 3, (N, 2))
        idx = np.arange(N)
        # this rank's block pairs: row-direction sums stay here, column-direction sums go to the owner of the column
        rowsum = np.zeros((N, 2))
        colshare = np.zeros((N, 2))
        for I, J, d in block_pairs(T):
            if not (Ib[rank] <= I < Ib[rank + 1]):
                continue
            n = idx[I * BLOCK:min((I + 1) * BLOCK, N)]
            m = idx[J * BLOCK:min((J + 1) * BLOCK, N)]
            nn, mm = np.meshgrid(n, m, indexing="ij")
            if d == 0:
                f = _pair_term(p, nn, mm)
                f[nn == mm] = 0.0
                rowsum[n] += f.sum(1)
            else:
                rowsum[n] += _pair_term(p, nn, mm).sum(1)
                colshare[m] += _pair_term(p, mm, nn).sum(0)
        # exchange: inbox[h][g, local row] = rank g's share for the rows of rank h (the layout of mlapm_sym_colpush_kernel)
        send = []
        for h in range(world):
            buf = torch.zeros(stride, 2, dtype=torch.float64)
            a, b = rows[h]
            buf[:b - a] = torch.from_numpy(colshare[a:b])
            send.append(buf)
        inbox = [torch.zeros(stride, 2, dtype=torch.float64) for _ in range(world)]
        for h in range(world):                                   # gloo has no all_to_all: one gather per owner
            dist.gather(send[h], inbox if rank == h else None, dst=h)
        a, b = rows[rank]
        total = rowsum[a:b].copy()
        for g in range(world):                                   # rank order, like the finalize kernel
            total += inbox[g][:b - a].numpy()
        nn, mm = np.meshgrid(idx[a:b], idx, indexing="ij")
        f = _pair_term(p, nn, mm)
        f[nn == mm] = 0.0
        want = f.sum(1)
        out[rank] = float(np.abs(total - want).max())
    finally:
        dist.destroy_process_group()


def test_symmetric_exchange_layout_world2():
    world, N = 2, 5 * BLOCK + 100
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), N, out), nprocs=world, join=True)
        assert set(out.keys()) == {0, 1}
        assert max(out.values()) < 1e-9, dict(out)
