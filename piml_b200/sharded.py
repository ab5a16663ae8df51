"""Agent-sharded MLAPM crowd across the GPUs of one node (BASELINE config 5a).

Rank g owns rows [g N/G, (g+1) N/G) and needs every agent's (position, velocity) each step -- the path's only exchange.
Instead of a separate all-gather, the finalize kernel of `piml_mlapm_advance_push_f32` stores each new row straight
into EVERY rank's next-state arrays over NVLink / NVSwitch peer memory (torch symmetric memory provides the peer
mappings), so the transfer rides on the kernel's own epilogue; a step then needs one cross-rank barrier only.

`symmetric=True` (the default for crowds of >= 16 384 agents per the library's own rule): every UNORDERED pair is
evaluated once for both rows.  Rank g then owns whole 512-agent blocks and evaluates the block pairs of its own row
blocks; the column-direction sums that belong to other ranks' agents are reduced per rank and stored into the owners'
inboxes by the pair stage (`piml_mlapm_sym_pairs_push_f32`), a barrier, and the owners' finalize kernel
(`piml_mlapm_sym_finalize_push_f32`) adds the shares in rank order and pushes the new state as above.

One process per GPU (`torchrun`), `torch.distributed` initialised with the NCCL backend.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L


def shard_rows(N, world, rank):
    """Block partition of the agents: rank g owns rows [g N/G, (g+1) N/G)."""
    if N % world:
        raise ValueError(f"{N} agents are not divisible by {world} ranks")
    shard = N // world
    return rank * shard, (rank + 1) * shard


def allgather_state(pos_next, vel_next, pos_rows, vel_rows, group=None):
    """The exchange as a separate collective (the baseline the fused push replaces; also what a backend without peer
    memory uses): every rank contributes its rows' new positions / velocities, all ranks end with the full (N,2)
    arrays.  Works on NCCL (GPU) and gloo (CPU tests of the partition logic)."""
    if dist.get_backend(group) == "gloo":
        world = dist.get_world_size(group)
        for full, part in ((pos_next, pos_rows), (vel_next, vel_rows)):
            chunks = list(full.chunk(world, dim=0))
            dist.all_gather(chunks, part.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(pos_next, pos_rows, group=group)
        dist.all_gather_into_tensor(vel_next, vel_rows, group=group)
    return pos_next, vel_next


def allgather_rows(full, part, group=None):
    """All-gather of equally sized row blocks: rank g contributes `part` = rows [g n/G, (g+1) n/G) of `full` (n, ...).
    NCCL: one all_gather_into_tensor; gloo (CPU tests of the host logic): all_gather into row-block views."""
    if dist.get_backend(group) == "gloo":
        chunks = list(full.chunk(dist.get_world_size(group), dim=0))
        dist.all_gather(chunks, part.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(full, part.contiguous(), group=group)
    return full


class ShardedCrowd(object):
    """Double-buffered crowd state in symmetric memory: buf[parity][0] = positions (N,2), buf[parity][1] = velocities."""

    SYM_MIN_AGENTS = 16384

    def __init__(self, N, group=None, device=None, symmetric=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = device if device is not None else L.cuda_device()
        self.N = N
        self.symmetric = (N >= self.SYM_MIN_AGENTS) if symmetric is None else bool(symmetric)
        if self.symmetric:
            r0, r1 = C.c_int64(), C.c_int64()
            L.check(L.load().piml_mlapm_sym_shard_rows(N, self.world, self.rank, C.byref(r0), C.byref(r1)),
                    "piml_mlapm_sym_shard_rows")
            self.rows = (int(r0.value), int(r1.value))
            self.inbox = symm.empty((int(L.load().piml_mlapm_sym_inbox_bytes(N, self.world)) // 4,),
                                    dtype=torch.float32, device=self.device)
            self.inbox_hdl = symm.rendezvous(self.inbox, self.group)
            self._inbox_tab = (C.c_uint64 * self.world)(*[int(p) for p in self.inbox_hdl.buffer_ptrs])
            self._sym_ws = torch.empty(int(L.load().piml_mlapm_sym_shard_workspace_bytes(N, self.world)),
                                       dtype=torch.uint8, device=self.device)
        else:
            self.rows = shard_rows(N, self.world, self.rank)
        self.shard = self.rows[1] - self.rows[0]
        self.buf = symm.empty((2, 2, N, 2), dtype=torch.float32, device=self.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        if len(self.ptrs) != self.world:
            raise RuntimeError("symmetric memory rendezvous returned an unexpected number of peer buffers")
        # destinations of ALL agents (the column side of the symmetric gates needs them): symmetric too, so that a rank
        # can upload only its own rows and hand them to its peers (scatter_rows)
        self.dest_buf = symm.empty((N, 2), dtype=torch.float32, device=self.device)
        self.dest_hdl = symm.rendezvous(self.dest_buf, self.group)
        self._dest_tab = (C.c_uint64 * self.world)(*[int(p) for p in self.dest_hdl.buffer_ptrs])
        self.parity = 0
        self.arrived = torch.empty(self.shard, dtype=torch.uint8, device=self.device)
        plane = N * 2 * 4                                        # bytes of one (N,2) fp32 array
        self._tables = []
        for par in (0, 1):
            pos = (C.c_uint64 * self.world)(*[p + (par * 2) * plane for p in self.ptrs])
            vel = (C.c_uint64 * self.world)(*[p + (par * 2 + 1) * plane for p in self.ptrs])
            self._tables.append((pos, vel))

    def load(self, position, velocity):
        """Every rank passes the full initial state (N,2)."""
        self.buf[self.parity, 0].copy_(position)
        self.buf[self.parity, 1].copy_(velocity)
        torch.cuda.current_stream(self.device).synchronize()
        dist.barrier(self.group)

    def scatter_rows(self, position_rows, velocity_rows, destination_rows=None):
        """The host-buffer entry of a sharded step: every rank passes ONLY ITS rows [rows[0], rows[1]) (pinned host or
        device tensors): one H2D of 1/G of the state, then the rows are stored into every rank's current-state arrays
        over NVLink peer memory and a barrier makes the crowd complete everywhere.  destination_rows likewise into
        .dest_buf."""
        r0, r1 = self.rows
        dev = self.device
        stage = [L.f32c(x.to(dev, non_blocking=True) if not x.is_cuda else x) for x in (position_rows, velocity_rows)]
        dst = destination_rows
        if dst is not None:
            dst = L.f32c(dst.to(dev, non_blocking=True) if not dst.is_cuda else dst)
        pos_tab, vel_tab = self._tables[self.parity]
        L.check(L.load().piml_scatter_rows_push_f32(
            L.ptr(stage[0]), L.ptr(stage[1]), L.ptr(dst), r0, r1 - r0, self.world, pos_tab, vel_tab,
            self._dest_tab if dst is not None else None, L.stream_ptr(dev)), "piml_scatter_rows_push_f32")
        self.hdl.barrier(channel=0)

    @property
    def position(self):
        return self.buf[self.parity, 0]

    @property
    def velocity(self):
        return self.buf[self.parity, 1]

    def step(self, model, desired_speed, destination, dt, radius=0.3, trace=None):
        """One iteration of src/main_mlapm.py:18-36 for this rank's rows, exchange included.  Returns arrived[bool]
        for the local rows; afterwards .position / .velocity hold the new state of ALL agents.
        trace: optional list that receives CUDA events at the stage boundaries of this step (pairs + share push |
        barrier | finalize + state push | barrier), for per-rank timelines."""
        nxt = 1 - self.parity
        pos_tab, vel_tab = self._tables[nxt]
        ds = desired_speed if desired_speed.dim() == 2 else desired_speed.unsqueeze(-1)
        r0, r1 = self.rows

        def mark():
            if trace is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                trace.append(e)
        mark()
        if self.symmetric:
            lib, prm, st = L.load(), model._params(), L.stream_ptr(self.device)
            L.check(lib.piml_mlapm_sym_pairs_push_f32(
                L.ptr(self.position), L.ptr(self.velocity), L.ptr(destination), self.N, self.world, self.rank,
                C.byref(prm), self._inbox_tab, L.ptr(self._sym_ws), self._sym_ws.numel(), st),
                "piml_mlapm_sym_pairs_push_f32")
            mark()
            self.hdl.barrier(channel=0)    # every rank's column-direction shares have landed in the owners' inboxes
            mark()
            L.check(lib.piml_mlapm_sym_finalize_push_f32(
                L.ptr(self.position), L.ptr(self.velocity), L.ptr(ds), ds.shape[1], L.ptr(destination), self.N,
                self.world, self.rank, C.byref(prm), float(dt), float(radius), L.ptr(self.inbox), pos_tab, vel_tab,
                L.ptr(self.arrived), L.ptr(self._sym_ws), self._sym_ws.numel(), st),
                "piml_mlapm_sym_finalize_push_f32")
            mark()
            self.hdl.barrier(channel=0)    # ... and every rank's new rows in every rank's next-state arrays
            mark()
            self.parity = nxt
            return self.arrived.bool()
        L.check(L.load().piml_mlapm_advance_push_f32(
            L.ptr(self.position), L.ptr(self.velocity), L.ptr(ds), ds.shape[1], L.ptr(destination), self.N, r0, r1,
            C.byref(model._params()), float(dt), float(radius), self.world, pos_tab, vel_tab, L.ptr(self.arrived),
            L.ptr(model._workspace(self.N, self.device)), L.stream_ptr(self.device)), "piml_mlapm_advance_push_f32")
        mark()
        self.hdl.barrier(channel=0)        # every rank's rows have landed in every rank's next-state arrays
        mark()
        self.parity = nxt
        return self.arrived.bool()


class ShardedNNCrowd(object):
    """Agent-sharded NN-augmented (or pure social-force) rollout of ONE large scene: the loop body of
    `get_multiple_rollouts` (simulators.py:602-652) with rank g computing rows [g N/G, (g+1) N/G).

    Every rank keeps the whole state.  Per step: feature rebuild of the OWN rows against all agents
    (`piml_state_features_rows_f32`, cell list) -> network forward of the own rows -> the path's one exchange, an
    all-gather of the new accelerations (8 B per agent -- the only quantity a rank cannot compute for rows it does not
    own) -> the cheap elementwise state update (Euler / arrival / waypoints, `piml_integrate_step_f32`) replicated on
    every rank, which keeps the replicas bit-identical without exchanging positions and velocities.
    Works on NCCL; results are bit-identical to the unsharded step sequence (`scripts/check_nn_sharded.py`)."""

    def __init__(self, model, args, N, obstacles, group=None, device=None):
        from . import models as M
        from .sfm import SocialForce
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = device if device is not None else L.cuda_device()
        self.N, self.rows = N, shard_rows(N, self.world, self.rank)
        self.args, self.model = args, model
        self.sfm = isinstance(model, SocialForce)
        if not self.sfm:
            self.spec = model.spec if hasattr(model, "spec") else M.spec_from_module(model)
            self.packed = M.pack_device(model.state_dict(), self.spec, self.device)
            self.packed_tc = M.pack_device_tc(model.state_dict(), self.spec, self.device)
        self.obstacles = L.f32c(obstacles.to(self.device))
        self.Mo = self.obstacles.shape[-2] if self.obstacles.numel() else 0
        n = self.rows[1] - self.rows[0]
        kp, ko = min(args.topk_ped, N), (min(args.topk_obs, self.Mo) if self.Mo else 0)
        dev = self.device
        self.ped_f, self.obs_f = torch.empty(n, kp, 6, device=dev), torch.empty(n, ko, 6, device=dev)
        self.self_f, self.dest_f = torch.empty(n, 7, device=dev), torch.empty(n, 2, device=dev)
        self.a_next = torch.empty(1, N, 2, device=dev)

    def load(self, position, velocity, acceleration, destination, dest_idx, dest_num, waypoints, desired_speed,
             hist_v=None):
        """Every rank passes the whole initial state: (N,2) x4, dest_idx / dest_num (N) int64, waypoints (D,N,2),
        desired_speed (N)."""
        dv = lambda x, dt=torch.float32: x.to(self.device, dt).contiguous()
        self.p, self.v, self.a, self.dest = [dv(x)[None].clone() for x in (position, velocity, acceleration,
                                                                           destination)]
        self.didx, self.dnum = dv(dest_idx, torch.int64)[None].clone(), dv(dest_num, torch.int64)[None].clone()
        self.wp, self.ds = dv(waypoints)[None].clone(), dv(desired_speed).reshape(1, self.N).clone()
        self.hist = (dv(hist_v)[None] if hist_v is not None else self.v).clone()
        self._features()

    def _features(self):
        from .features import cos_threshold
        a, (r0, r1) = self.args, self.rows
        L.check(L.load().piml_state_features_rows_f32(
            L.ptr(self.p), L.ptr(self.v), L.ptr(self.a), L.ptr(self.dest), L.ptr(self.obstacles) if self.Mo else None,
            self.N, self.Mo, r0, r1, a.topk_ped, cos_threshold(a.sight_angle_ped), float(a.dist_threshold_ped),
            a.topk_obs, cos_threshold(a.sight_angle_obs), float(a.dist_threshold_obs), L.ptr(self.hist),
            L.ptr(self.ds), L.ptr(self.ped_f), L.ptr(self.obs_f) if self.Mo else None, L.ptr(self.self_f),
            L.ptr(self.dest_f), L.stream_ptr(self.device)), "piml_state_features_rows_f32")

    def step(self, dt=None, remove_on_arrival=True):
        """One rollout step; afterwards .p / .v / .a hold the new state of ALL agents on every rank."""
        from . import models as M
        from .rollout import integrate_step
        dt = float(self.args.time_unit if dt is None else dt)
        if self.sfm:
            a_own = self.model(self.ped_f, self.obs_f, self.self_f)[0]
        else:
            a_own = M.pinnsf_forward(self.spec, self.packed, self.ped_f, self.obs_f, self.self_f, need_msgs=False,
                                     packed_tc=self.packed_tc)[0]
        if self.world > 1:                   # the path's one exchange: 8 B per agent
            allgather_rows(self.a_next.view(self.N, 2), a_own, self.group)
        else:
            self.a_next.view(self.N, 2).copy_(a_own)
        integrate_step(self.p, self.v, self.a, self.a_next, self.dest, self.didx, self.dnum, self.wp, dt,
                       remove_on_arrival, hist_v=self.hist)
        self._features()
