"""One rollout-training step -- `test_multiple_rollouts_for_training` (reference src/models/simulators.py:659-832),
`loss.backward()` (:359) and the optimizer step (:360) -- captured ONCE into a CUDA graph and replayed per batch.

Why: at the reference's batch sizes (C = 32 channels x T = 5 frames x N = 144 slots) every kernel of the step takes a
few microseconds, and the step is ~190 launches issued from Python through autograd: 13 ms of host time for well under
1 ms of GPU work.  A replayed graph removes the host from the step.  Everything captured is the library's kernels and
torch's own elementwise / optimizer kernels; the step's data-dependent host decisions are hoisted out of it:

  * `if torch.sum(mask) > 0` (:705) is evaluated for all frames of the batch BEFORE the replay (one device -> host
    read); a batch with an empty frame takes the eager path;
  * the NaN assert (:745) and the collision counters (:788-789) are read AFTER the replay.

Batches must keep the shapes of the batch the graph was captured with (the reference's channelled windows do);
their tensors are copied into static buffers before every replay (the step modifies `data.labels` and
`data.dest_idx` in place, like the reference).  The optimizer must be constructed with `capturable=True`.
"""
import copy

import torch

from . import train_rollout as TRO


class GraphedRolloutTraining(object):
    def __init__(self, simulator, optimizer, example_batch, warmup=3):
        self.sim, self.opt = simulator, optimizer
        for g in optimizer.param_groups:
            if not g.get("capturable", False):
                raise ValueError("GraphedRolloutTraining needs an optimizer constructed with capturable=True")
        self.static = self._clone(example_batch)
        self._keys = [k for k, v in self.static.__dict__.items() if torch.is_tensor(v)]
        self._shapes = {k: tuple(getattr(self.static, k).shape) for k in self._keys}
        dev = self.static.position.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up on the capture stream: sizes every scratch buffer
            for _ in range(warmup):
                self._load(example_batch)
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._load(example_batch)
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph, stream=side):
            self._out = TRO.test_multiple_rollouts_for_training(self.sim, self.static, _sync_free=True)
            self._out[0].backward()
            self.opt.step()
        self._deferred = self.sim._deferred

    @staticmethod
    def _clone(batch):
        b = copy.copy(batch)
        for k, v in batch.__dict__.items():
            if torch.is_tensor(v):
                setattr(b, k, v.clone())
        return b

    def _load(self, batch):
        for k in self._keys:
            src = getattr(batch, k)
            if tuple(src.shape) != self._shapes[k]:
                raise ValueError(f"batch.{k} has shape {tuple(src.shape)}, the graph was captured with {self._shapes[k]}")
            getattr(self.static, k).copy_(src)

    def _eager_step(self):
        self.opt.zero_grad(set_to_none=True)
        out = TRO.test_multiple_rollouts_for_training(self.sim, self.static)
        out[0].backward()
        self.opt.step()
        return out

    def step(self, batch):
        """One training step on `batch`; returns the reference's 7-tuple (tensors owned by the graph: read or clone them
        before the next step).  Updates simulator.collision_count / hard_collision_count like the reference."""
        self._load(batch)
        frames_ok = bool((self.static.mask_p_pred.long().sum(dim=(0, 2)) > 0).all())      # simulators.py:705, hoisted
        if not frames_ok:
            return self._eager_step()
        self.graph.replay()
        nan_flag, coll, hard = self._deferred
        stats = torch.stack([nan_flag.float(), coll, hard]).tolist()                       # the step's one read-back
        assert stats[0] == 0.0, 'find nan in a rollout-training step'                      # :745
        self.sim.collision_count = getattr(self.sim, 'collision_count', 0) + stats[1]
        self.sim.hard_collision_count = getattr(self.sim, 'hard_collision_count', 0) + stats[2]
        return self._out
