"""Build-container only (skipped where /root/reference is absent, e.g. on the GPU box): the oracle against the
UNMODIFIED reference imported live, on EVERY frame of the reference's own clips (toy, GC 1000-1060,
synthetic 1560-1620, UCY 0-54) -- the committed golden vectors hold a subset of these frames."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests.util import valid_sets

HAVE_REF = os.path.isdir("/root/reference/src")
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")


def _harness():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import _refharness as H
    return H


@pytest.mark.parametrize("clip,step", [("TOY_CLIP", 1), ("GC_CLIP", 3), ("SYN_CLIP", 3), ("UCY_CLIP", 4)])
def test_oracle_features_equal_live_reference_on_whole_clips(clip, step):
    """get_relative_features (data.py:466-512) frame by frame (T = 1 calls, as the rollout makes them) over the clip."""
    H = _harness()
    DATA, *_ = H.import_reference()
    torch.set_num_threads(1)
    raw = H.load_raw(getattr(H, clip))
    ped = DATA.Pedestrians()
    T = raw.position.shape[0]
    n = 0
    for t in range(0, T, step):
        p, d = raw.position[t:t + 1].clone(), raw.destination[t:t + 1].clone()
        v, a = raw.velocity[t:t + 1].clone(), raw.acceleration[t:t + 1].clone()
        if not bool((~torch.isnan(p[..., 0])).any()):
            continue
        with H.quiet():
            pf, of, df = ped.get_relative_features(p, v, a, d, raw.obstacles.clone(), 6, 90, 4, 10, 90, 4)
        v2, a2 = raw.velocity[t:t + 1].numpy().copy(), raw.acceleration[t:t + 1].numpy().copy()
        want = O.relative_features(raw.position[t:t + 1].numpy(), v2, a2, raw.destination[t:t + 1].numpy(),
                                   raw.obstacles.numpy())
        assert np.array_equal(want[0], pf.numpy()), (clip, t)
        assert np.array_equal(want[1], of.numpy()), (clip, t)
        assert np.array_equal(want[2], df.numpy()), (clip, t)
        assert np.array_equal(v2, v.numpy()) and np.array_equal(a2, a.numpy())      # in-place NaN -> 0 side effect
        n += 1
    assert n >= 40


def test_oracle_selection_equals_live_reference_on_gc_frames():
    """get_nearby_obj_in_sight (data.py:416-447): valid sets {idx: dist <= thr} and distances, angles 90 and 100."""
    H = _harness()
    DATA, *_ = H.import_reference()
    raw = H.load_raw(H.GC_CLIP)
    ped = DATA.Pedestrians()
    for t in range(30, raw.position.shape[0], 45):
        p, v = raw.position[t:t + 1].clone(), raw.velocity[t:t + 1].clone()
        v[torch.isnan(v)] = 0
        head = ped.get_heading_direction(v)
        for angle in (90, 100):
            dist, idx = ped.get_nearby_obj_in_sight(p.clone(), p.clone(), head, 6, angle)
            od, oi = O.select(p.numpy(), p.numpy(), head.numpy(), 6, angle)
            assert np.array_equal(od, dist.numpy()), (t, angle)
            assert valid_sets(oi, od, 4) == valid_sets(idx.numpy(), dist.numpy(), 4), (t, angle)
