"""The drop-in, end to end: the UNMODIFIED reference (its scripts, simulator class and data containers, imported from
/root/reference here or from the verbatim copy baseline/_ref/ on the GPU box) with piml_b200.patch installed must run
its hot path in libpiml_b200.so and reproduce what the unpatched reference produced (the committed golden vectors,
and a live unpatched run on the host cores of the same box for the side effects callers can observe).

    src/main_mlapm.py:18-36                                   -> mlapm.npz `circle`
    BaseSimulator.get_multiple_rollouts (simulators.py:556)    -> rollout_*.npz, + in-place dest_idx / hist_v side effects
    test_multiple_rollouts_for_training + backward (:659, :359) -> training_rollout.npz
    TimeIndexedPedData.make_dataset (data.py:746)              -> the unpatched make_dataset of the same clip
"""
import copy
import os
import runpy
import sys
import types

import numpy as np
import pytest
import torch

from tests.util import golden, group

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import _refharness as H  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not H.available(), reason="no reference tree (run __graft_entry__.build() first)")]


@pytest.fixture()
def ref():
    """The reference modules with the CUDA path patched in; restored afterwards."""
    DATA, MODEL, MLAPM, SIM, UTILS = H.import_reference()
    import functions.metrics as METRIC
    import piml_b200.patch as patch
    names = patch.install(DATA=DATA, MLAPM_MOD=MLAPM, MODEL=MODEL, SIM=SIM, UTILS=UTILS, METRIC=METRIC)
    assert len(names) == 19
    try:
        yield types.SimpleNamespace(DATA=DATA, MODEL=MODEL, MLAPM=MLAPM, SIM=SIM, UTILS=UTILS, METRIC=METRIC,
                                    patch=patch)
    finally:
        patch.uninstall()


def _launches():
    from piml_b200 import _lib as L
    return L.launch_count()


def test_main_mlapm_script_runs_on_the_cuda_path(ref, monkeypatch):
    """python src/main_mlapm.py, unmodified (7 agents on a circle, 200 steps): only `matplotlib` (not installed; the
    script's last three lines plot) is replaced by a recorder, which is also how the trajectories are read back."""
    seen = {}
    plt = types.ModuleType("matplotlib.pyplot")
    plt.plot = lambda x, y, *a, **k: seen.update(x=x.clone(), y=y.clone())
    plt.axis = lambda *a, **k: None
    plt.show = lambda *a, **k: None
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    g = group(golden("mlapm"), "circle")
    before = _launches()
    torch.manual_seed(0)                                       # make_golden.py seeds the script's torch.rand the same way
    runpy.run_path(os.path.join(H.REF_SRC, "main_mlapm.py"), run_name="__main__")
    steps = g["position"].shape[1] - 1
    assert _launches() - before >= 2 * steps                   # every MLAPM.step ran in the library
    got = np.stack([seen["x"].numpy().T, seen["y"].numpy().T], -1)          # plt.plot(position[:, :, 0].T, ...)
    want = g["position"]
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want))                    # same arrival frames
    drift = np.nanmax(np.linalg.norm(got - want, axis=-1))
    print(f"main_mlapm.py through the patch: {steps} steps, free-running drift {drift:.2e} m")
    assert drift < 1e-3


ROLLOUTS = [("rollout_gc_bm", H.GC_CLIP, "pinnsf_bm", "gc1560", None, False),
            ("rollout_toy5_m", H.TOY_CLIP, "pinnsf_m", "ucy", None, False),
            ("rollout_ucy_bm", H.UCY_CLIP, "pinnsf_bm", "ucy", 200, False),
            ("rollout_syn_sfm", H.SYN_CLIP, "pinnsf_bm", "gc1560", None, True)]


def _make_sim(r, args, sfm):
    torch.manual_seed(666)
    with H.quiet():
        sim = r.SIM.BaseSimulator(args)
    if sfm:
        import piml_b200 as P
        sim.model = P.SocialForce(args.dataset_name)
    sim.model.eval()
    return sim


@pytest.mark.parametrize("name,clip,kind,dsn,max_frames,sfm", ROLLOUTS)
def test_get_multiple_rollouts_on_real_reference_objects(ref, name, clip, kind, dsn, max_frames, sfm):
    """BaseSimulator(args).get_multiple_rollouts(TimeIndexedPedData, t_start=25, load_model=False) with the patch on:
    returns the reference's RawData, same masks / arrival pattern as the unpatched run, trajectories within the
    chaos envelope, and weights taken from the reference module itself."""
    g = golden(name)
    args = H.default_args(model=kind, dataset_name=dsn)
    raw = H.load_raw(clip)
    data = H.make_time_indexed(args, raw)                      # patched make_dataset (features on the GPU)
    if max_frames is not None and max_frames < data.num_frames:
        data.num_frames = max_frames
    sim = _make_sim(ref, args, sfm)
    work = copy.deepcopy(data)
    before = _launches()
    with H.quiet(), torch.no_grad():
        res = sim.get_multiple_rollouts(work, t_start=25, load_model=False)
    assert type(res) is ref.DATA.RawData
    T = data.num_frames
    assert _launches() - before >= 3 * (T - 25) or sfm
    assert np.array_equal(res.mask_p[:T].numpy(), g["out/mask_p"])
    pos, want = res.position[:T].numpy(), g["out/position"]
    assert np.array_equal(np.isnan(pos), np.isnan(want))
    d = np.linalg.norm(np.nan_to_num(pos) - np.nan_to_num(want), axis=-1)
    early, late = d[:25 + 100].max(), d.max()
    print(f"{name}: patched rollout vs the reference's own: <=100 steps {early:.2e} m, all {T - 25} steps {late:.2e} m")
    assert early < 2e-4                                        # before chaos amplifies fp32 rounding
    assert late < 0.05
    assert np.array_equal(res.position[:26].numpy(), data.position[:26].numpy(), equal_nan=True)


def test_rollout_side_effects_match_the_unpatched_reference(ref):
    """The reference's loop works on VIEWS of the dataset (simulators.py:571-578): afterwards data.dest_idx[t_start]
    holds the final waypoint indices and data.self_features[t_start, :, 2:4] the first step's velocities.  Compare
    with a live unpatched run on the CPU over 60 steps of the GC clip."""
    args = H.default_args(model="pinnsf_bm", dataset_name="gc1560")
    raw = H.load_raw(H.GC_CLIP)
    data = H.make_time_indexed(args, raw)
    data.num_frames = 25 + 60
    sim = _make_sim(ref, args, False)
    a, b = copy.deepcopy(data), copy.deepcopy(data)
    with H.quiet(), torch.no_grad():
        got = sim.get_multiple_rollouts(a, t_start=25, load_model=False)
        ref.patch.uninstall()
        try:
            want = sim.get_multiple_rollouts(b, t_start=25, load_model=False)      # the reference's own loop, CPU
        finally:
            ref.patch.install(DATA=ref.DATA, MLAPM_MOD=ref.MLAPM, MODEL=ref.MODEL, SIM=ref.SIM, UTILS=ref.UTILS,
                              METRIC=ref.METRIC)
    assert torch.equal(a.dest_idx, b.dest_idx)
    assert torch.allclose(a.self_features, b.self_features, rtol=0, atol=1e-5, equal_nan=True)
    assert not torch.equal(b.self_features[25], data.self_features[25])            # the side effect exists
    for k in ("position", "velocity", "acceleration", "destination", "mask_p", "mask_p_pred"):
        assert torch.equal(getattr(a, k), getattr(b, k)) or torch.allclose(getattr(a, k), getattr(b, k),
                                                                           equal_nan=True)
    T = data.num_frames
    assert torch.equal(got.mask_p[:T], want.mask_p[:T])
    assert torch.allclose(got.position[:T], want.position[:T], rtol=0, atol=1e-4, equal_nan=True)
    assert torch.allclose(got.velocity[:T], want.velocity[:T], rtol=0, atol=1e-4, equal_nan=True)


@pytest.mark.parametrize("case,clip,kind,dsn,chans,steps,over", [
    ("ucy_bm", H.UCY_CLIP, "pinnsf_bm", "ucy", slice(200, 206), 5, {}),
    ("gc_bm_full", H.GC_CLIP, "pinnsf_bm", "gc1560", slice(300, 304), 6,
     dict(reg_weight=1e-3, teacher_weight=0.5, new_collision_loss_flag=1, time_decay=0.9))])
def test_training_rollout_and_backward_on_real_reference_objects(ref, case, clip, kind, dsn, chans, steps, over):
    """sim.test_multiple_rollouts_for_training(ChanneledTimeIndexedPedData batch on cuda) + loss.backward() with the
    patch on, against the unpatched reference's losses, collision counts and parameter gradients."""
    g = group(golden("training_rollout"), case)
    args = H.default_args(model=kind, dataset_name=dsn, valid_steps=steps, device="cuda", **over)
    raw = H.load_raw(clip)
    data = H.make_time_indexed(args, raw)
    ch = ref.DATA.ChanneledTimeIndexedPedData()
    with H.quiet():
        ch.load_from_time_indexed_peddata(data, stride=steps, mode='slice')
    batch = ref.DATA.ChanneledTimeIndexedPedData.slice(ch, chans)
    for f in ("ped_features", "obs_features", "self_features", "labels", "mask_p", "mask_p_pred", "position",
              "velocity", "acceleration", "destination", "dest_idx", "waypoints", "obstacles", "dest_num",
              "abnormal_mask"):
        setattr(batch, f, getattr(batch, f).clone().cuda())
    assert np.array_equal(batch.ped_features.cpu().numpy(), g["in/ped_features"])   # same inputs as the golden run
    sim = _make_sim(ref, args, False)
    assert next(sim.model.parameters()).is_cuda
    sim.collision_count, sim.hard_collision_count, sim.epoch, sim.batch_idx = 0, 0, 0, 0
    before = _launches()
    with H.quiet():
        res = sim.test_multiple_rollouts_for_training(batch)
    res[0].backward()
    assert _launches() - before > 10 * steps
    for i, r_ in enumerate(res):
        want = float(g[f"out{i}"])
        assert abs(float(r_.detach()) - want) <= 2e-5 * max(abs(want), 1e-3), (i, float(r_.detach()), want)
    assert sim.collision_count == float(g["collision_count"])
    assert sim.hard_collision_count == float(g["hard_collision_count"])
    assert np.array_equal(batch.dest_idx.cpu().numpy(), g["dest_idx_after"])
    named = dict(sim.model.named_parameters())
    worst = 0.0
    for k, v in g.items():
        if k.startswith("grad/"):
            got = named[k[5:]].grad.cpu().numpy()
            worst = max(worst, float(np.abs(got - v).max() / max(np.abs(v).max(), 1e-12)))
    assert worst < 5e-5, worst
    dead = [n for n, p_ in named.items() if p_.grad is None]
    assert dead and all("processor" in n for n in dead)        # the discarded ResBlock Linear keeps grad=None (a6)


@pytest.mark.parametrize("clip,dsn", [(H.GC_CLIP, "gc1560"), (H.UCY_CLIP, "ucy")])
def test_make_dataset_whole_clip_matches_the_unpatched_reference(ref, clip, dsn):
    """TimeIndexedPedData.make_dataset over all 750 / 676 frames: patched (feature kernels with T as the batch
    dimension, desired-speed kernel, collision labels) vs the reference's own build on the host cores."""
    args = H.default_args(dataset_name=dsn)
    before = _launches()
    got = H.make_time_indexed(args, H.load_raw(clip))
    assert _launches() - before >= 4
    ref.patch.uninstall()
    try:
        want = H.make_time_indexed(args, H.load_raw(clip))
    finally:
        ref.patch.install(DATA=ref.DATA, MLAPM_MOD=ref.MLAPM, MODEL=ref.MODEL, SIM=ref.SIM, UTILS=ref.UTILS,
                          METRIC=ref.METRIC)
    for k in ("ped_features", "obs_features", "labels", "mask_p_pred", "mask_v_pred", "mask_a_pred", "abnormal_mask",
              "position", "velocity", "acceleration", "destination"):
        a, b = getattr(got, k), getattr(want, k)
        assert a.shape == b.shape and a.dtype == b.dtype, k
        assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)), k     # bit-exact
    a, b = got.self_features, want.self_features
    assert torch.equal(a[..., :6], b[..., :6])
    # desired speed: a mean of <= 25 fp32 norms; torch.mean's summation order is unspecified -> 1e-6 relative
    assert torch.allclose(a[..., 6], b[..., 6], rtol=1e-6, atol=0, equal_nan=True)
    for k in ("num_frames", "num_pedestrians", "ped_feature_dim", "obs_feature_dim", "self_feature_dim", "topk_obs"):
        assert getattr(got, k) == getattr(want, k), k
