"""Pins the CPU oracle (oracle/piml_oracle.c) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Bit-exact for neighbour sets, distances and features; 1e-5 for forces / MLP."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import accel_err, golden, group, rel_vec_err, untied_finite, valid_sets

FEATURE_CASES = ["gc", "gc_window", "ucy", "toy", "syn512", "syn300_wide", "channelled"]


@pytest.mark.parametrize("case", FEATURE_CASES)
def test_relative_features_bit_exact(case):
    g = group(golden("features"), case)
    kp, ap, tp, ko, ao, to = [int(v) for v in g["params"]]
    vel, acc = g["velocity"].copy(), g["acceleration"].copy()
    pf, of, df, (pi, pd, oi, od) = O.relative_features(g["position"], vel, acc, g["destination"], g["obstacles"],
                                                      kp, ap, tp, ko, ao, to, return_selection=True)
    assert np.array_equal(pf, g["ped_features"])
    assert np.array_equal(of, g["obs_features"])
    assert np.array_equal(df, g["dest_features"])
    # in-place NaN -> 0 side effect (data.py:483-484)
    assert np.array_equal(vel, g["velocity_after"]) and np.array_equal(acc, g["acceleration_after"])
    # distances are bit-equal everywhere; indices wherever the distance is finite (inf ties are unordered in torch)
    assert np.array_equal(pd, g["ped_dist"]) and np.array_equal(od, g["obs_dist"])
    fin = untied_finite(g["ped_dist"])
    assert np.array_equal(pi[fin], g["ped_idx"][fin].astype(np.int64))
    fin = untied_finite(g["obs_dist"])
    assert np.array_equal(oi[fin], g["obs_idx"][fin].astype(np.int64))
    assert valid_sets(pi, pd, tp) == valid_sets(g["ped_idx"], g["ped_dist"], tp)
    assert valid_sets(oi, od, to) == valid_sets(g["obs_idx"], g["obs_dist"], to)


@pytest.mark.parametrize("case", FEATURE_CASES)
def test_heading_bit_exact(case):
    g = group(golden("features"), case)
    assert np.array_equal(O.heading(g["velocity_after"]), g["heading"])


def test_mlapm_circle_rollout():
    g = group(golden("mlapm"), "circle")
    pos, vel, mask = g["position"], g["velocity"], g["mask"]
    worst = 0.0
    for t in range(pos.shape[1] - 1):
        m = mask[:, t]
        if not m.any():
            break
        act = O.mlapm_step(pos[m, t], vel[m, t], g["desired_speed"][m], g["destination"][m], 0.08, "GC")
        worst = max(worst, rel_vec_err(act, vel[m, t + 1]))
        # main_mlapm.py:26  p = position + v * dt  (fp32, un-fused) from the reference's own new velocity
        assert np.array_equal(pos[m, t] + vel[m, t + 1] * np.float32(0.08), pos[m, t + 1])
    assert worst < 1e-5, worst


@pytest.mark.parametrize("case,ver", [("syn257", "raw"), ("syn257", "GC"), ("syn1000", "raw"), ("syn1000", "GC")])
def test_mlapm_step(case, ver):
    g = group(golden("mlapm"), case)
    act = O.mlapm_step(g["position"], g["velocity"], g["desired_speed"], g["destination"], 0.08, ver)
    assert rel_vec_err(act, g[ver + "/action"]) < 1e-5


def test_mlapm_params_and_wide_speed():
    g = group(golden("mlapm"), "syn300b")
    tau, A, B, C_, D, th, dt = [float(v) for v in g["params"]]
    act = O.mlapm_step(g["position"], g["velocity"], g["desired_speed"], g["destination"], dt, "GC", tau, A, B, C_,
                       D, th)
    assert rel_vec_err(act, g["GC/action"]) < 1e-5


def test_mlapm_view_gate_bit_exact():
    """mlapm.py:27: einsum('nk,nmk->nm') > 0 == fmaf(v1, r1, v0*r0) > 0 for every ordered pair."""
    for case in ("syn257", "syn1000"):
        g = group(golden("mlapm"), case)
        p, v = g["position"], g["velocity"]
        n = p.shape[0]
        want = np.unpackbits(g["view_bits"])[:n * n].reshape(n, n).astype(bool)
        rx = (p[None, :, 0] - p[:, None, 0]).astype(np.float32)
        ry = (p[None, :, 1] - p[:, None, 1]).astype(np.float32)
        t = (v[:, None, 0] * rx).astype(np.float32)
        # fused multiply-add emulated in float64 (exact product + one rounding for fp32 operands)
        got = (v[:, None, 1].astype(np.float64) * ry.astype(np.float64) + t.astype(np.float64)).astype(np.float32) > 0
        assert np.array_equal(got, want)


@pytest.mark.parametrize("key", ["v0/gc1560", "v0/ucy", "v1/gc2344", "v1/ucy", "v2/gc2344"])
def test_calc_acceleration(key):
    z = golden("sfm")
    ver, ds = key.split("/")
    out = O.calc_acceleration(z["ped"], ver, ds)
    assert rel_vec_err(out, z[key]) < 1e-5
    assert np.array_equal(out == 0, z[key] == 0)          # padded slots stay exactly zero


def test_calc_acceleration_channelled():
    z = golden("sfm")
    assert rel_vec_err(O.calc_acceleration(z["ped_c"], "v2", "gc2344"), z["v2c/gc2344"]) < 1e-5
    assert rel_vec_err(O.calc_acceleration(z["ped_c"], "v0", "gc1560"), z["v0c/gc1560"]) < 1e-5


def _oracle_model(z, kind):
    import torch
    from piml_b200 import models as M
    g = group(z, kind)
    hs_e, hs_p, hs_d, nl_e, nl_p, nl_d, has_obs = [int(v) for v in g["cfg"]]
    kind_id, coll = M.MODEL_KINDS[kind]
    enc = [6] + [hs_e] * nl_e
    dec = [hs_p] + [hs_d] * nl_d
    coll_dims = [dec[-1], dec[-1], 1] if coll == 'dec' else ([hs_p, dec[-1], 1] if coll == 'proc' else [])
    spec = M.NetSpec(enc, 0 if nl_p > 1 else 1, dec, coll_dims, kind_id, bool(has_obs), float(g["tau"]))
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd/")}
    packed = M.pack_state_dict(sd, spec).numpy()
    desc = O.net_desc(spec.enc_dims, spec.proc_mode, spec.dec_dims, spec.coll_dims, spec.kind)
    return g, spec, desc, packed


@pytest.mark.parametrize("kind", ["pinnsf_bm", "pinnsf_m", "pinnsf_bottleneck", "pinnsf"])
def test_pinnsf_forward(kind):
    z = golden("models")
    g, spec, desc, packed = _oracle_model(z, kind)
    for suffix, chan in (("", False), ("c", True)):
        ped, obs, slf = z["ped" + ("_c" if chan else "")], z["obs" + ("_c" if chan else "")], \
            z["self" + ("_c" if chan else "")]
        out = O.pinnsf_forward(desc, packed, spec.tau, ped, obs, slf, spec.has_obs, channelled=chan)
        n_out = len([k for k in g if k.startswith("out" + suffix) and k[len("out" + suffix):].isdigit()])
        assert len(out) == n_out
        for i, o in enumerate(out):
            ref = g[f"out{suffix}{i}"]
            o = o.reshape(ref.shape)
            if o.ndim >= 2 and o.shape[-1] in (2, 128, 32):
                err = rel_vec_err(o, ref, floor=1e-3)
            else:
                err = float(np.abs(o - ref).max())
            assert err < 1e-5, (kind, suffix, i, err)


def test_integrate_step_matches_reference_rollout():
    """Re-synchronised single steps: from the reference's recorded state at t, one oracle update must
    reproduce the reference's p and v at t+1 bit-for-bit (a comes from the model and is taken from the golden)."""
    for name in ("rollout_gc_bm", "rollout_toy5_m", "rollout_ucy_bm"):
        z = golden(name)
        i, o = group(z, "in"), group(z, "out")
        T, t0, dt = int(i["num_frames"]), int(i["t_start"]), float(i["time_unit"])
        flag = (i["mask_p"] - i["mask_p_pred"]).astype(np.int64)
        P, V, A = o["position"], o["velocity"], o["acceleration"]
        dest = i["destination"][t0].copy()
        didx = i["dest_idx"][t0].astype(np.int64)
        for t in range(t0, T - 1):
            p, v, a, dest, didx, hist = O.integrate_step(
                P[t], V[t], A[t], A[t + 1], dest, didx, i["dest_num"], i["waypoints"], dt, True, flag[t + 1],
                i["position"][t + 1], i["velocity"][t + 1], i["acceleration"][t + 1], i["destination"][t + 1],
                i["dest_idx"][t + 1].astype(np.int64))
            assert np.array_equal(p, P[t + 1], equal_nan=True), (name, t)
            assert np.array_equal(v, V[t + 1]), (name, t)
            assert np.array_equal(dest, o["dest_after_step"][t - t0], equal_nan=True), (name, t)


@pytest.mark.parametrize("name", ["rollout_gc_bm", "rollout_toy5_m", "rollout_ucy_bm"])
def test_rollout_resynchronised_network_output(name):
    """Oracle features + oracle network on the reference's recorded state at t reproduce the reference's a[t+1]."""
    import torch
    from piml_b200 import models as M
    from tests.golden_args import base_args
    z = golden(name)
    i, o = group(z, "in"), group(z, "out")
    kind, dsn = str(i["model"]), str(i["dataset_name"])
    spec = M.spec_from_args(kind, base_args(model=kind, dataset_name=dsn))
    sd = {k[3:]: torch.from_numpy(v) for k, v in group(golden("models"), kind).items() if k.startswith("sd/")}
    packed = M.pack_state_dict(sd, spec).numpy()
    desc = O.net_desc(spec.enc_dims, spec.proc_mode, spec.dec_dims, spec.coll_dims, spec.kind)
    T, t0 = int(i["num_frames"]), int(i["t_start"])
    flag = i["mask_p"] - i["mask_p_pred"]
    worst = 0.0
    for t in range(t0 + 1, T - 1, max(1, (T - t0) // 40)):
        p, v, a = o["position"][t][None], o["velocity"][t][None].copy(), o["acceleration"][t][None].copy()
        dest = o["dest_after_step"][t - t0 - 1][None]
        pf, of, df = O.relative_features(p, v, a, dest, i["obstacles"])
        slf = np.concatenate([df[0], v[0], a[0], i["desired_speed"][:, None]], -1)
        acc = O.pinnsf_forward(desc, packed, spec.tau, pf[0], of[0], slf)[0]
        sim = flag[t + 1] == 0
        worst = max(worst, accel_err(acc[sim], o["acceleration"][t + 1][sim], slf[sim], spec.tau))
    assert worst < 1e-5, worst


# ---- pure social-force mode (BASELINE config 2) ------------------------------------------------------------------------
def test_sfm_forward_matches_composed_reference_module():
    g = group(golden("rollout_syn_sfm"), "fwd")
    acc, pm, om = O.sfm_forward(g["ped"], g["obs"], g["self"], "gc1560")
    assert rel_vec_err(pm, g["ped_msgs"], floor=1e-4) < 1e-5
    assert rel_vec_err(om, g["obs_msgs"], floor=1e-4) < 1e-5
    assert np.array_equal(pm == 0, g["ped_msgs"] == 0) and np.array_equal(om == 0, g["obs_msgs"] == 0)
    assert accel_err(acc, g["acc"], g["self"], 0.5) < 1e-5


def test_sfm_rollout_resynchronised():
    """From the reference's recorded state at t: oracle features + oracle social-force module reproduce a[t+1], and
    one oracle state update reproduces p, v at t+1 bit for bit (the composed module rolled out by the unmodified
    get_multiple_rollouts on the synthetic clip, 725 steps)."""
    z = golden("rollout_syn_sfm")
    i, o = group(z, "in"), group(z, "out")
    T, t0, dt = int(i["num_frames"]), int(i["t_start"]), float(i["time_unit"])
    assert str(i["model"]) == "sfm" and T == 750
    flag = (i["mask_p"] - i["mask_p_pred"]).astype(np.int64)
    P, V, A = o["position"], o["velocity"], o["acceleration"]
    dest = i["destination"][t0].copy()
    didx = i["dest_idx"][t0].astype(np.int64)
    worst = 0.0
    for t in range(t0, T - 1):
        p, v, a, dest, didx, hist = O.integrate_step(
            P[t], V[t], A[t], A[t + 1], dest, didx, i["dest_num"], i["waypoints"], dt, True, flag[t + 1],
            i["position"][t + 1], i["velocity"][t + 1], i["acceleration"][t + 1], i["destination"][t + 1],
            i["dest_idx"][t + 1].astype(np.int64))
        assert np.array_equal(p, P[t + 1], equal_nan=True), t
        assert np.array_equal(v, V[t + 1]), t
        assert np.array_equal(dest, o["dest_after_step"][t - t0], equal_nan=True), t
        if t > t0 and (t - t0) % 12 == 0:
            pp, vv, aa = P[t][None], V[t][None].copy(), A[t][None].copy()
            d_prev = o["dest_after_step"][t - t0 - 1][None]
            pf, of, df = O.relative_features(pp, vv, aa, d_prev, i["obstacles"])
            slf = np.concatenate([df[0], vv[0], aa[0], i["desired_speed"][:, None]], -1)
            acc = O.sfm_forward(pf[0], of[0], slf, "gc1560")[0]
            sim = flag[t + 1] == 0
            worst = max(worst, accel_err(acc[sim], A[t + 1][sim], slf[sim], 0.5))
    assert worst < 1e-5, worst


def test_predicate_arithmetic_against_torch_primitives():
    """the fp32 evaluation of the neighbour-selection predicate (torch.norm over a 2-vector,
    torch.cosine_similarity, the 3.14-based threshold; data.py:432-443) on ~10^6 random pairs incl. zero, huge, tiny,
    inf and NaN coordinates: the oracle's gated distances must equal torch's CPU kernels bit for bit."""
    import torch
    rng = np.random.default_rng(11)
    N, M = 256, 4096
    pos = rng.normal(0, 3, (1, N, 2)).astype(np.float32)
    obj = rng.normal(0, 3, (1, M, 2)).astype(np.float32)
    obj[0, :64] = pos[0, :64]                                  # coincident points (distance 0, cos of a zero vector)
    obj[0, 64:96] *= 1e18
    obj[0, 96:128] *= 1e-30
    obj[0, 128:140] = np.nan
    obj[0, 140:150, 0] = np.inf
    pos[0, -3:] = np.nan                                       # absent agents
    head = rng.normal(0, 1, (1, N, 2)).astype(np.float32)
    head[0, :16] = 0.0                                         # stationary agents see nobody at <= 90 degrees
    for angle in (90, 100):
        dist, idx = O.select(pos, obj, head, M, angle)
        p, o, h = torch.from_numpy(pos), torch.from_numpy(obj), torch.from_numpy(head)
        rel = o[:, None, :, :] - p[:, :, None, :]              # (1,N,M,2) = B - A
        rel[torch.isnan(rel)] = float('inf')
        d = torch.norm(rel, p=2, dim=-1)
        cos = torch.cosine_similarity(rel, h[:, :, None, :].expand_as(rel), dim=-1)
        cos[torch.isnan(cos)] = -1
        d[cos < torch.tensor(np.float32(np.cos(3.14 * angle / 180)))] = float('inf')
        want = torch.sort(d, dim=-1)[0].numpy()
        assert np.array_equal(np.asarray(dist).view(np.uint32), want.view(np.uint32)), angle
        fin = np.isfinite(want)
        assert fin.sum() > 1e5


# ---- f-4: evaluation metrics -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["rollout_gc_bm", "rollout_ucy_bm", "rollout_syn_sfm"])
def test_metrics_match_reference(name):
    """The numpy restatement of METRIC.mae / ot / mmd _with_time_mask against the reference's own values on its own
    rollouts (per frame where the reference loops over frames)."""
    g = group(golden("metrics"), name)
    p, q, mask = g["p_pred"], g["labels"], g["mask"]
    assert abs(O.mae_with_time_mask(p, q, mask) - float(g["mae_sum"])) <= 1e-5 * float(g["mae_sum"])
    step = 5 if name != "rollout_ucy_bm" else 2                  # every 5th frame keeps the CPU suite short
    sub = np.zeros_like(mask)
    sub[::step] = mask[::step]
    frames, ot = O.ot_with_time_mask(p, q, sub)
    idx = np.searchsorted(g["frames"], frames)
    assert np.array_equal(g["frames"][idx], frames)
    assert np.allclose(ot, g["ot"][idx], rtol=2e-4, atol=1e-6)
    frames2, mm = O.mmd_with_time_mask(p, q, sub)
    assert np.array_equal(frames2, frames)
    assert np.allclose(mm, g["mmd"][idx], rtol=2e-4, atol=2e-6)


@pytest.mark.parametrize("name", ["rollout_gc_bm", "rollout_ucy_bm", "rollout_toy5_m", "rollout_syn_sfm"])
def test_desired_speed_matches_reference_make_dataset(name):
    """data.py:797-806: the golden rollouts carry data.self_features[t, :, -1] of the reference's own make_dataset
    (skip_frames = 25) next to the clip's velocities."""
    g = golden(name)
    got = O.desired_speed(g["in/velocity"], 25)
    want = g["in/desired_speed"]
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-6, atol=0, equal_nan=True)
