"""ctypes/numpy wrapper over oracle/liboracle.so (the C restatement in piml_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (piml_b200) never imports this module.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = os.path.join(_HERE, "piml_oracle.c")


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only; seconds)."""
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(_SRC)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


class NetDesc(C.Structure):
    """Mirror of orc_net_t."""
    _fields_ = [("n_enc", C.c_int), ("enc_dims", C.c_int * 9), ("proc_mode", C.c_int),
                ("n_dec", C.c_int), ("dec_dims", C.c_int * 9), ("n_coll", C.c_int),
                ("coll_dims", C.c_int * 5), ("kind", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_version.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def cos_threshold(angle):
    """data.py:442-443: math.cos(3.14 * angle / 180), cast to fp32 by the tensor comparison."""
    return np.float32(math.cos(3.14 * angle / 180))


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(int(n)))


def heading(vel):
    """vel (T,N,2) or (C,T,N,2) -> heading direction, same shape (data.py:351-395)."""
    v = _f32(vel)
    shp = v.shape
    Cc = shp[0] if v.ndim == 4 else 1
    T, N = shp[-3], shp[-2]
    out = np.empty_like(v)
    lib().orc_heading(_ptr(v), C.c_int(Cc), C.c_int(T), C.c_int(N), _ptr(out))
    return out


def select(pos, obj, head, k, angle):
    """get_nearby_obj_in_sight (data.py:416-447). pos (B,N,2), obj (B,M,2) or (M,2), head (B,N,2)."""
    pos, obj, head = _f32(pos), _f32(obj), _f32(head)
    B, N = pos.shape[0], pos.shape[1]
    M = obj.shape[-2]
    kk = min(k, M)
    stride = M * 2 if obj.ndim == 3 else 0
    dist = np.empty((B, N, kk), np.float32)
    idx = np.empty((B, N, kk), np.int64)
    lib().orc_select(_ptr(pos), _ptr(obj), C.c_int64(stride), _ptr(head), C.c_int(B), C.c_int(N), C.c_int(M),
                     C.c_int(k), C.c_float(cos_threshold(angle)), _ptr(dist), _ptr(idx))
    return dist, idx


def relative_features(position, velocity, acceleration, destination, obstacles, topk_ped=6, sight_angle_ped=90,
                      dist_threshold_ped=4, topk_obs=10, sight_angle_obs=90, dist_threshold_obs=4,
                      return_selection=False):
    """get_relative_features (data.py:466-512). Inputs (T,N,2) or (C,T,N,2); velocity/acceleration arrays that
    are writable fp32 C-contiguous are sanitised in place like the reference."""
    pos, dest = _f32(position), _f32(destination)
    vel = velocity if (isinstance(velocity, np.ndarray) and velocity.dtype == np.float32
                       and velocity.flags.c_contiguous and velocity.flags.writeable) else _f32(velocity).copy()
    acc = acceleration if (isinstance(acceleration, np.ndarray) and acceleration.dtype == np.float32
                           and acceleration.flags.c_contiguous and acceleration.flags.writeable) \
        else _f32(acceleration).copy()
    obs = _f32(obstacles)
    lead = pos.shape[:-2]
    Cc = pos.shape[0] if pos.ndim == 4 else 1
    T, N = pos.shape[-3], pos.shape[-2]
    M = obs.shape[-2] if obs.size else 0
    per_ch = 1 if obs.ndim == 3 else 0
    kp, ko = min(topk_ped, N), (min(topk_obs, M) if M else 0)
    ped_f = np.empty(lead + (N, kp, 6), np.float32)
    obs_f = np.empty(lead + (N, ko, 6), np.float32)
    dest_f = np.empty(lead + (N, 2), np.float32)
    pi = np.empty(lead + (N, kp), np.int64)
    pd = np.empty(lead + (N, kp), np.float32)
    oi = np.empty(lead + (N, ko), np.int64)
    od = np.empty(lead + (N, ko), np.float32)
    lib().orc_relative_features(
        _ptr(pos), _ptr(vel), _ptr(acc), _ptr(dest), _ptr(obs), C.c_int(per_ch), C.c_int(Cc), C.c_int(T),
        C.c_int(N), C.c_int(M), C.c_int(topk_ped), C.c_float(cos_threshold(sight_angle_ped)),
        C.c_float(dist_threshold_ped), C.c_int(topk_obs), C.c_float(cos_threshold(sight_angle_obs)),
        C.c_float(dist_threshold_obs), _ptr(ped_f), _ptr(obs_f), _ptr(dest_f), _ptr(pi), _ptr(pd), _ptr(oi),
        _ptr(od))
    if return_selection:
        return ped_f, obs_f, dest_f, (pi, pd, oi, od)
    return ped_f, obs_f, dest_f


def relative_features_rows(position, velocity, acceleration, destination, obstacles, rows, topk_ped=6,
                           sight_angle_ped=90, dist_threshold_ped=4, topk_obs=10, sight_angle_obs=90,
                           dist_threshold_obs=4, return_selection=False):
    """Rows [r0,r1) of get_relative_features for ONE frame: inputs (N,2) (or (1,N,2)); outputs carry r1-r0 rows.
    Same arithmetic as relative_features; for crowds whose full N x N scan would take minutes on the CPU."""
    pos, dest = _f32(position).reshape(-1, 2), _f32(destination).reshape(-1, 2)
    vel, acc = _f32(velocity).reshape(-1, 2).copy(), _f32(acceleration).reshape(-1, 2).copy()
    obs = _f32(obstacles).reshape(-1, 2)
    N, M = pos.shape[0], obs.shape[0]
    r0, r1 = rows
    R = r1 - r0
    kp, ko = min(topk_ped, N), (min(topk_obs, M) if M else 0)
    ped_f, obs_f = np.empty((R, kp, 6), np.float32), np.empty((R, ko, 6), np.float32)
    dest_f = np.empty((R, 2), np.float32)
    pi, pd = np.empty((R, kp), np.int64), np.empty((R, kp), np.float32)
    oi, od = np.empty((R, ko), np.int64), np.empty((R, ko), np.float32)
    lib().orc_relative_features_rows(
        _ptr(pos), _ptr(vel), _ptr(acc), _ptr(dest), _ptr(obs), C.c_int(N), C.c_int(M), C.c_int(topk_ped),
        C.c_float(cos_threshold(sight_angle_ped)), C.c_float(dist_threshold_ped), C.c_int(topk_obs),
        C.c_float(cos_threshold(sight_angle_obs)), C.c_float(dist_threshold_obs), C.c_int64(r0), C.c_int64(r1),
        _ptr(ped_f), _ptr(obs_f), _ptr(dest_f), _ptr(pi), _ptr(pd), _ptr(oi), _ptr(od))
    if return_selection:
        return ped_f, obs_f, dest_f, (pi, pd, oi, od)
    return ped_f, obs_f, dest_f


def desired_speed(velocity, skip_frames=25):
    """The double loop of TimeIndexedPedData.make_dataset (data.py:797-806): per pedestrian, mean ||v|| over the
    `skip_frames` frames from its first moving frame (frame 0 if it never moves).  velocity (T,N,2) -> (N,)."""
    v = _f32(velocity)
    T, N = v.shape[0], v.shape[1]
    speed = np.sqrt((v[..., 1] * v[..., 1] + v[..., 0] * v[..., 0]).astype(np.float32))       # norm2f
    out = np.empty(N, np.float32)
    for i in range(N):
        moving = np.nonzero(speed[:, i] > 0)[0]
        s = int(moving[0]) if len(moving) else 0
        out[i] = np.float32(speed[s:s + skip_frames, i].astype(np.float64).mean())
    return out


def collision_label(ped_f):
    f = _f32(ped_f)
    out = np.empty(f.shape[:-1], np.float32)
    lib().orc_collision_label(_ptr(f), C.c_int64(out.size), _ptr(out))
    return out


MLAPM_VERSIONS = {"raw": 0, "GC": 1}


def mlapm_step(position, velocity, desired_speed, destination, dt, version="GC", tau=0.5, A=7.55, B=-3.0,
               C_=0.2, D=-0.3, theta=56, rows=None):
    """MLAPM.step (mlapm.py:10-58). rows=(r0,r1) restricts the computed rows (bounded CPU-baseline samples)."""
    pos, vel, dest = _f32(position), _f32(velocity), _f32(destination)
    ds = _f32(desired_speed)
    if ds.ndim == 1:
        ds = ds[:, None]
    N = pos.shape[0]
    r0, r1 = rows if rows is not None else (0, N)
    out = np.empty((r1 - r0, 2), np.float32)
    lib().orc_mlapm_step(_ptr(pos), _ptr(vel), _ptr(ds), C.c_int(ds.shape[1]), _ptr(dest), C.c_int64(N),
                         C.c_int(MLAPM_VERSIONS[version]), C.c_float(tau), C.c_float(A), C.c_float(B),
                         C.c_float(C_), C.c_float(D), C.c_float(theta), C.c_float(dt), C.c_int64(r0),
                         C.c_int64(r1), _ptr(out))
    return out


def mlapm_step_diag(position, velocity, desired_speed, destination, dt, version="GC", tau=0.5, A=7.55, B=-3.0,
                    C_=0.2, D=-0.3, theta=56, rows=None):
    """mlapm_step plus, per row, the force before the Euler step (fp64, (R,2)) and the operand magnitude
    S = |dest term| + sum_m |pair term| (fp64, (R,)): returns (action, force, S)."""
    pos, vel, dest = _f32(position), _f32(velocity), _f32(destination)
    ds = _f32(desired_speed)
    if ds.ndim == 1:
        ds = ds[:, None]
    N = pos.shape[0]
    r0, r1 = rows if rows is not None else (0, N)
    out = np.empty((r1 - r0, 2), np.float32)
    force = np.empty((r1 - r0, 2), np.float64)
    opsum = np.empty((r1 - r0,), np.float64)
    lib().orc_mlapm_step_diag(_ptr(pos), _ptr(vel), _ptr(ds), C.c_int(ds.shape[1]), _ptr(dest), C.c_int64(N),
                              C.c_int(MLAPM_VERSIONS[version]), C.c_float(tau), C.c_float(A), C.c_float(B),
                              C.c_float(C_), C.c_float(D), C.c_float(theta), C.c_float(dt), C.c_int64(r0),
                              C.c_int64(r1), _ptr(out), _ptr(force), _ptr(opsum))
    return out, force, opsum


_SFM = {("v0", "gc1560"): (8.75, -2.5, 0, 0, 0), ("v0", "gc2344"): (8.75, -2.5, 0, 0, 0),
        ("v0", "ucy"): (10.67, -3.33, 0, 0, 0),
        ("v1", "gc1560"): (8.75, -2.5, 0, 0, 0), ("v1", "gc2344"): (8.75, -2.5, 0, 0, 0),
        ("v1", "ucy"): (10.67, -3.33, 0, 0, 0),
        ("v2", "gc2344"): (9.00, -2.75, 0.06, -0.3, 10 * 3.1415 / 180)}


def calc_acceleration(relative_data, equation_version="v0", dataset="gc1560", eps=1e-6):
    """UTILS.calc_acceleration (utils.py:31-100)."""
    rel = _f32(relative_data)
    A, B, Cc, D, th = _SFM[(equation_version, dataset)]
    out = np.empty(rel.shape[:-1] + (2,), np.float32)
    lib().orc_calc_acceleration(_ptr(rel), C.c_int64(out.size // 2), C.c_int(rel.shape[-1]),
                                C.c_int(int(equation_version[1])), C.c_float(A), C.c_float(B), C.c_float(Cc),
                                C.c_float(D), C.c_float(th), C.c_float(eps), _ptr(out))
    return out


SFM_CONSTS = {"gc1560": (8.75, -2.5), "gc2344": (8.75, -2.5), "ucy": (10.67, -3.33)}     # utils.py:47-52


def sfm_forward(ped, obs, slf, dataset="gc1560", tau=0.5, A_obs=10 / 0.2, B_obs=-1 / 0.2, eps=1e-6):
    """The composed social-force module (tests/golden/make_golden.py: SocialForceComposed) for (N,..) inputs.
    Returns [acc (N,2), ped_msgs (N,kp,2), obs_msgs (N,ko,2)]."""
    ped, slf = _f32(ped), _f32(slf)
    obs = _f32(obs) if obs is not None and np.asarray(obs).size else None
    R, kp = ped.shape[0], ped.shape[1]
    ko = obs.shape[1] if obs is not None else 0
    A, B = SFM_CONSTS[dataset]
    acc = np.empty((R, 2), np.float32)
    pm = np.empty((R, kp, 2), np.float32)
    om = np.empty((R, ko, 2), np.float32)
    lib().orc_sfm_forward(_ptr(ped), _ptr(obs), _ptr(slf), C.c_int64(R), C.c_int(kp), C.c_int(ko), C.c_float(A),
                          C.c_float(B), C.c_float(A_obs), C.c_float(B_obs), C.c_float(eps), C.c_float(tau), _ptr(acc),
                          _ptr(pm), _ptr(om) if ko else None)
    return [acc, pm, om]


def net_desc(enc_dims, proc_mode, dec_dims, coll_dims, kind):
    d = NetDesc()
    d.n_enc = len(enc_dims) - 1
    for i, v in enumerate(enc_dims):
        d.enc_dims[i] = v
    d.proc_mode = proc_mode
    d.n_dec = len(dec_dims) - 1
    for i, v in enumerate(dec_dims):
        d.dec_dims[i] = v
    d.n_coll = max(len(coll_dims) - 1, 0)
    for i, v in enumerate(coll_dims):
        d.coll_dims[i] = v
    d.kind = kind
    return d


def pinnsf_forward(desc, params, tau, ped, obs, self_f, has_obs=True, channelled=False):
    """Forward of a PINNSF-family model (eval mode). `desc` from net_desc, `params` the packed fp32 vector.
    Returns [acc, ped_msgs, (obs_msgs), (coll)] like the reference's list."""
    ped, self_f = _f32(ped), _f32(self_f)
    obs = _f32(obs) if has_obs else np.zeros(ped.shape[:-2] + (0, 6), np.float32)
    params = _f32(params)
    lead = self_f.shape[:-1]
    R = int(np.prod(lead))
    kp, ko = ped.shape[-2], obs.shape[-2]
    pw = desc.enc_dims[desc.n_enc]
    msgw = 2 if desc.kind == 0 else pw
    acc = np.empty(lead + (2,), np.float32)
    pm = np.empty(lead + (kp, msgw), np.float32)
    om = np.empty(lead + (ko, msgw), np.float32)
    coll = np.empty(lead + (kp,), np.float32)
    group = lead[-1] if (channelled and len(lead) >= 2) else 0
    lib().orc_pinnsf_forward(C.byref(desc), _ptr(params), C.c_int(1 if has_obs else 0), C.c_float(tau),
                             _ptr(ped), _ptr(obs), _ptr(self_f), C.c_int64(R), C.c_int(kp), C.c_int(ko),
                             C.c_int(group), _ptr(acc), _ptr(pm), _ptr(om), _ptr(coll))
    out = [acc, pm]
    if has_obs:
        out.append(om)
    if desc.n_coll:
        out.append(coll)
    return out


def integrate_step(p, v, a, a_next, dest, dest_idx, dest_num, waypoints, dt, remove_on_arrival=True,
                   entry=None, p_gt=None, v_gt=None, a_gt=None, dest_gt=None, dest_idx_gt=None):
    """simulators.py:603-639 on ONE scene. Returns new (p, v, a, dest, dest_idx, hist_v) (inputs untouched)."""
    p, v, a, dest = _f32(p).copy(), _f32(v).copy(), _f32(a).copy(), _f32(dest).copy()
    a_next, wp = _f32(a_next), _f32(waypoints)
    di = np.ascontiguousarray(dest_idx, dtype=np.int64).copy()
    dn = np.ascontiguousarray(dest_num, dtype=np.int64)
    N = p.shape[0]
    hist = np.empty((N, 2), np.float32)
    if entry is not None:
        en = np.ascontiguousarray(entry, dtype=np.int64)
        gts = [_f32(p_gt), _f32(v_gt), _f32(a_gt), _f32(dest_gt)]
        dig = np.ascontiguousarray(dest_idx_gt, dtype=np.int64)
    else:
        en, gts, dig = None, [None] * 4, None
    lib().orc_integrate_step(_ptr(p), _ptr(v), _ptr(a), _ptr(a_next), _ptr(dest), _ptr(di), _ptr(dn), _ptr(wp),
                             C.c_int(N), C.c_float(dt), C.c_int(1 if remove_on_arrival else 0), _ptr(en),
                             _ptr(gts[0]), _ptr(gts[1]), _ptr(gts[2]), _ptr(gts[3]), _ptr(dig), _ptr(hist))
    return p, v, a, dest, di, hist


# ---- f-4: evaluation metrics (reference src/functions/metrics.py), numpy fp32 restatement -----------------------------
def mae_with_time_mask(p, q, mask):
    """metrics.py:29-42 with reduction 'sum': sum over mask == 1 of ||p - q||_2."""
    p, q, mask = _f32(p), _f32(q), np.asarray(mask)
    d = (p - q)[mask == 1]
    return float(np.sqrt((d * d).sum(-1, dtype=np.float32)).sum(dtype=np.float32))


def _logsumexp(x, axis):
    m = x.max(axis=axis, keepdims=True)
    return (np.log(np.exp(x - m).sum(axis=axis, dtype=np.float32)) + np.squeeze(m, axis)).astype(np.float32)


def sinkhorn_distance(x, y, eps=0.1, max_iter=100):
    """SinkhornDistance.forward (metrics.py:131-192) for one frame: x (n,2), y (n,2) -> entropic OT cost.
    Log-domain updates, equal weights, stops when sum |u - u_prev| < 0.1."""
    x, y = _f32(x), _f32(y)
    eps32 = np.float32(eps)
    C = (np.abs(x[:, None, :] - y[None, :, :]) ** 2).sum(-1, dtype=np.float32)            # :194-199
    n, m = x.shape[0], y.shape[0]
    lmu = np.log(np.float32(1.0 / n) + np.float32(1e-8), dtype=np.float32)
    lnu = np.log(np.float32(1.0 / m) + np.float32(1e-8), dtype=np.float32)
    u, v = np.zeros(n, np.float32), np.zeros(m, np.float32)
    M = lambda: ((-C + u[:, None]) + v[None, :]) / eps32                                 # :186-189
    for _ in range(max_iter):
        u1 = u
        u = eps32 * (lmu - _logsumexp(M(), 1)) + u
        v = eps32 * (lnu - _logsumexp(M().T, 1)) + v
        if float(np.abs(u - u1).sum(dtype=np.float32)) < 1e-1:                            # :166-170
            break
    pi = np.exp(M())
    return float((pi * C).sum(dtype=np.float32))


def ot_with_time_mask(p, q, mask, eps=0.1, max_iter=100):
    """metrics.py:45-67: one Sinkhorn problem per frame with more than one masked point; returns (frames, costs)."""
    p, q, mask = _f32(p), _f32(q), np.asarray(mask)
    frames, out = [], []
    for t in range(mask.shape[0]):
        sel = mask[t] == 1
        if sel.sum() > 1:
            frames.append(t)
            out.append(sinkhorn_distance(p[t][sel], q[t][sel], eps, max_iter))
    return np.asarray(frames), np.asarray(out, np.float64)


def mmd(source, target, kernel_mul=2.0, kernel_num=5):
    """MaximumMeanDiscrepancy.__call__ (metrics.py:213-273): multi-bandwidth Gaussian-kernel MMD of two point sets."""
    s, t = _f32(source), _f32(target)
    n, m = s.shape[0], t.shape[0]
    tot = np.concatenate([s, t], 0)
    L2 = ((tot[None, :, :] - tot[:, None, :]) ** 2).sum(2, dtype=np.float32)
    ns = n + m
    bw = np.float32(L2.sum(dtype=np.float32) / np.float32(ns * ns - ns))
    bw = np.float32(bw / np.float32(kernel_mul ** (kernel_num // 2)))
    K = np.zeros_like(L2)
    for i in range(kernel_num):
        with np.errstate(divide='ignore', invalid='ignore'):
            K = K + np.exp(-L2 / np.float32(bw * np.float32(kernel_mul ** i)))
    XX = (K[:n, :n] / np.float32(n * n)).sum(1, dtype=np.float32)
    XY = (K[:n, n:] / np.float32(-n * m)).sum(1, dtype=np.float32)
    YX = (K[n:, :n] / np.float32(-m * n)).sum(1, dtype=np.float32)
    YY = (K[n:, n:] / np.float32(m * m)).sum(1, dtype=np.float32)
    return float((XX + XY).sum(dtype=np.float32) + (YX + YY).sum(dtype=np.float32))


def mmd_with_time_mask(p, q, mask):
    """metrics.py:70-91; returns (frames, values)."""
    p, q, mask = _f32(p), _f32(q), np.asarray(mask)
    frames, out = [], []
    for t in range(mask.shape[0]):
        sel = mask[t] == 1
        if sel.sum() > 1:
            frames.append(t)
            out.append(mmd(p[t][sel], q[t][sel]))
    return np.asarray(frames), np.asarray(out, np.float64)
