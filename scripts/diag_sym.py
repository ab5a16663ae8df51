"""Diagnostic: symmetric vs ordered MLAPM kernels vs the fp32 oracle and an fp64 numpy evaluation."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import piml_b200 as P
from piml_b200 import _lib as L
from oracle import oracle as O
from tests.util import rel_vec_err

KW = dict(version='GC', tau=0.5, A=7.55, B=-3.00, C=0.2, D=-0.3, theta=56)


def f64(p, v, ds, d, dt, ver):
    p, v, ds, d = [x.astype(np.float64) for x in (p, v, ds, d)]
    N = len(p)
    e = d - p
    e /= np.maximum(np.linalg.norm(e, axis=-1, keepdims=True), 1e-12)
    F = (ds * e - v) / KW['tau']
    th0 = np.float64(np.float32(KW['theta']) / np.float32(180.0) * np.float32(3.14159274101257324))
    for a in range(0, N, 512):
        sl = slice(a, min(a + 512, N))
        vr = p[None, :, :] - p[sl, None, :]
        r = np.linalg.norm(vr, axis=-1)
        view = (np.einsum('nk,nmk->nm', v[sl].astype(np.float32), vr.astype(np.float32)) > 0)
        with np.errstate(all='ignore'):
            rh = vr / r[..., None]
            if ver == 'raw':
                w = KW['A'] * np.exp(KW['B'] * r)
                contrib = np.where(view[..., None], w[..., None] * rh, 0.0)
            else:
                vv = v[None, :, :] - v[sl, None, :]
                cs = (vr * vv).sum(-1) / np.maximum(r, 1e-8) / np.maximum(np.linalg.norm(vv, axis=-1), 1e-8)
                cross = (vr[..., 0].astype(np.float32) * e[sl, None, 1].astype(np.float32)
                         - vr[..., 1].astype(np.float32) * e[sl, None, 0].astype(np.float32))
                th = np.where(cross > 0, -th0, th0)
                c, s_ = np.cos(th), np.sin(th)
                dx = c * rh[..., 0] - s_ * rh[..., 1]
                dy = s_ * rh[..., 0] + c * rh[..., 1]
                w = KW['A'] * np.exp(KW['B'] * r + KW['C'] * cs + KW['D'] * r * cs)
                contrib = np.where(view[..., None], w[..., None] * np.stack([dx, dy], -1), 0.0)
        contrib[np.arange(sl.stop - sl.start), np.arange(sl.start, sl.stop)] = 0.0
        F[sl] -= contrib.sum(1)
    return v + F * dt


def main():
    for N, ver in ((8192, 'raw'), (8192, 'GC'), (5000, 'raw')):
        rng = np.random.default_rng(7 * N + len(ver))
        Ls = np.sqrt(N / 0.5)
        p = (rng.random((N, 2)) * Ls).astype(np.float32)
        d = (rng.random((N, 2)) * Ls).astype(np.float32)
        v = rng.normal(0, 1, (N, 2)).astype(np.float32)
        v[rng.random(N) < 0.05] = 0.0
        ds = (1.34 + 0.3 * rng.normal(0, 1, (N, 1))).astype(np.float32)
        want = O.mlapm_step(p, v, ds, d, 0.08, ver)
        ref64 = f64(p, v, ds, d, 0.08, ver)
        model = P.MLAPM(**dict(KW, version=ver))
        t = lambda x: torch.from_numpy(x).cuda()
        res = {}
        for algo, name in ((2, 'sym'), (1, 'ordered')):
            L.check(L.load().piml_set_mlapm_algorithm(algo), "x")
            res[name] = model.step(t(p), t(v), t(ds), t(d), 0.08).cpu().numpy()
        L.load().piml_set_mlapm_algorithm(0)
        print(N, ver, "sym-vs-oracle %.2e ordered-vs-oracle %.2e | vs f64: sym %.2e ordered %.2e oracle %.2e" % (
            rel_vec_err(res['sym'], want), rel_vec_err(res['ordered'], want), rel_vec_err(res['sym'], ref64),
            rel_vec_err(res['ordered'], ref64), rel_vec_err(want, ref64)))
        num = np.linalg.norm(res['sym'].astype(np.float64) - want, axis=-1)
        den = np.maximum(np.linalg.norm(want, axis=-1), 1e-6)
        i = int(np.argmax(num / den))
        print("  worst agent", i, "v", v[i], "want", want[i], "sym", res['sym'][i], "ordered", res['ordered'][i],
              "f64", ref64[i])


main()
