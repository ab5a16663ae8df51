/*
 * piml_oracle.c -- CPU restatement of the PIML crowd-rollout hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity ORACLE for the CUDA path in piml_b200/csrc.  Only tests/, the smoke() check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (the piml_b200 package) never imports it and has no CPU fallback.
 *
 * Every function restates one reference function (file:line under /root/reference/src) in plain C with the
 * exact fp32 evaluation order torch 2.11's CPU kernels use:
 *   torch.norm over a 2-vector          -> sqrtf(fmaf(y, y, x*x))
 *   torch.cosine_similarity (2-vectors) -> normalise each operand by max(norm, eps), then a0*b0 + a1*b1 (no FMA)
 *   einsum('nk,nmk->nm') (K=2, bmm)     -> fmaf(v1, r1, v0*r0)
 * Compile with -ffp-contract=off (see oracle/Makefile) so that only the explicit fmaf() calls fuse.
 * Pinned against golden vectors generated from the unmodified reference (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

static inline float norm2f(float x, float y) { return sqrtf(fmaf(y, y, x * x)); }

ORC_API int orc_version(void) { return 1; }

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* data.py:483-484  acceleration[isnan]=0 ; velocity[isnan]=0   (in place on the caller's tensors) */
ORC_API void orc_nan_to_zero(float *x, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        if (isnan(x[i])) x[i] = 0.0f;
}

/* data.py:351-395  get_heading_direction.  vel: (C,T,N,2) -> out (C,T,N,2).
 * Zero-speed frames take the nearest LATER non-zero velocity of the same pedestrian, else the nearest earlier
 * one (backward pass then forward pass, :366-377); then v / ||v|| with ||v||==0 -> 0.1 (:391-394). */
ORC_API void orc_heading(const float *vel, int C, int T, int N, float *out) {
    memcpy(out, vel, sizeof(float) * (size_t)C * T * N * 2);
    for (int c = 0; c < C; ++c)
        for (int i = 0; i < N; ++i) {
            float tx = 0.f, ty = 0.f;
            for (int t = T - 1; t >= 0; --t) {
                float *h = out + (((size_t)c * T + t) * N + i) * 2;
                if (norm2f(h[0], h[1]) == 0.f) { h[0] = tx; h[1] = ty; } else { tx = h[0]; ty = h[1]; }
            }
            for (int t = 0; t < T; ++t) {
                float *h = out + (((size_t)c * T + t) * N + i) * 2;
                if (norm2f(h[0], h[1]) == 0.f) { h[0] = tx; h[1] = ty; } else { tx = h[0]; ty = h[1]; }
            }
        }
    size_t tot = (size_t)C * T * N;
    for (size_t j = 0; j < tot; ++j) {
        float n = norm2f(out[2 * j], out[2 * j + 1]);
        if (n == 0.f) n += 0.1f;
        out[2 * j] = out[2 * j] / n;
        out[2 * j + 1] = out[2 * j + 1] / n;
    }
}

/* Distance with the field-of-view gate applied, data.py:432-443.
 * rel = obj - pos ; NaN -> +inf ; dist = ||rel|| ; cos = cosine_similarity(rel, heading) (eps 1e-8), NaN -> -1 ;
 * dist = inf where cos < thr. */
static inline float gated_distance(float px, float py, float ox, float oy, float hx, float hy, float cos_thr) {
    float rx = ox - px, ry = oy - py;
    if (isnan(rx)) rx = INFINITY;
    if (isnan(ry)) ry = INFINITY;
    float d = norm2f(rx, ry);
    float nr = fmaxf(d, 1e-8f);
    float nh = fmaxf(norm2f(hx, hy), 1e-8f);
    float a = (rx / nr) * (hx / nh);
    float b = (ry / nr) * (hy / nh);
    float c = a + b;
    if (isnan(c)) c = -1.0f;
    if (c < cos_thr) d = INFINITY;
    return d;
}

typedef struct { float d; int64_t i; } cand_t;

/* (dist, index) lexicographic = torch.sort's CPU behaviour (stable ascending), data.py:445 */
static inline int cand_less(cand_t a, cand_t b) { return a.d < b.d || (a.d == b.d && a.i < b.i); }

/* keep the kk smallest candidates sorted in best[0..cnt) */
static inline void topk_insert(cand_t *best, int *cnt, int kk, cand_t c) {
    int n = *cnt;
    if (n == kk && !cand_less(c, best[n - 1])) return;
    int j = (n < kk) ? n : n - 1;
    while (j > 0 && cand_less(c, best[j - 1])) { best[j] = best[j - 1]; --j; }
    best[j] = c;
    if (n < kk) *cnt = n + 1;
}

/* data.py:416-447  get_nearby_obj_in_sight.
 * pos (B,N,2), obj (B,M,2) or (M,2) when obj_stride==0 (stride in floats between frames), head (B,N,2).
 * out_dist (B,N,kk) , out_idx (B,N,kk) with kk = min(k, M). */
ORC_API void orc_select(const float *pos, const float *obj, int64_t obj_stride, const float *head, int B, int N,
                        int M, int k, float cos_thr, float *out_dist, int64_t *out_idx) {
    int kk = k < M ? k : M;
    if (kk <= 0) return;
#pragma omp parallel
    {
        cand_t *best = (cand_t *)malloc(sizeof(cand_t) * kk);
#pragma omp for schedule(static)
        for (int64_t row = 0; row < (int64_t)B * N; ++row) {
            int b = (int)(row / N);
            const float *p = pos + row * 2;
            const float *h = head + row * 2;
            const float *o = obj + (size_t)b * obj_stride;
            int cnt = 0;
            for (int m = 0; m < M; ++m) {
                cand_t c;
                c.d = gated_distance(p[0], p[1], o[2 * m], o[2 * m + 1], h[0], h[1], cos_thr);
                c.i = m;
                topk_insert(best, &cnt, kk, c);
            }
            for (int j = 0; j < kk; ++j) {
                out_dist[row * kk + j] = best[j].d;
                out_idx[row * kk + j] = best[j].i;
            }
        }
        free(best);
    }
}

/* data.py:466-512  get_relative_features (with :449-464 get_filtered_features and :398-414 folded in:
 * gathering the k selected rows is bit-equivalent to materialise-then-gather, SURVEY.md App. C spec_check).
 * pos/vel/acc/dest: (C,T,N,2); vel and acc are sanitised IN PLACE (NaN->0) like the reference.
 * obs: (M,2) if obs_per_channel==0 else (C,M,2).   Outputs:
 *   ped_f (C,T,N,kp',6)  obs_f (C,T,N,ko',6)  dest_f (C,T,N,2)   kp'=min(kp,N) ko'=min(ko,M)
 *   optional (may be NULL): ped_idx/ped_dist (C,T,N,kp'), obs_idx/obs_dist (C,T,N,ko'). */
ORC_API void orc_relative_features(const float *pos, float *vel, float *acc, const float *dest, const float *obs,
                                   int obs_per_channel, int C, int T, int N, int M, int kp, float cos_p,
                                   float thr_p, int ko, float cos_o, float thr_o, float *ped_f, float *obs_f,
                                   float *dest_f, int64_t *ped_idx, float *ped_dist, int64_t *obs_idx,
                                   float *obs_dist) {
    int64_t B = (int64_t)C * T;
    int64_t rows = B * N;
    orc_nan_to_zero(acc, rows * 2);
    orc_nan_to_zero(vel, rows * 2);
    float *head = (float *)malloc(sizeof(float) * rows * 2);
    orc_heading(vel, C, T, N, head);

    int kpp = kp < N ? kp : N;
    int kop = (M > 0) ? (ko < M ? ko : M) : 0;
    float *pd = (float *)malloc(sizeof(float) * rows * (kpp > 0 ? kpp : 1));
    int64_t *pi = (int64_t *)malloc(sizeof(int64_t) * rows * (kpp > 0 ? kpp : 1));
    float *od = (float *)malloc(sizeof(float) * rows * (kop > 0 ? kop : 1));
    int64_t *oi = (int64_t *)malloc(sizeof(int64_t) * rows * (kop > 0 ? kop : 1));

    /* ped-ped: objects = positions of the same frame (data.py:489-490) */
    orc_select(pos, pos, (int64_t)N * 2, head, (int)B, N, N, kp, cos_p, pd, pi);
    /* ped-obstacle (data.py:504-505): obstacles broadcast over time */
    if (M > 0) {
        if (!obs_per_channel) {
            orc_select(pos, obs, 0, head, (int)B, N, M, ko, cos_o, od, oi);
        } else {
            for (int c = 0; c < C; ++c)
                orc_select(pos + (size_t)c * T * N * 2, obs + (size_t)c * M * 2, 0, head + (size_t)c * T * N * 2, T,
                           N, M, ko, cos_o, od + (size_t)c * T * N * kop, oi + (size_t)c * T * N * kop);
        }
    }

#pragma omp parallel for schedule(static)
    for (int64_t row = 0; row < rows; ++row) {
        int64_t b = row / N;
        int c = (int)(b / T);
        const float *p = pos + row * 2, *v = vel + row * 2, *a = acc + row * 2;
        const float *fp = pos + b * N * 2, *fv = vel + b * N * 2, *fa = acc + b * N * 2;
        for (int j = 0; j < kpp; ++j) {
            float *f = ped_f + (row * kpp + j) * 6;
            int64_t m = pi[row * kpp + j];
            if (pd[row * kpp + j] > thr_p) {        /* data.py:459-462 zero padding */
                for (int q = 0; q < 6; ++q) f[q] = 0.f;
            } else {                                  /* data.py:412 relative = B - A */
                f[0] = fp[2 * m] - p[0]; f[1] = fp[2 * m + 1] - p[1];
                f[2] = fv[2 * m] - v[0]; f[3] = fv[2 * m + 1] - v[1];
                f[4] = fa[2 * m] - a[0]; f[5] = fa[2 * m + 1] - a[1];
            }
        }
        /* data.py:496-497 */
        float dx = dest[row * 2] - p[0], dy = dest[row * 2 + 1] - p[1];
        dest_f[row * 2] = isnan(dx) ? 0.f : dx;
        dest_f[row * 2 + 1] = isnan(dy) ? 0.f : dy;
        const float *o = obs + (obs_per_channel ? (size_t)c * M * 2 : 0);
        for (int j = 0; j < kop; ++j) {
            float *f = obs_f + (row * kop + j) * 6;
            int64_t m = oi[row * kop + j];
            if (od[row * kop + j] > thr_o) {
                for (int q = 0; q < 6; ++q) f[q] = 0.f;
            } else {                                  /* data.py:506-508 obs = (o, 0, 0) */
                f[0] = o[2 * m] - p[0]; f[1] = o[2 * m + 1] - p[1];
                f[2] = 0.f - v[0]; f[3] = 0.f - v[1];
                f[4] = 0.f - a[0]; f[5] = 0.f - a[1];
            }
        }
    }
    if (ped_idx) memcpy(ped_idx, pi, sizeof(int64_t) * rows * kpp);
    if (ped_dist) memcpy(ped_dist, pd, sizeof(float) * rows * kpp);
    if (obs_idx && kop) memcpy(obs_idx, oi, sizeof(int64_t) * rows * kop);
    if (obs_dist && kop) memcpy(obs_dist, od, sizeof(float) * rows * kop);
    free(head); free(pd); free(pi); free(od); free(oi);
}

/* Row-range form of orc_relative_features for ONE frame (C = T = 1): rows [row0,row1) of the frame against all N
 * agents / M obstacles -- the same arithmetic, used to check large crowds (N = 1e5) on a bounded row sample.
 * vel / acc are sanitised in place over the WHOLE frame like the reference (data.py:483-484); outputs hold
 * (row1-row0) rows. */
ORC_API void orc_relative_features_rows(const float *pos, float *vel, float *acc, const float *dest,
                                        const float *obs, int N, int M, int kp, float cos_p, float thr_p, int ko,
                                        float cos_o, float thr_o, int64_t row0, int64_t row1, float *ped_f,
                                        float *obs_f, float *dest_f, int64_t *ped_idx, float *ped_dist,
                                        int64_t *obs_idx, float *obs_dist) {
    int64_t R = row1 - row0;
    orc_nan_to_zero(acc, (int64_t)N * 2);
    orc_nan_to_zero(vel, (int64_t)N * 2);
    float *head = (float *)malloc(sizeof(float) * (size_t)N * 2);
    orc_heading(vel, 1, 1, N, head);
    int kpp = kp < N ? kp : N;
    int kop = (M > 0) ? (ko < M ? ko : M) : 0;
    float *pd = (float *)malloc(sizeof(float) * R * (kpp > 0 ? kpp : 1));
    int64_t *pi = (int64_t *)malloc(sizeof(int64_t) * R * (kpp > 0 ? kpp : 1));
    float *od = (float *)malloc(sizeof(float) * R * (kop > 0 ? kop : 1));
    int64_t *oi = (int64_t *)malloc(sizeof(int64_t) * R * (kop > 0 ? kop : 1));
    orc_select(pos + row0 * 2, pos, 0, head + row0 * 2, 1, (int)R, N, kp, cos_p, pd, pi);
    if (M > 0) orc_select(pos + row0 * 2, obs, 0, head + row0 * 2, 1, (int)R, M, ko, cos_o, od, oi);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < R; ++r) {
        int64_t row = row0 + r;
        const float *p = pos + row * 2, *v = vel + row * 2, *a = acc + row * 2;
        for (int j = 0; j < kpp; ++j) {
            float *f = ped_f + (r * kpp + j) * 6;
            int64_t m = pi[r * kpp + j];
            if (pd[r * kpp + j] > thr_p) {
                for (int q = 0; q < 6; ++q) f[q] = 0.f;
            } else {
                f[0] = pos[2 * m] - p[0]; f[1] = pos[2 * m + 1] - p[1];
                f[2] = vel[2 * m] - v[0]; f[3] = vel[2 * m + 1] - v[1];
                f[4] = acc[2 * m] - a[0]; f[5] = acc[2 * m + 1] - a[1];
            }
        }
        float dx = dest[row * 2] - p[0], dy = dest[row * 2 + 1] - p[1];
        dest_f[r * 2] = isnan(dx) ? 0.f : dx;
        dest_f[r * 2 + 1] = isnan(dy) ? 0.f : dy;
        for (int j = 0; j < kop; ++j) {
            float *f = obs_f + (r * kop + j) * 6;
            int64_t m = oi[r * kop + j];
            if (od[r * kop + j] > thr_o) {
                for (int q = 0; q < 6; ++q) f[q] = 0.f;
            } else {
                f[0] = obs[2 * m] - p[0]; f[1] = obs[2 * m + 1] - p[1];
                f[2] = 0.f - v[0]; f[3] = 0.f - v[1];
                f[4] = 0.f - a[0]; f[5] = 0.f - a[1];
            }
        }
    }
    if (ped_idx) memcpy(ped_idx, pi, sizeof(int64_t) * R * kpp);
    if (ped_dist) memcpy(ped_dist, pd, sizeof(float) * R * kpp);
    if (obs_idx && kop) memcpy(obs_idx, oi, sizeof(int64_t) * R * kop);
    if (obs_dist && kop) memcpy(obs_dist, od, sizeof(float) * R * kop);
    free(head); free(pd); free(pi); free(od); free(oi);
}

/* data.py:515-535 calculate_collision_label: will the pair come within 0.5 m in the next second (10 samples) */
ORC_API void orc_collision_label(const float *ped_f, int64_t slots, float *out) {
    for (int64_t s = 0; s < slots; ++s) {
        const float *f = ped_f + s * 6;
        float acc = 0.f;
        for (int q = 0; q < 10; ++q) {
            float tq = (float)q * 0.1f;             /* torch.arange(10) * 0.1 : int64 -> fp32 product */
            float x = f[0] + f[2] * tq, y = f[1] + f[3] * tq;
            float d = norm2f(x, y);
            float cflag = (d < 0.5f && d != 0.f) ? 1.f : 0.f;
            acc += cflag;
        }
        out[s] = acc > 0.f ? 1.f : 0.f;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * MLAPM.step, mlapm.py:10-58.  version: 0 = 'raw', 1 = 'GC'.  ('UCY' crashes in the reference, SURVEY B-15.)
 * pos, vel, dest (N,2); ds (N,ds_dim) with ds_dim 1 or 2 (main_mlapm.py:13 passes (N,2)).
 * Computes rows [row0,row1) against ALL N columns; out (row1-row0, 2) = velocity + force*dt.
 * Per-pair terms follow the reference's fp32 op order; the row sum is accumulated in double (the reference's
 * fp32 .sum(dim=1) order is an ATen implementation detail; it differs from this by ~1e-6).
 * NaN semantics: view(bool)*A*exp(..)*direc is a product, so a NaN column poisons every row (0*NaN = NaN). */
static void mlapm_step_impl(const float *pos, const float *vel, const float *ds, int ds_dim, const float *dest,
                            int64_t N, int version, float tau, float A, float Bc, float Cc, float Dc,
                            float theta_deg, float dt, int64_t row0, int64_t row1, float *out, double *force,
                            double *opsum) {
    /* theta = -sign(cross) * theta/180*pi as fp32 tensor ops: (sign*theta)/180*pi ; theta==0 -> theta/180*pi (double->fp32) */
    const float pi_f = (float)3.141592653589793;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t n = row0; n < row1; ++n) {
        float px = pos[2 * n], py = pos[2 * n + 1], vx = vel[2 * n], vy = vel[2 * n + 1];
        /* F.normalize(destination - position): x / max(||x||, 1e-12) */
        float ex = dest[2 * n] - px, ey = dest[2 * n + 1] - py;
        float en = fmaxf(norm2f(ex, ey), 1e-12f);
        ex = ex / en; ey = ey / en;
        float dsx = ds[n * ds_dim], dsy = ds[n * ds_dim + (ds_dim > 1 ? 1 : 0)];
        float fx0 = (dsx * ex - vx) / tau, fy0 = (dsy * ey - vy) / tau;
        double sx = 0.0, sy = 0.0, mag = 0.0;
        for (int64_t m = 0; m < N; ++m) {
            float rx = pos[2 * m] - px, ry = pos[2 * m + 1] - py;
            float r = norm2f(rx, ry);
            float viewf = (fmaf(vy, ry, vx * rx) > 0.f) ? 1.f : 0.f;
            float nr = fmaxf(r, 1e-12f);
            float nx = rx / nr, ny = ry / nr;
            float tx, ty;
            if (version == 0) {
                float e = expf(Bc * r);
                tx = ((viewf * A) * e) * nx;
                ty = ((viewf * A) * e) * ny;
            } else {
                float ux = vel[2 * m] - vx, uy = vel[2 * m + 1] - vy;
                float cr = fmaxf(r, 1e-8f), cu = fmaxf(norm2f(ux, uy), 1e-8f);
                float cosv = (rx / cr) * (ux / cu) + (ry / cr) * (uy / cu);
                float cross = rx * ey - ry * ex;
                float sg = (cross > 0.f) ? 1.f : ((cross < 0.f) ? -1.f : (isnan(cross) ? NAN : 0.f));
                float th = ((-sg) * theta_deg) / 180.f * pi_f;
                if (th == 0.f) th = (float)((double)theta_deg / 180.0 * 3.141592653589793);
                float Ct = cosf(th), St = sinf(th);
                float dxr = Ct * nx + (-St) * ny;
                float dyr = St * nx + Ct * ny;
                float e = expf(Bc * r + Cc * cosv + (Dc * r) * cosv);
                tx = ((viewf * A) * e) * dxr;
                ty = ((viewf * A) * e) * dyr;
            }
            sx += (double)tx; sy += (double)ty;
            if (opsum) mag += sqrt((double)tx * tx + (double)ty * ty);
        }
        float fx = fx0 - (float)sx, fy = fy0 - (float)sy;
        out[2 * (n - row0)] = vx + fx * dt;
        out[2 * (n - row0) + 1] = vy + fy * dt;
        if (force) { force[2 * (n - row0)] = (double)fx0 - sx; force[2 * (n - row0) + 1] = (double)fy0 - sy; }
        if (opsum) opsum[n - row0] = mag + sqrt((double)fx0 * fx0 + (double)fy0 * fy0);
    }
}

ORC_API void orc_mlapm_step(const float *pos, const float *vel, const float *ds, int ds_dim, const float *dest,
                            int64_t N, int version, float tau, float A, float Bc, float Cc, float Dc,
                            float theta_deg, float dt, int64_t row0, int64_t row1, float *out) {
    mlapm_step_impl(pos, vel, ds, ds_dim, dest, N, version, tau, A, Bc, Cc, Dc, theta_deg, dt, row0, row1, out, NULL,
                    NULL);
}

/* Diagnostic form for the force-level parity gate: additionally returns, per row, the force BEFORE the Euler step
 * (fp32 pair terms summed in fp64, destination term added in fp64: (rows,2) doubles) and the operand magnitude
 * S = |dest term| + sum_m |term_m| -- the scale fp32 rounding of the row sum is relative to (S / |F| is the
 * condition number of the sum). */
ORC_API void orc_mlapm_step_diag(const float *pos, const float *vel, const float *ds, int ds_dim, const float *dest,
                                 int64_t N, int version, float tau, float A, float Bc, float Cc, float Dc,
                                 float theta_deg, float dt, int64_t row0, int64_t row1, float *out, double *force,
                                 double *opsum) {
    mlapm_step_impl(pos, vel, ds, ds_dim, dest, N, version, tau, A, Bc, Cc, Dc, theta_deg, dt, row0, row1, out, force,
                    opsum);
}

/* ------------------------------------------------------------------------------------------------------------
 * utils.py:31-100 calc_acceleration.  rel (S, stride>=4) -> out (S,2).  version 0='v0', 1='v1', 2='v2'.
 * v1/v2 use the POSITIONS as dv (reference quirk, utils.py:67,84 -> cos == r^2/(r+eps)^2). */
ORC_API void orc_calc_acceleration(const float *rel, int64_t S, int stride, int version, float A, float Bc,
                                   float Cc, float Dc, float theta, float eps, float *out) {
    float ct = (float)cos((double)theta), st = (float)sin((double)theta);
    for (int64_t s = 0; s < S; ++s) {
        float dx = rel[s * stride], dy = rel[s * stride + 1];
        float r = norm2f(dx, dy) + eps;
        if (version == 0) {
            float a = A * expf(Bc * r);
            out[2 * s] = -a * (dx / r);
            out[2 * s + 1] = -a * (dy / r);
        } else {
            float v = r;                              /* dv == dr */
            float c = (dx * dx + dy * dy) / r / v;
            float a = (version == 1) ? A * expf(Bc * r + Cc * c) : A * expf(Bc * r + Cc * c + (Dc * r) * c);
            float nx = dx / r, ny = dy / r;
            if (version == 1) {
                out[2 * s] = -a * nx; out[2 * s + 1] = -a * ny;
            } else {
                out[2 * s] = -a * (ct * nx + (-st) * ny);
                out[2 * s + 1] = -a * (st * nx + ct * ny);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * Interaction networks, model.py:40-119 (MLP / ResBlock / ResDNN) and the four PINNSF forwards
 * (:762-792 pinnsf, :1104-1135 pinnsf_bottleneck, :1185-1221 pinnsf_bm, :1271-1305 pinnsf_m).
 * Accumulation is done in double and rounded to fp32 per layer: the reference's fp32 MKL GEMM and the CUDA
 * kernel's fp32 FMA chain both sit within ~1e-6 of this.
 *
 * Packed parameter layout (see piml_b200/models.py: pack_params):  for each Linear in forward order,
 *   W (out,in) row-major then b (out).   Branch = encoder layers, [processor block-0 linear if proc_mode==1],
 *   decoder layers, predictor; then (for the ped branch) the collision head. */
typedef struct {
    int n_enc; int enc_dims[9];      /* enc_dims[0]=input dim, then n_enc widths */
    int proc_mode;                   /* 0: ResDNN == 2x (processor_hidden_layers > 1) ; 1: relu(Wx+b)+x */
    int n_dec; int dec_dims[9];      /* dec_dims[0]=processor width, then n_dec widths */
    int n_coll; int coll_dims[5];    /* collision head widths incl. input; n_coll==0: none */
    int kind;                        /* 0: bottleneck (per-slot decode, sum of 2-d msgs) ; 1: sum embeddings then decode */
} orc_net_t;

static const float *linear_d(const float *w, const float *x, int in, int out, int relu, float *y) {
    const float *W = w, *b = w + (size_t)in * out;
    for (int o = 0; o < out; ++o) {
        double s = (double)b[o];
        for (int i = 0; i < in; ++i) s += (double)W[(size_t)o * in + i] * (double)x[i];
        float r = (float)s;
        y[o] = (relu && r < 0.f) ? 0.f : r;
    }
    return w + (size_t)in * out + out;
}

/* One branch on one slot-row up to and including the processor.  Returns pointer past consumed params. */
static const float *branch_embed(const orc_net_t *net, const float *w, const float *x, float *emb, float *tmp) {
    float bufa[512], bufb[512];
    const float *cur = x;
    float *nxt = bufa;
    int in = net->enc_dims[0];
    for (int l = 0; l < net->n_enc; ++l) {
        int out = net->enc_dims[l + 1];
        w = linear_d(w, cur, in, out, l < net->n_enc - 1, nxt);     /* MLP: last act = Identity (model.py:57) */
        cur = nxt; nxt = (nxt == bufa) ? bufb : bufa; in = out;
    }
    if (net->proc_mode == 0) {
        for (int i = 0; i < in; ++i) emb[i] = cur[i] + cur[i];        /* ResDNN == x + x (model.py:115-119) */
    } else {
        w = linear_d(w, cur, in, in, 1, tmp);                          /* ResBlock: relu(Wx+b) + x */
        for (int i = 0; i < in; ++i) emb[i] = tmp[i] + cur[i];
    }
    return w;
}

static const float *mlp_chain(const float *w, const float *x, const int *dims, int n, float *y) {
    float bufa[512], bufb[512];
    const float *cur = x;
    float *nxt = bufa;
    int in = dims[0];
    for (int l = 0; l < n; ++l) {
        int out = dims[l + 1];
        w = linear_d(w, cur, in, out, l < n - 1, nxt);
        cur = nxt; nxt = (nxt == bufa) ? bufb : bufa; in = out;
    }
    memcpy(y, cur, sizeof(float) * in);
    return w;
}

static size_t branch_param_count(const orc_net_t *net) {
    size_t c = 0;
    for (int l = 0; l < net->n_enc; ++l) c += (size_t)net->enc_dims[l] * net->enc_dims[l + 1] + net->enc_dims[l + 1];
    int pw = net->enc_dims[net->n_enc];
    if (net->proc_mode == 1) c += (size_t)pw * pw + pw;
    for (int l = 0; l < net->n_dec; ++l) c += (size_t)net->dec_dims[l] * net->dec_dims[l + 1] + net->dec_dims[l + 1];
    c += (size_t)net->dec_dims[net->n_dec] * 2 + 2;
    return c;
}

/* Forward of one PINNSF-family model in eval mode.
 * ped (R,kp,6) obs (R,ko,6) self (R,7); R = all leading dims flattened.  norm_group: number of consecutive rows
 * that share ONE destination norm column-wise -- 0 for the normal (N,7) call (norm over dim=1 == per row);
 * for channelled (C,N,7) inputs the reference reduces over the AGENT axis (dim=1 quirk, SURVEY B-3) and
 * norm_group = N.
 * Outputs: acc (R,2); ped_msgs (R,kp,msgw) ; obs_msgs (R,ko,msgw) ; coll (R,kp) (if head present)
 *   msgw = 2 for kind 0, processor width for kind 1. */
ORC_API void orc_pinnsf_forward(const orc_net_t *net, const float *params, int has_obs, float tau,
                                const float *ped, const float *obs, const float *self, int64_t R, int kp, int ko,
                                int norm_group, float *acc, float *ped_msgs, float *obs_msgs, float *coll) {
    int pw = net->enc_dims[net->n_enc];
    int dw = net->dec_dims[net->n_dec];
    int msgw = net->kind == 0 ? 2 : pw;
    size_t bp = branch_param_count(net);
    const float *wp = params, *wo = params + bp, *wc = params + 2 * bp;
    int cin = net->n_coll ? net->coll_dims[0] : 0;

    /* destination norm (model.py:1206 torch.norm(self[..., :2], dim=1, keepdim=True)) */
    float *dn = (float *)malloc(sizeof(float) * R * 2);
    if (norm_group <= 0) {
        for (int64_t r = 0; r < R; ++r) dn[2 * r] = dn[2 * r + 1] = norm2f(self[r * 7], self[r * 7 + 1]);
    } else {
        for (int64_t g = 0; g < R / norm_group; ++g)
            for (int q = 0; q < 2; ++q) {
                double s = 0.0;
                for (int i = 0; i < norm_group; ++i) {
                    double x = self[(g * norm_group + i) * 7 + q];
                    s += x * x;
                }
                float nv = (float)sqrt(s);
                for (int i = 0; i < norm_group; ++i) dn[(g * norm_group + i) * 2 + q] = nv;
            }
    }

#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < R; ++r) {
        float emb[512], tmp[512], dec[512], sum[512], msg[2];
        double ax = 0.0, ay = 0.0;
        for (int br = 0; br < (has_obs ? 2 : 1); ++br) {
            const float *w0 = br == 0 ? wp : wo;
            const float *feat = br == 0 ? ped + r * kp * 6 : obs + r * ko * 6;
            float *msgs = br == 0 ? ped_msgs + r * kp * msgw : obs_msgs + r * ko * msgw;
            int k = br == 0 ? kp : ko;
            for (int i = 0; i < pw; ++i) sum[i] = 0.f;
            const float *wdec = NULL;
            for (int j = 0; j < k; ++j) {
                wdec = branch_embed(net, w0, feat + j * 6, emb, tmp);
                if (net->kind == 0) {
                    const float *wpred = mlp_chain(wdec, emb, net->dec_dims, net->n_dec, dec);
                    int pd[2] = {dw, 2};
                    mlp_chain(wpred, dec, pd, 1, msg);
                    msgs[j * 2] = msg[0]; msgs[j * 2 + 1] = msg[1];
                    ax += msg[0]; ay += msg[1];
                    if (br == 0 && net->n_coll) {
                        float c1[1];
                        mlp_chain(wc, dec, net->coll_dims, net->n_coll, c1);
                        coll[r * kp + j] = 1.f / (1.f + expf(-c1[0]));
                    }
                } else {
                    for (int i = 0; i < pw; ++i) { msgs[j * pw + i] = emb[i]; sum[i] += emb[i]; }
                    if (br == 0 && net->n_coll) {
                        float c1[1];
                        mlp_chain(wc, emb, net->coll_dims, net->n_coll, c1);
                        coll[r * kp + j] = 1.f / (1.f + expf(-c1[0]));
                    }
                }
            }
            if (net->kind == 1 && k > 0) {
                const float *wpred = mlp_chain(wdec, sum, net->dec_dims, net->n_dec, dec);
                int pd[2] = {dw, 2};
                mlp_chain(wpred, dec, pd, 1, msg);
                ax += msg[0]; ay += msg[1];
            }
        }
        (void)cin;
        /* destination term, model.py:1205-1210 */
        const float *s = self + r * 7;
        float nx = dn[2 * r], ny = dn[2 * r + 1];
        if (nx == 0.f) nx = nx + 0.1f;
        if (ny == 0.f) ny = ny + 0.1f;
        float dxs = (s[6] * (s[0] / nx) - s[2]) / tau;
        float dys = (s[6] * (s[1] / ny) - s[3]) / tau;
        acc[2 * r] = (float)ax + dxs;
        acc[2 * r + 1] = (float)ay + dys;
    }
    free(dn);
}

/* ------------------------------------------------------------------------------------------------------------
 * Pure social-force "model" with the model(ped, obs, self) -> [acc, ped_msgs, obs_msgs] interface (BASELINE config 2,
 * SURVEY.md 8c): per slot the v0 repulsion of utils.py:53-58 (ped constants A_p,B_p; the same form with the obstacle
 * constants A_o,B_o of socialforce.yaml:52-56), summed over slots in slot order (torch.sum over dim -2), plus the
 * destination term model.py:1205-1210 for (N,7) inputs.  Padded slots (dr = 0) give exactly 0.
 * ped (R,kp,6), obs (R,ko,6) or NULL, self (R,7) -> acc (R,2), ped_msgs (R,kp,2), obs_msgs (R,ko,2) (may be NULL). */
ORC_API void orc_sfm_forward(const float *ped, const float *obs, const float *self, int64_t R, int kp, int ko,
                             float A_p, float B_p, float A_o, float B_o, float eps, float tau, float *acc,
                             float *ped_msgs, float *obs_msgs) {
    for (int64_t r = 0; r < R; ++r) {
        float ax = 0.f, ay = 0.f, ox = 0.f, oy = 0.f;
        for (int j = 0; j < kp; ++j) {
            const float *f = ped + (r * kp + j) * 6;
            float rr = norm2f(f[0], f[1]) + eps;
            float a = A_p * expf(B_p * rr);
            float mx = -a * (f[0] / rr), my = -a * (f[1] / rr);
            if (ped_msgs) { ped_msgs[(r * kp + j) * 2] = mx; ped_msgs[(r * kp + j) * 2 + 1] = my; }
            ax += mx; ay += my;
        }
        for (int j = 0; obs && j < ko; ++j) {
            const float *f = obs + (r * ko + j) * 6;
            float rr = norm2f(f[0], f[1]) + eps;
            float a = A_o * expf(B_o * rr);
            float mx = -a * (f[0] / rr), my = -a * (f[1] / rr);
            if (obs_msgs) { obs_msgs[(r * ko + j) * 2] = mx; obs_msgs[(r * ko + j) * 2 + 1] = my; }
            ox += mx; oy += my;
        }
        const float *s = self + r * 7;
        float n = norm2f(s[0], s[1]);
        if (n == 0.f) n = n + 0.1f;
        float dxs = (s[6] * (s[0] / n) - s[2]) / tau;
        float dys = (s[6] * (s[1] / n) - s[3]) / tau;
        acc[2 * r] = (ax + ox) + dxs;
        acc[2 * r + 1] = (ay + oy) + dys;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * One inference-rollout state update, simulators.py:603-639, everything except the model
 * forward and the feature rebuild.  All arrays for ONE scene of N slots.
 *   p,v,a (N,2) in/out ; a_next (N,2) model output ; dest (N,2) in/out ; dest_idx (N) in/out ; dest_num (N)
 *   waypoints (D,N,2) ; hist_v (N,2) out (num_history_velocity == 1)
 *   entry (N) int: new_peds_flag[t+1] or NULL when t == T-1 ; *_next: ground truth at t+1 (N,..)
 *   remove_on_arrival: 1 for get_multiple_rollouts (:611), 0 for the training rollout (:748-751). */
ORC_API void orc_integrate_step(float *p, float *v, float *a, const float *a_next, float *dest, int64_t *dest_idx,
                                const int64_t *dest_num, const float *waypoints, int N, float dt,
                                int remove_on_arrival, const int64_t *entry, const float *p_gt, const float *v_gt,
                                const float *a_gt, const float *dest_gt, const int64_t *dest_idx_gt, float *hist_v) {
    for (int n = 0; n < N; ++n) {
        float px = p[2 * n], py = p[2 * n + 1];
        float vnx = v[2 * n] + a[2 * n] * dt, vny = v[2 * n + 1] + a[2 * n + 1] * dt;    /* :603 */
        float pnx = px + v[2 * n] * dt, pny = py + v[2 * n + 1] * dt;                        /* :604 */
        float dis = norm2f(px - dest[2 * n], py - dest[2 * n + 1]);                          /* :608 */
        if (dis < 0.5f) dest_idx[n] += 1;                                                    /* :609 */
        if (dest_idx[n] > dest_num[n] - 1) {
            if (remove_on_arrival) { pnx = NAN; pny = NAN; }                                 /* :611 */
            dest_idx[n] -= 1;                                                                /* :613 */
        }
        dest[2 * n] = waypoints[((size_t)dest_idx[n] * N + n) * 2];                          /* :614-616 */
        dest[2 * n + 1] = waypoints[((size_t)dest_idx[n] * N + n) * 2 + 1];
        p[2 * n] = pnx; p[2 * n + 1] = pny;
        v[2 * n] = vnx; v[2 * n + 1] = vny;
        a[2 * n] = a_next[2 * n]; a[2 * n + 1] = a_next[2 * n + 1];
        if (hist_v) { hist_v[2 * n] = vnx; hist_v[2 * n + 1] = vny; }                        /* :624-626 */
        if (entry && entry[n] == 1) {                                                        /* :629-639 */
            p[2 * n] = p_gt[2 * n]; p[2 * n + 1] = p_gt[2 * n + 1];
            v[2 * n] = v_gt[2 * n]; v[2 * n + 1] = v_gt[2 * n + 1];
            a[2 * n] = a_gt[2 * n]; a[2 * n + 1] = a_gt[2 * n + 1];
            dest[2 * n] = dest_gt[2 * n]; dest[2 * n + 1] = dest_gt[2 * n + 1];
            dest_idx[n] = dest_idx_gt[n];
            if (hist_v) { hist_v[2 * n] = v_gt[2 * n]; hist_v[2 * n + 1] = v_gt[2 * n + 1]; }
        }
    }
}
