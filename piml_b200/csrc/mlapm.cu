// mlapm.cu -- dense all-pairs MLAPM force for sm_100a.
//
// Replaces MLAPM.step (reference src/models/mlapm.py:10-58) and the loop body of src/main_mlapm.py:19-34.
// The reference materialises (N,N,2)x4 + (N,N,2,2) + (N,N)x6 temporaries (~100 B per ordered pair); here each thread
// keeps R rows in registers and streams the columns through shared memory (TMA bulk copies of position / velocity
// tiles, double buffered), so DRAM traffic is O(N) and the kernel is bound by the FP32 and MUFU pipes.
//
// Grid: (row blocks of 128*R rows) x (column splits).  Each CTA writes its partial row sums to workspace[split][row];
// a finalize kernel adds the splits in a fixed order (deterministic), applies the destination term and the Euler
// update.  The field-of-view gate `v_n . (p_m - p_n) > 0` uses the reference's exact fp32 form
// fmaf(v1, r1, v0*r0) (einsum -> bmm); the force magnitude only has to hold 1e-5 relative, which
// both math variants below do:
//   EXACT: IEEE div / sqrt / expf in the reference's operation order (validation variant)
//   fast : rsqrt.approx + ex2.approx, shared 1/r, constant-angle rotation folded into two FMAs (production)
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace piml {

constexpr int ML_THREADS = 128;
constexpr int ML_TILE = 512;          // columns per shared-memory stage: 4 KB positions + 4 KB velocities

struct MlConst {
    int version;
    float A, B, C, D;                  // reference constants
    float Bl, Cl, Dl;                  // pre-multiplied by log2(e) for ex2
    float cos_t, sin_t;                // cos/sin of theta (fp32 angle as the reference builds it)
    float inv_tau, tau;
};

__device__ __forceinline__ float ex2_approx(float x) {            // one MUFU.EX2
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {          // one MUFU.RSQ
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Stage one tile of positions and velocities (n float2 each) on ONE mbarrier phase.  All threads call it.
__device__ __forceinline__ bool stage_tile_pv(float2 *dp, float2 *dv, const float2 *gp, const float2 *gv, int n,
                                              uint64_t *bar) {
    const bool tma = ((reinterpret_cast<uintptr_t>(gp) | reinterpret_cast<uintptr_t>(gv)) & 15u) == 0 && n >= 2;
    if (tma) {
        if (threadIdx.x == 0) {
            const uint32_t bytes = static_cast<uint32_t>(n & ~1) * 8u;
            mbar_expect_tx(bar, 2u * bytes);
            tma_bulk_g2s(dp, gp, bytes, bar);
            tma_bulk_g2s(dv, gv, bytes, bar);
            if (n & 1) { dp[n - 1] = gp[n - 1]; dv[n - 1] = gv[n - 1]; }
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) { dp[i] = gp[i]; dv[i] = gv[i]; }
    }
    return tma;
}

// One ordered pair, production math.  ~37 issue slots for version GC.
template <int VERSION>
__device__ __forceinline__ void pair_fast(float px, float py, float vx, float vy, float ex, float ey, float2 pm,
                                          float2 vm, const MlConst &k, float &fx, float &fy) {
    const float rx = __fsub_rn(pm.x, px);
    const float ry = __fsub_rn(pm.y, py);
    const float gate = (__fmaf_rn(vy, ry, __fmul_rn(vx, rx)) > 0.f) ? k.A : 0.f;     // exact FoV gate (mlapm.py:27)
    const float r2 = fmaf(ry, ry, rx * rx);
    const float inv_r = rsqrt_approx(fmaxf(r2, 1e-30f));
    const float r = r2 * inv_r;
    if (VERSION == 0) {
        const float w = gate * ex2_approx(k.Bl * r) * inv_r;          // view*A*exp(B r) * vr/|vr|   (mlapm.py:29)
        fx = fmaf(w, rx, fx);
        fy = fmaf(w, ry, fy);
    } else {
        const float ux = vm.x - vx, uy = vm.y - vy;
        const float u2 = fmaf(uy, uy, ux * ux);
        const float dot = fmaf(ry, uy, rx * ux);
        const float cosv = dot * inv_r * rsqrt_approx(fmaxf(u2, 1e-30f));   // cosine_similarity(vr, vv)  (mlapm.py:32)
        const float cross = fmaf(rx, ey, -(ry * ex));                 // sign picks +-theta           (mlapm.py:33-34)
        const float s = (cross > 0.f) ? -k.sin_t : k.sin_t;
        const float dx = fmaf(k.cos_t, rx, -(s * ry));                // R(theta) * vr                (mlapm.py:35-38)
        const float dy = fmaf(s, rx, k.cos_t * ry);
        const float arg = fmaf(fmaf(k.Dl, r, k.Cl), cosv, k.Bl * r);  // (B r + C cos + D r cos) * log2 e
        const float w = gate * ex2_approx(arg) * inv_r;
        fx = fmaf(w, dx, fx);
        fy = fmaf(w, dy, fy);
    }
}

// One ordered pair in the reference's own operation order with IEEE arithmetic (validation variant).
template <int VERSION>
__device__ __forceinline__ void pair_exact(float px, float py, float vx, float vy, float ex, float ey, float2 pm,
                                           float2 vm, const MlConst &k, float &fx, float &fy) {
    const float rx = __fsub_rn(pm.x, px);
    const float ry = __fsub_rn(pm.y, py);
    const float r = norm2_rn(rx, ry);
    const float gate = (__fmaf_rn(vy, ry, __fmul_rn(vx, rx)) > 0.f) ? 1.f : 0.f;
    const float nr = fmaxf(r, 1e-12f);
    const float nx = __fdiv_rn(rx, nr), ny = __fdiv_rn(ry, nr);
    const float ga = __fmul_rn(gate, k.A);
    if (VERSION == 0) {
        const float e = expf(__fmul_rn(k.B, r));
        fx = __fadd_rn(fx, __fmul_rn(__fmul_rn(ga, e), nx));
        fy = __fadd_rn(fy, __fmul_rn(__fmul_rn(ga, e), ny));
    } else {
        const float ux = __fsub_rn(vm.x, vx), uy = __fsub_rn(vm.y, vy);
        const float cr = fmaxf(r, 1e-8f), cu = fmaxf(norm2_rn(ux, uy), 1e-8f);
        const float cosv = __fadd_rn(__fmul_rn(__fdiv_rn(rx, cr), __fdiv_rn(ux, cu)),
                                     __fmul_rn(__fdiv_rn(ry, cr), __fdiv_rn(uy, cu)));
        const float cross = __fsub_rn(__fmul_rn(rx, ey), __fmul_rn(ry, ex));
        const float s = (cross > 0.f) ? -k.sin_t : k.sin_t;
        const float dx = __fadd_rn(__fmul_rn(k.cos_t, nx), __fmul_rn(-s, ny));
        const float dy = __fadd_rn(__fmul_rn(s, nx), __fmul_rn(k.cos_t, ny));
        const float arg = __fadd_rn(__fadd_rn(__fmul_rn(k.B, r), __fmul_rn(k.C, cosv)),
                                    __fmul_rn(__fmul_rn(k.D, r), cosv));
        const float e = expf(arg);
        fx = __fadd_rn(fx, __fmul_rn(__fmul_rn(ga, e), dx));
        fy = __fadd_rn(fy, __fmul_rn(__fmul_rn(ga, e), dy));
    }
}

// partial[split][row - row0] = sum over this split's columns of  view*A*exp(..)*direc
template <int VERSION, int R, bool EXACT>
__global__ void __launch_bounds__(ML_THREADS) mlapm_pairs_kernel(const float2 *__restrict__ pos,
                                                                 const float2 *__restrict__ vel,
                                                                 const float2 *__restrict__ dest, int N, int row0,
                                                                 int row1, int cols_per_split, MlConst k,
                                                                 float2 *__restrict__ partial) {
    __shared__ __align__(16) float2 sp[2][ML_TILE];
    __shared__ __align__(16) float2 sv[2][ML_TILE];
    __shared__ __align__(8) uint64_t bars[2];
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int nrows = row1 - row0;
    const int rbase = blockIdx.x * (ML_THREADS * R);
    float px[R], py[R], vx[R], vy[R], ex[R], ey[R], fx[R], fy[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        int rl = rbase + i * ML_THREADS + threadIdx.x;
        rl = rl < nrows ? rl : nrows - 1;
        const int n = row0 + rl;
        const float2 p = pos[n], v = vel[n], d = dest[n];
        px[i] = p.x; py[i] = p.y; vx[i] = v.x; vy[i] = v.y;
        // ed = F.normalize(destination - position)   (mlapm.py:21)
        const float dx = __fsub_rn(d.x, p.x), dy = __fsub_rn(d.y, p.y);
        const float dn = fmaxf(norm2_rn(dx, dy), 1e-12f);
        ex[i] = __fdiv_rn(dx, dn); ey[i] = __fdiv_rn(dy, dn);
        fx[i] = 0.f; fy[i] = 0.f;
    }

    const int c0 = blockIdx.y * cols_per_split;
    const int c1 = min(N, c0 + cols_per_split);
    const int ncols = c1 - c0;
    const int ntiles = (ncols + ML_TILE - 1) / ML_TILE;
    uint32_t phase_bits = 0;
    bool waits[2] = {false, false};
    if (ntiles > 0) {
        const int tn = min(ncols, ML_TILE);
        waits[0] = stage_tile_pv(sp[0], sv[0], pos + c0, vel + c0, tn, &bars[0]);
    }
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) {
            const int m1 = c0 + (t + 1) * ML_TILE;
            const int tn1 = min(c1 - m1, ML_TILE);
            waits[buf ^ 1] = stage_tile_pv(sp[buf ^ 1], sv[buf ^ 1], pos + m1, vel + m1, tn1, &bars[buf ^ 1]);
        }
        if (waits[buf]) {
            mbar_wait(&bars[buf], (phase_bits >> buf) & 1u);
            phase_bits ^= (1u << buf);
        }
        const int tn = min(c1 - (c0 + t * ML_TILE), ML_TILE);
        const float2 *tp = sp[buf];
        const float2 *tv = sv[buf];
#pragma unroll 2
        for (int j = 0; j < tn; ++j) {
            const float2 pm = tp[j];
            const float2 vm = tv[j];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                if (EXACT) pair_exact<VERSION>(px[i], py[i], vx[i], vy[i], ex[i], ey[i], pm, vm, k, fx[i], fy[i]);
                else pair_fast<VERSION>(px[i], py[i], vx[i], vy[i], ex[i], ey[i], pm, vm, k, fx[i], fy[i]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int rl = rbase + i * ML_THREADS + threadIdx.x;
        if (rl < nrows) partial[static_cast<int64_t>(blockIdx.y) * nrows + rl] = make_float2(fx[i], fy[i]);
    }
}

// force = (v0*ed - v)/tau - sum_splits partial ; action = v + force*dt ; optional p' = p + action*dt and arrival.
__global__ void mlapm_finalize_kernel(const float2 *__restrict__ pos, const float2 *__restrict__ vel,
                                      const float *__restrict__ ds, int ds_dim, const float2 *__restrict__ dest,
                                      int row0, int row1, int nsplit, const float2 *__restrict__ partial, float tau,
                                      float dt, float radius, float2 *__restrict__ action,
                                      float2 *__restrict__ pos_new, uint8_t *__restrict__ arrived) {
    const int rl = blockIdx.x * blockDim.x + threadIdx.x;
    const int nrows = row1 - row0;
    if (rl >= nrows) return;
    const int n = row0 + rl;
    const float2 p = pos[n], v = vel[n], d = dest[n];
    const float dx = __fsub_rn(d.x, p.x), dy = __fsub_rn(d.y, p.y);
    const float dn = fmaxf(norm2_rn(dx, dy), 1e-12f);
    const float ex = __fdiv_rn(dx, dn), ey = __fdiv_rn(dy, dn);
    const float dsx = ds[static_cast<int64_t>(n) * ds_dim];
    const float dsy = ds[static_cast<int64_t>(n) * ds_dim + (ds_dim > 1 ? 1 : 0)];
    // force += (desired_speed * ed - velocity) / tau      (mlapm.py:22)
    float fx = __fdiv_rn(__fsub_rn(__fmul_rn(dsx, ex), v.x), tau);
    float fy = __fdiv_rn(__fsub_rn(__fmul_rn(dsy, ey), v.y), tau);
    float sx = 0.f, sy = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        const float2 q = partial[static_cast<int64_t>(s) * nrows + rl];
        sx = __fadd_rn(sx, q.x); sy = __fadd_rn(sy, q.y);
    }
    fx = __fsub_rn(fx, sx); fy = __fsub_rn(fy, sy);               // force -= (...).sum(dim=1)   (mlapm.py:29/39)
    const float ax = __fadd_rn(v.x, __fmul_rn(fx, dt));           // action = velocity + force*dt (mlapm.py:57)
    const float ay = __fadd_rn(v.y, __fmul_rn(fy, dt));
    action[rl] = make_float2(ax, ay);
    if (pos_new || arrived) {
        const float qx = __fadd_rn(p.x, __fmul_rn(ax, dt));       // p = position + v*dt          (main_mlapm.py:26)
        const float qy = __fadd_rn(p.y, __fmul_rn(ay, dt));
        if (pos_new) pos_new[rl] = make_float2(qx, qy);
        if (arrived)                                              // ||p - destination|| < radius (main_mlapm.py:34)
            arrived[rl] = norm2_rn(__fsub_rn(qx, d.x), __fsub_rn(qy, d.y)) < radius ? 1 : 0;
    }
}


// =====================================================================================================================
// Production pair kernel (v2): packed FP32 (Blackwell FFMA2/FMUL2/FADD2, two rows per instruction), SoA-duplicated
// column tiles staged by ONE TMA bulk copy per stage, rotation and amplitude hoisted out of the pair loop.
//
//   sum_m view*A*exp(..)*R(theta_nm) r^  =  A * [ cos_t*Sx - sin_t*Ty' ,  sin_t*Tx' + cos_t*Sy ]
//     Sx = sum w rx,  Sy = sum w ry,  Ty' = sum (w sigma) ry,  Tx' = sum (w sigma) rx,
//     w = view * exp(..)/r,  sigma = -1 if (vr x e) > 0 else +1                        (mlapm.py:33-39)
//
// Per ordered pair: 26 packed-FP32 instructions per TWO pairs + 3 MUFU + 2 ALU-pipe selects per pair, i.e. 19 issue
// slots per pair (v1: 37), FP32 pipe 26 and MUFU pipe 24 cycles per 32 pairs per SM sub-partition.
// The two discontinuous gates use the reference's exact fp32 arithmetic: view = fmaf(vy,ry,vx*rx) > 0 (einsum/bmm),
// cross = fl(fl(rx*ey) - fl(ry*ex)) un-fused, whose sign is the order of the two rounded products (equal products ->
// sigma = +1, the reference's masked_fill_(theta == 0, +theta)).
// =====================================================================================================================
constexpr int M2_THREADS = 128;
constexpr int M2_TILE = 512;                      // columns per stage: 512 * 16 B = 8 KB
constexpr int M2_COLF = 4;                        // floats per column record {px,py,vx,vy}; the packed instructions
                                                  // read them as scalar-broadcast operands (SASS `Rn.F32`)

__device__ __forceinline__ float2 splat(float x) { return make_float2(x, x); }

// col8[j] = {px,py,vx,vy} for j < N, padded up to Npad with copies of the last agent (never evaluated:
// the pair loop stops at N; the padding only keeps the fixed-size TMA bulk copies in bounds).
__global__ void mlapm_prep_kernel(const float2 *__restrict__ pos, const float2 *__restrict__ vel, int N, int Npad,
                                  float4 *__restrict__ col8) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Npad) return;
    const int s = j < N ? j : N - 1;
    const float2 p = pos[s], v = vel[s];
    col8[j] = make_float4(p.x, p.y, v.x, v.y);
}

struct M2Const {
    float Bl, Cl, Dl;                  // B, C, D times log2(e)
};

// Two rows (packed lanes) against one column.
template <int VERSION>
__device__ __forceinline__ void pair2(const float2 npx, const float2 npy, const float2 vx, const float2 vy,
                                      const float2 nvx, const float2 nvy, const float2 ex, const float2 ey,
                                      const float4 cp, const float2 Bl, const float2 Cl,
                                      const float2 Dl, float2 &Sx, float2 &Sy, float2 &Tx, float2 &Ty) {
    const float2 eps = splat(1e-30f);
    const float2 rx = __fadd2_rn(make_float2(cp.x, cp.x), npx);            // vr = p_m - p_n          (mlapm.py:25)
    const float2 ry = __fadd2_rn(make_float2(cp.y, cp.y), npy);
    const float2 g = __ffma2_rn(vy, ry, __fmul2_rn(vx, rx));               // einsum('nk,nmk->nm')    (mlapm.py:27)
    const float2 r2 = __ffma2_rn(ry, ry, __ffma2_rn(rx, rx, eps));
    const float2 ir = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
    const float2 r = __fmul2_rn(r2, ir);
    float2 w;
    if (VERSION == 0) {
        const float2 arg = __fmul2_rn(Bl, r);
        w = __fmul2_rn(make_float2(ex2_approx(arg.x), ex2_approx(arg.y)), ir);
        w.x = g.x > 0.f ? w.x : 0.f;
        w.y = g.y > 0.f ? w.y : 0.f;
        Sx = __ffma2_rn(w, rx, Sx);
        Sy = __ffma2_rn(w, ry, Sy);
    } else {
        const float2 ux = __fadd2_rn(make_float2(cp.z, cp.z), nvx);        // vv = v_m - v_n          (mlapm.py:31)
        const float2 uy = __fadd2_rn(make_float2(cp.w, cp.w), nvy);
        const float2 u2 = __ffma2_rn(uy, uy, __ffma2_rn(ux, ux, eps));
        const float2 iu = make_float2(rsqrt_approx(u2.x), rsqrt_approx(u2.y));
        const float2 dot = __ffma2_rn(ry, uy, __fmul2_rn(rx, ux));
        // cosine_similarity cs = dot*ir*iu (mlapm.py:32); B r + (C + D r) cs = B r + (dot*iu) (C/r + D): one
        // packed instruction fewer than forming cs (r*ir == 1 to 2^-22, far inside the 1e-5 gate)
        const float2 q = __fmul2_rn(dot, iu);
        // vr x e = fl(rx*ey) - fl(ry*ex) un-fused: its sign is the order of the two rounded products    (mlapm.py:33)
        const float2 m1 = __fmul2_rn(rx, ey), m2 = __fmul2_rn(ry, ex);
        const float2 arg = __ffma2_rn(q, __ffma2_rn(Cl, ir, Dl), __fmul2_rn(Bl, r));
        w = __fmul2_rn(make_float2(ex2_approx(arg.x), ex2_approx(arg.y)), ir);
        w.x = g.x > 0.f ? w.x : 0.f;                                       // view gate
        w.y = g.y > 0.f ? w.y : 0.f;
        float2 ws;                                                         // w * sigma; cross == 0 -> +theta  (:34)
        ws.x = m1.x > m2.x ? -w.x : w.x;
        ws.y = m1.y > m2.y ? -w.y : w.y;
        Sx = __ffma2_rn(w, rx, Sx);
        Sy = __ffma2_rn(w, ry, Sy);
        Tx = __ffma2_rn(ws, rx, Tx);
        Ty = __ffma2_rn(ws, ry, Ty);
    }
}

// partial[split][row - row0] = (Sx, Sy, Tx', Ty') over this split's columns.  Each thread owns 2*RP rows.
template <int VERSION, int RP, int UNROLL>
__global__ void __launch_bounds__(M2_THREADS) mlapm_pairs2_kernel(const float2 *__restrict__ pos,
                                                                  const float2 *__restrict__ vel,
                                                                  const float2 *__restrict__ dest,
                                                                  const float4 *__restrict__ col8, int N, int row0,
                                                                  int row1, int cols_per_split, M2Const k,
                                                                  float4 *__restrict__ partial) {
    __shared__ __align__(128) float4 tile[2][M2_TILE];
    __shared__ __align__(8) uint64_t bars[2];
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int nrows = row1 - row0;
    const int rbase = blockIdx.x * (M2_THREADS * 2 * RP);
    float2 npx[RP], npy[RP], vx[RP], vy[RP], nvx[RP], nvy[RP], ex[RP], ey[RP], Sx[RP], Sy[RP], Tx[RP], Ty[RP];
#pragma unroll
    for (int i = 0; i < RP; ++i) {
        float t[2][6];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int rl = rbase + (2 * i + h) * M2_THREADS + threadIdx.x;
            rl = rl < nrows ? rl : nrows - 1;
            const int n = row0 + rl;
            const float2 p = pos[n], v = vel[n], d = dest[n];
            // ed = F.normalize(destination - position)   (mlapm.py:21)
            const float dx = __fsub_rn(d.x, p.x), dy = __fsub_rn(d.y, p.y);
            const float dn = fmaxf(norm2_rn(dx, dy), 1e-12f);
            t[h][0] = p.x; t[h][1] = p.y; t[h][2] = v.x; t[h][3] = v.y;
            t[h][4] = __fdiv_rn(dx, dn); t[h][5] = __fdiv_rn(dy, dn);
        }
        npx[i] = make_float2(-t[0][0], -t[1][0]); npy[i] = make_float2(-t[0][1], -t[1][1]);
        vx[i] = make_float2(t[0][2], t[1][2]); vy[i] = make_float2(t[0][3], t[1][3]);
        nvx[i] = make_float2(-t[0][2], -t[1][2]); nvy[i] = make_float2(-t[0][3], -t[1][3]);
        ex[i] = make_float2(t[0][4], t[1][4]); ey[i] = make_float2(t[0][5], t[1][5]);
        Sx[i] = Sy[i] = Tx[i] = Ty[i] = make_float2(0.f, 0.f);
    }
    const float2 Bl = splat(k.Bl), Cl = splat(k.Cl), Dl = splat(k.Dl);

    const int c0 = blockIdx.y * cols_per_split;                 // multiple of M2_TILE
    const int c1 = min(N, c0 + cols_per_split);
    const int ntiles = (c1 - c0 + M2_TILE - 1) / M2_TILE;
    constexpr uint32_t TILE_BYTES = M2_TILE * M2_COLF * sizeof(float);
    if (threadIdx.x == 0 && ntiles > 0) {
        mbar_expect_tx(&bars[0], TILE_BYTES);
        tma_bulk_g2s(tile[0], col8 + static_cast<int64_t>(c0), TILE_BYTES, &bars[0]);
    }
    uint32_t phase_bits = 0;
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (threadIdx.x == 0 && t + 1 < ntiles) {               // tile[buf^1] was released by the barrier below
            mbar_expect_tx(&bars[buf ^ 1], TILE_BYTES);
            tma_bulk_g2s(tile[buf ^ 1], col8 + static_cast<int64_t>(c0 + (t + 1) * M2_TILE), TILE_BYTES,
                         &bars[buf ^ 1]);
        }
        mbar_wait(&bars[buf], (phase_bits >> buf) & 1u);
        phase_bits ^= (1u << buf);
        const int tn = min(c1 - (c0 + t * M2_TILE), M2_TILE);
        const float4 *tl = tile[buf];
#pragma unroll(UNROLL)
        for (int j = 0; j < tn; ++j) {
            const float4 cp = tl[j];
#pragma unroll
            for (int i = 0; i < RP; ++i)
                pair2<VERSION>(npx[i], npy[i], vx[i], vy[i], nvx[i], nvy[i], ex[i], ey[i], cp, Bl, Cl, Dl, Sx[i],
                               Sy[i], Tx[i], Ty[i]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RP; ++i) {
        const int ra = rbase + (2 * i) * M2_THREADS + threadIdx.x;
        const int rb = ra + M2_THREADS;
        float4 *out = partial + static_cast<int64_t>(blockIdx.y) * nrows;
        if (ra < nrows) out[ra] = make_float4(Sx[i].x, Sy[i].x, Tx[i].x, Ty[i].x);
        if (rb < nrows) out[rb] = make_float4(Sx[i].y, Sy[i].y, Tx[i].y, Ty[i].y);
    }
}

// =====================================================================================================================
// Symmetric pair kernel (v3): every UNORDERED pair {n,m} is evaluated once and feeds both rows.
//
// Everything smooth in the GC formula is symmetric under n <-> m: vr -> -vr, vv -> -vv, so r, |vv|, cos(vr,vv) and
// the exponent B r + C cos + D r cos are the same numbers for (n,m) and (m,n); only the two discontinuous gates (the
// view test of the row's own velocity and the sign of vr x e of the row's own destination direction) and the
// accumulation differ.  The 16 packed instructions + 3 MUFU of the shared part are therefore spent once per unordered
// pair, and each direction adds 8 packed instructions (gate products, cross products, 4 accumulations):
// 32 packed instructions + 6 MUFU per 128 ordered pairs instead of 48 + 12.  The gates keep the reference's exact
// fp32 arithmetic in BOTH directions because fp32 subtraction, multiplication and fma are odd-symmetric:
//   vr[m,n] = fl(p_n - p_m) = -fl(p_m - p_n),  v_m . vr[m,n] = -fmaf(v_my, ry, fl(v_mx rx))  ->  view' = (gc < 0),
//   vr[m,n] x e_m = fl(fl(ry e_mx) - fl(rx e_my))                                          ->  sigma' = -1 iff m2c > m1c.
//
// Schedule: agents are cut into blocks of 512; block pair (I, J = I + d mod T) is evaluated by the CTA of row block I
// for d = 0 .. floor(T/2) (a circulant schedule: every unordered block pair exactly once, every row block the same
// amount of work; for even T the pairs at d = T/2 belong to I < T/2).  d = 0 is the diagonal block, evaluated one
// direction at a time like v2.  A thread keeps 8 rows (4 packed lane pairs) in registers and streams J's columns from
// shared memory (TMA bulk copies, double buffered), so the row direction accumulates in registers as before.  The
// column direction needs, per column, the sum over all 512 rows of the CTA: the two packed lanes are added, each lane
// parks its 4 sums for a batch of 8 columns in a per-warp shared-memory scratch, the warp transposes (lane -> column
// lane/4, every 4th entry) and finishes with two shuffle stages; the 4 warps' column sums are added in a fixed order
// and written to partialC[(I, d)][column] -- deterministic, no atomics.  Cost of the reduction per column and warp:
// 9 FADD + 1 SHFL + 2 x 128-bit shared accesses against 64 packed instructions of pair work.
// The finalize kernel subtracts the column-direction sums (vr' = -vr) from the row-direction sums.
// =====================================================================================================================
constexpr int MS_THREADS = 128;
constexpr int MS_BLOCK = 512;                     // agents per block of the schedule = rows per CTA (4 per thread)
constexpr int MS_CT = 128;                        // default columns per shared-memory stage (4 KB; 35 KB per CTA)
constexpr int MS_BATCH = 8;                       // default columns per warp-level reduction
constexpr int MS_RED_STRIDE = 36;                 // float4 per column of the scratch: 32 lanes + 64 B skew (no conflicts)
constexpr int MS_RECF = 8;                        // floats per agent record {px,py,vx,vy, ex,ey,0,0}
constexpr float MS_FAR = 1.0e18f;                 // padding agents sit here: exp(B r) == 0 exactly, nothing overflows

template <int CT, int BATCH, int THREADS = MS_THREADS>
struct MsSmem {
    float4 tile[2][CT * 2];
    float4 red[THREADS / 32][BATCH * MS_RED_STRIDE];
    float4 colacc[THREADS / 32][CT];
    uint64_t bars[2];
};

// rec[j] = {px,py,vx,vy},{ex,ey,0,0}, e = F.normalize(destination - position) (mlapm.py:21); j >= N: padding agents
// far away (their weight underflows to exactly 0 in both directions), so block tails need no masks.
__global__ void mlapm_prep_sym_kernel(const float2 *__restrict__ pos, const float2 *__restrict__ vel,
                                      const float2 *__restrict__ dest, int N, int Npad, float4 *__restrict__ rec,
                                      ulonglong2 *__restrict__ colsum) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Npad) return;
    colsum[2 * j] = make_ulonglong2(0ull, 0ull);                   // this step's column-direction accumulators
    colsum[2 * j + 1] = make_ulonglong2(0ull, 0ull);
    if (j >= N) {
        rec[2 * j] = make_float4(MS_FAR, MS_FAR, 0.f, 0.f);
        rec[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float2 p = pos[j], v = vel[j], d = dest[j];
    const float dx = __fsub_rn(d.x, p.x), dy = __fsub_rn(d.y, p.y);
    const float dn = fmaxf(norm2_rn(dx, dy), 1e-12f);
    rec[2 * j] = make_float4(p.x, p.y, v.x, v.y);
    rec[2 * j + 1] = make_float4(__fdiv_rn(dx, dn), __fdiv_rn(dy, dn), 0.f, 0.f);
}

// Two rows (packed lanes) against one column, both directions.  Row sums accumulate in S/T; the column's share is
// returned in cS/cT (FIRST: overwritten, else accumulated) in the row's sign convention (the caller subtracts it).
template <int VERSION, bool FIRST>
__device__ __forceinline__ void pair2sym(const float2 npx, const float2 npy, const float2 vx, const float2 vy,
                                         const float2 nvx, const float2 nvy, const float2 ex, const float2 ey,
                                         const float4 cp, const float4 ce, const float2 Bl, const float2 Cl,
                                         const float2 Dl, float2 &Sx, float2 &Sy, float2 &Tx, float2 &Ty,
                                         float2 &cSx, float2 &cSy, float2 &cTx, float2 &cTy) {
    const float2 eps = splat(1e-30f);
    const float2 rx = __fadd2_rn(splat(cp.x), npx);                        // vr[n,m] = p_m - p_n     (mlapm.py:25)
    const float2 ry = __fadd2_rn(splat(cp.y), npy);
    const float2 g = __ffma2_rn(vy, ry, __fmul2_rn(vx, rx));               // v_n . vr[n,m]           (mlapm.py:27)
    const float2 gc = __ffma2_rn(splat(cp.w), ry, __fmul2_rn(splat(cp.z), rx));   // = -(v_m . vr[m,n])
    const float2 r2 = __ffma2_rn(ry, ry, __ffma2_rn(rx, rx, eps));
    const float2 ir = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
    const float2 r = __fmul2_rn(r2, ir);
    float2 w0;
    if (VERSION == 0) {
        const float2 arg = __fmul2_rn(Bl, r);
        w0 = __fmul2_rn(make_float2(ex2_approx(arg.x), ex2_approx(arg.y)), ir);
    } else {
        const float2 ux = __fadd2_rn(splat(cp.z), nvx);                    // vv[n,m] = v_m - v_n     (mlapm.py:31)
        const float2 uy = __fadd2_rn(splat(cp.w), nvy);
        const float2 u2 = __ffma2_rn(uy, uy, __ffma2_rn(ux, ux, eps));
        const float2 iu = make_float2(rsqrt_approx(u2.x), rsqrt_approx(u2.y));
        const float2 dot = __ffma2_rn(ry, uy, __fmul2_rn(rx, ux));
        const float2 q = __fmul2_rn(dot, iu);
        const float2 arg = __ffma2_rn(q, __ffma2_rn(Cl, ir, Dl), __fmul2_rn(Bl, r));
        w0 = __fmul2_rn(make_float2(ex2_approx(arg.x), ex2_approx(arg.y)), ir);
    }
    float2 w, wc;
    w.x = g.x > 0.f ? w0.x : 0.f;                                          // view gate of the row
    w.y = g.y > 0.f ? w0.y : 0.f;
    wc.x = gc.x < 0.f ? w0.x : 0.f;                                        // view gate of the column
    wc.y = gc.y < 0.f ? w0.y : 0.f;
    Sx = __ffma2_rn(w, rx, Sx);
    Sy = __ffma2_rn(w, ry, Sy);
    cSx = FIRST ? __fmul2_rn(wc, rx) : __ffma2_rn(wc, rx, cSx);
    cSy = FIRST ? __fmul2_rn(wc, ry) : __ffma2_rn(wc, ry, cSy);
    if (VERSION != 0) {
        // sign of vr x e from the order of the two rounded products, both directions             (mlapm.py:33-34)
        const float2 m1 = __fmul2_rn(rx, ey), m2 = __fmul2_rn(ry, ex);
        const float2 m1c = __fmul2_rn(rx, splat(ce.y)), m2c = __fmul2_rn(ry, splat(ce.x));
        float2 ws, wcs;
        ws.x = m1.x > m2.x ? -w.x : w.x;
        ws.y = m1.y > m2.y ? -w.y : w.y;
        wcs.x = m2c.x > m1c.x ? -wc.x : wc.x;
        wcs.y = m2c.y > m1c.y ? -wc.y : wc.y;
        Tx = __ffma2_rn(ws, rx, Tx);
        Ty = __ffma2_rn(ws, ry, Ty);
        cTx = FIRST ? __fmul2_rn(wcs, rx) : __ffma2_rn(wcs, rx, cTx);
        cTy = FIRST ? __fmul2_rn(wcs, ry) : __ffma2_rn(wcs, ry, cTy);
    } else if (FIRST) {
        cTx = cTy = make_float2(0.f, 0.f);
    }
}

// partialR[split][local row] = row-direction (Sx,Sy,Tx',Ty') over the split's block pairs;
// partialC[(I - I0) * D + d - 1][column of block J] = column-direction sums of block pair (I, J = I + d mod T).
// RPS = packed row pairs per thread (2 * RPS rows): the CTA's 512 rows are spread over THREADS = 256 / RPS threads.
template <int VERSION, int CT, int BATCH, int RPS = 2>
__global__ void __launch_bounds__(MS_BLOCK / (2 * RPS)) mlapm_sym_kernel(const float4 *__restrict__ rec, int T, int D,
                                                                         int I0, int per, int sub, M2Const k,
                                                                         float4 *__restrict__ partialR, int nrows_pad,
                                                                         unsigned long long *__restrict__ colsum,
                                                                         float colscale) {
    constexpr int THREADS = MS_BLOCK / (2 * RPS);
    static_assert(MS_BLOCK % CT == 0 && CT % BATCH == 0 && (BATCH == 8 || BATCH == 4), "bad stage shape");
    constexpr int SPB = MS_BLOCK / CT;                      // stages per block pair
    constexpr int LPC = 32 / BATCH;                         // lanes that share a column in the transposed sum
    extern __shared__ __align__(128) unsigned char ms_smem_raw[];
    MsSmem<CT, BATCH, THREADS> &sm = *reinterpret_cast<MsSmem<CT, BATCH, THREADS> *>(ms_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&sm.bars[0], 1);
        mbar_init(&sm.bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int I = I0 + blockIdx.x;
    int L = D + 1;                                          // d = 0 .. D
    if (!(T & 1) && 2 * I >= T) L = D;                      // even T: the pairs at d = T/2 belong to I < T/2
    // blockIdx.y = (group of `per` block pairs) * sub + (which 1/sub of each pair's column stages): small shards use
    // sub > 1 so that the grid still has several waves of CTAs
    const int sy = blockIdx.y / sub, su = blockIdx.y - sy * sub;
    const int spc = SPB / sub;                               // stages of a block pair this CTA handles
    const int d_lo = sy * per;
    const int d_hi = min(d_lo + per, L);

    float2 npx[RPS], npy[RPS], vx[RPS], vy[RPS], nvx[RPS], nvy[RPS], ex[RPS], ey[RPS], Sx[RPS], Sy[RPS], Tx[RPS],
        Ty[RPS];
#pragma unroll
    for (int i = 0; i < RPS; ++i) {
        const int64_t na = static_cast<int64_t>(I) * MS_BLOCK + (2 * i) * THREADS + tid;
        const int64_t nb = na + THREADS;
        const float4 a0 = rec[2 * na], a1 = rec[2 * na + 1], b0 = rec[2 * nb], b1 = rec[2 * nb + 1];
        npx[i] = make_float2(-a0.x, -b0.x); npy[i] = make_float2(-a0.y, -b0.y);
        vx[i] = make_float2(a0.z, b0.z); vy[i] = make_float2(a0.w, b0.w);
        nvx[i] = make_float2(-a0.z, -b0.z); nvy[i] = make_float2(-a0.w, -b0.w);
        ex[i] = make_float2(a1.x, b1.x); ey[i] = make_float2(a1.y, b1.y);
        Sx[i] = Sy[i] = Tx[i] = Ty[i] = make_float2(0.f, 0.f);
    }
    const float2 Bl = splat(k.Bl), Cl = splat(k.Cl), Dl = splat(k.Dl);

    const int nst = d_hi > d_lo ? (d_hi - d_lo) * spc : 0;
    constexpr uint32_t STAGE_BYTES = CT * MS_RECF * sizeof(float);
    auto stage_src = [&](int s) {
        int J = I + d_lo + s / spc;
        J = J >= T ? J - T : J;
        return rec + (static_cast<int64_t>(J) * MS_BLOCK + (su * spc + s % spc) * CT) * 2;
    };
    if (tid == 0 && nst > 0) {
        mbar_expect_tx(&sm.bars[0], STAGE_BYTES);
        tma_bulk_g2s(sm.tile[0], stage_src(0), STAGE_BYTES, &sm.bars[0]);
    }
    uint32_t phase_bits = 0;
    for (int s = 0; s < nst; ++s) {
        const int buf = s & 1;
        if (tid == 0 && s + 1 < nst) {                      // tile[buf^1] was released by the barrier below
            mbar_expect_tx(&sm.bars[buf ^ 1], STAGE_BYTES);
            tma_bulk_g2s(sm.tile[buf ^ 1], stage_src(s + 1), STAGE_BYTES, &sm.bars[buf ^ 1]);
        }
        mbar_wait(&sm.bars[buf], (phase_bits >> buf) & 1u);
        phase_bits ^= (1u << buf);
        const float4 *tl = sm.tile[buf];
        const int d = d_lo + s / spc;
        if (d == 0) {
            // diagonal block: rows and columns are the same agents, every ordered pair appears -> one direction each
#pragma unroll 4
            for (int j = 0; j < CT; ++j) {
                const float4 cp = tl[2 * j];
#pragma unroll
                for (int i = 0; i < RPS; ++i)
                    pair2<VERSION>(npx[i], npy[i], vx[i], vy[i], nvx[i], nvy[i], ex[i], ey[i], cp, Bl, Cl, Dl, Sx[i],
                                   Sy[i], Tx[i], Ty[i]);
            }
        } else {
            float4 *red = sm.red[warp];
            for (int b = 0; b < CT / BATCH; ++b) {
#pragma unroll
                for (int c = 0; c < BATCH; ++c) {
                    const float4 cp = tl[2 * (b * BATCH + c)];
                    const float4 ce = tl[2 * (b * BATCH + c) + 1];
                    float2 cSx, cSy, cTx, cTy;
                    pair2sym<VERSION, true>(npx[0], npy[0], vx[0], vy[0], nvx[0], nvy[0], ex[0], ey[0], cp, ce, Bl, Cl,
                                            Dl, Sx[0], Sy[0], Tx[0], Ty[0], cSx, cSy, cTx, cTy);
#pragma unroll
                    for (int i = 1; i < RPS; ++i)
                        pair2sym<VERSION, false>(npx[i], npy[i], vx[i], vy[i], nvx[i], nvy[i], ex[i], ey[i], cp, ce,
                                                 Bl, Cl, Dl, Sx[i], Sy[i], Tx[i], Ty[i], cSx, cSy, cTx, cTy);
                    red[c * MS_RED_STRIDE + lane] = make_float4(cSx.x + cSx.y, cSy.x + cSy.y, cTx.x + cTx.y,
                                                                cTy.x + cTy.y);
                }
                __syncwarp();
                // lane -> column lane/LPC, entries (lane%LPC), (lane%LPC)+LPC, ...: 32/LPC of the 32 row lanes each
                const float4 *col = red + (lane / LPC) * MS_RED_STRIDE + (lane % LPC);
                float4 acc = col[0];
#pragma unroll
                for (int e = 1; e < 32 / LPC; ++e) {
                    const float4 t = col[LPC * e];
                    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                }
#pragma unroll
                for (int o = 1; o < LPC; o <<= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
                    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
                }
                if (lane % LPC == 0) sm.colacc[warp][b * BATCH + lane / LPC] = acc;
                __syncwarp();
            }
            __syncthreads();
            // Column-direction sums of this stage's columns (block J), over the CTA's 512 rows: added into the crowd-wide
            // FIXED-POINT accumulators colsum[agent][4] (int64, 2^-s resolution) with fire-and-forget 64-bit atomics.
            // Integer addition is associative, so the total does not depend on the order in which the T/2 block pairs
            // that share a column arrive: deterministic like the per-pair partial arrays this replaces, with O(N)
            // instead of O(N^2 / 64) workspace and DRAM traffic.
            int J = I + d;
            J = J >= T ? J - T : J;
            unsigned long long *dst = colsum + (static_cast<int64_t>(J) * MS_BLOCK + (su * spc + s % spc) * CT) * 4;
            for (int c = tid; c < CT; c += THREADS) {
                float4 a = sm.colacc[0][c];
#pragma unroll
                for (int wv = 1; wv < THREADS / 32; ++wv) {
                    const float4 t = sm.colacc[wv][c];
                    a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
                }
                atomicAdd(dst + c * 4 + 0, static_cast<unsigned long long>(__float2ll_rn(a.x * colscale)));
                atomicAdd(dst + c * 4 + 1, static_cast<unsigned long long>(__float2ll_rn(a.y * colscale)));
                atomicAdd(dst + c * 4 + 2, static_cast<unsigned long long>(__float2ll_rn(a.z * colscale)));
                atomicAdd(dst + c * 4 + 3, static_cast<unsigned long long>(__float2ll_rn(a.w * colscale)));
            }
        }
        __syncthreads();
    }
    float4 *out = partialR + static_cast<int64_t>(blockIdx.y) * nrows_pad + static_cast<int64_t>(blockIdx.x) * MS_BLOCK;
#pragma unroll
    for (int i = 0; i < RPS; ++i) {
        out[(2 * i) * THREADS + tid] = make_float4(Sx[i].x, Sy[i].x, Tx[i].x, Ty[i].x);
        out[(2 * i + 1) * THREADS + tid] = make_float4(Sx[i].y, Sy[i].y, Tx[i].y, Ty[i].y);
    }
}

// Exchange step of an agent-sharded crowd fused into the finalize kernel: rank g's NEXT-state arrays, mapped into this
// process over NVLink peer memory.  world == 0: no exchange.
constexpr int ML_MAX_PEERS = 16;
struct PeerPush { int world; float2 *pos[ML_MAX_PEERS]; float2 *vel[ML_MAX_PEERS]; };
// Column-direction sums of the symmetric kernel (colsum == nullptr: ordered-pair kernel, row sums only): fixed-point
// int64 accumulators per agent, value = colsum * inv_colscale.
// inbox != nullptr (agent-sharded crowd): the column-direction sums were reduced per rank and delivered to the owner
// of the rows (mlapm_sym_colpush_kernel); inbox[g * inbox_stride + local row] is rank g's share.
struct SymPartials { const long long *colsum; float inv_colscale; int nrows_pad; const float4 *inbox; int world;
                     int64_t inbox_stride; };
// Owners of the 512-agent blocks (rank g owns blocks [Ib[g], Ib[g+1])) and every rank's inbox as mapped here.
struct PeerInbox { int world, rank; int Ib[ML_MAX_PEERS + 1]; float4 *inbox[ML_MAX_PEERS]; int64_t stride; };

// Agent-sharded symmetric evaluation, first half of the exchange: this rank evaluated the block pairs (I, I + d) of
// ITS row blocks I in [I0, I1); the column-direction sum of every agent m (its own or another rank's) over those block
// pairs sits in this rank's fixed-point accumulators and is stored, as fp32, into the inbox of the rank that owns m,
// over NVLink peer memory.
__global__ void mlapm_sym_colpush_kernel(const long long *__restrict__ colsum, float inv_colscale, int N,
                                         const __grid_constant__ PeerInbox peers) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < N) {
        const int J = m / MS_BLOCK;
        const longlong2 c0 = reinterpret_cast<const longlong2 *>(colsum)[2 * m];
        const longlong2 c1 = reinterpret_cast<const longlong2 *>(colsum)[2 * m + 1];
        const double is = static_cast<double>(inv_colscale);
        const float4 u = make_float4(static_cast<float>(static_cast<double>(c0.x) * is),
                                     static_cast<float>(static_cast<double>(c0.y) * is),
                                     static_cast<float>(static_cast<double>(c1.x) * is),
                                     static_cast<float>(static_cast<double>(c1.y) * is));
        int h = 0;
        while (h + 1 < peers.world && J >= peers.Ib[h + 1]) ++h;
        peers.inbox[h][peers.rank * peers.stride + (m - peers.Ib[h] * MS_BLOCK)] = u;
    }
    __syncthreads();                                                      // one cumulative system fence per CTA:
    if (threadIdx.x == 0) __threadfence_system();                         // visible to the owner before the barrier
}

constexpr int FZ_LPR = 8;                        // lanes per row in the finalize kernel

// force = (v0*ed - v)/tau - A*R(sum partial) ; action = v + force*dt ; optional p' = p + action*dt and arrival.
__global__ void mlapm_finalize2_kernel(const float2 *__restrict__ pos, const float2 *__restrict__ vel,
                                       const float *__restrict__ ds, int ds_dim, const float2 *__restrict__ dest,
                                       int row0, int row1, int nsplit, const float4 *__restrict__ partial, float A,
                                       float cos_t, float sin_t, int version, float tau, float dt, float radius,
                                       float2 *__restrict__ action, float2 *__restrict__ pos_new,
                                       uint8_t *__restrict__ arrived, const __grid_constant__ PeerPush push,
                                       const __grid_constant__ SymPartials sym) {
    // FZ_LPR lanes share a row: lane j adds the split partials j, j + FZ_LPR, ... and the lanes' sums are combined in
    // lane order -- a fixed order (deterministic), but FZ_LPR loads in flight per row instead of one dependent chain of
    // up to 396 (an 8-way shard of 100k agents): the kernel was latency bound at 70 us of a 1.05 ms step.
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int rl_raw = gt / FZ_LPR, sub = gt % FZ_LPR;
    const int nrows = row1 - row0;
    const bool live = rl_raw < nrows;
    const int rl = live ? rl_raw : nrows - 1;                     // idle lanes shadow the last row (shuffles stay converged)
    const int64_t pstride = (sym.colsum || sym.inbox) ? sym.nrows_pad : nrows;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = sub; q < nsplit; q += FZ_LPR) {
        const float4 t = partial[static_cast<int64_t>(q) * pstride + rl];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
#pragma unroll
    for (int o = 1; o < FZ_LPR; o <<= 1) {                        // tree in a fixed shape: deterministic
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, o);
        s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
    }
    const int n = row0 + rl;
    float qx = 0.f, qy = 0.f, ax = 0.f, ay = 0.f;
    if (live && sub == 0) {
    const float2 p = pos[n], v = vel[n], d = dest[n];
    const float dx = __fsub_rn(d.x, p.x), dy = __fsub_rn(d.y, p.y);
    const float dn = fmaxf(norm2_rn(dx, dy), 1e-12f);
    const float ex = __fdiv_rn(dx, dn), ey = __fdiv_rn(dy, dn);
    const float dsx = ds[static_cast<int64_t>(n) * ds_dim];
    const float dsy = ds[static_cast<int64_t>(n) * ds_dim + (ds_dim > 1 ? 1 : 0)];
    // force += (desired_speed * ed - velocity) / tau      (mlapm.py:22)
    float fx = __fdiv_rn(__fsub_rn(__fmul_rn(dsx, ex), v.x), tau);
    float fy = __fdiv_rn(__fsub_rn(__fmul_rn(dsy, ey), v.y), tau);
    if (sym.inbox) {
        // agent-sharded symmetric evaluation: one pre-reduced share per rank, added in rank order
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int g = 0; g < sym.world; ++g) {
            const float4 t = sym.inbox[g * sym.inbox_stride + rl];
            u.x += t.x; u.y += t.y; u.z += t.z; u.w += t.w;
        }
        s.x -= u.x; s.y -= u.y; s.z -= u.z; s.w -= u.w;
    } else if (sym.colsum) {
        // symmetric evaluation: this agent was a COLUMN of T/2 block pairs; their column-direction sums (exact integer
        // total of the per-pair fp32 block sums) enter with the opposite sign (vr[m,n] = -vr[n,m]).
        const longlong2 c0 = reinterpret_cast<const longlong2 *>(sym.colsum)[2 * n];
        const longlong2 c1 = reinterpret_cast<const longlong2 *>(sym.colsum)[2 * n + 1];
        const double is = static_cast<double>(sym.inv_colscale);
        s.x -= static_cast<float>(static_cast<double>(c0.x) * is);
        s.y -= static_cast<float>(static_cast<double>(c0.y) * is);
        s.z -= static_cast<float>(static_cast<double>(c1.x) * is);
        s.w -= static_cast<float>(static_cast<double>(c1.y) * is);
    }
    float sx, sy;
    if (version == 0) { sx = s.x; sy = s.y; }
    else {                                                        // R(sigma theta) r^ summed    (mlapm.py:35-38)
        sx = fmaf(cos_t, s.x, -(sin_t * s.w));
        sy = fmaf(sin_t, s.z, cos_t * s.y);
    }
    fx = __fsub_rn(fx, __fmul_rn(A, sx));                         // force -= (...).sum(dim=1)   (mlapm.py:29/39)
    fy = __fsub_rn(fy, __fmul_rn(A, sy));
    ax = __fadd_rn(v.x, __fmul_rn(fx, dt));                       // action = velocity + force*dt (mlapm.py:57)
    ay = __fadd_rn(v.y, __fmul_rn(fy, dt));
    if (action) action[rl] = make_float2(ax, ay);
    if (pos_new || arrived || push.world > 0) {
        qx = __fadd_rn(p.x, __fmul_rn(ax, dt));                   // p = position + v*dt          (main_mlapm.py:26)
        qy = __fadd_rn(p.y, __fmul_rn(ay, dt));
        if (pos_new) pos_new[rl] = make_float2(qx, qy);
        if (arrived)                                              // ||p - destination|| < radius (main_mlapm.py:34)
            arrived[rl] = norm2_rn(__fsub_rn(qx, d.x), __fsub_rn(qy, d.y)) < radius ? 1 : 0;
    }
    }
    if (push.world > 0) {
        // the path's one exchange, fused: the row's new state goes into EVERY rank's next-state arrays; the row's
        // FZ_LPR lanes take one peer each, so the remote stores of a row are issued in parallel
        const int lead = (threadIdx.x & 31) & ~(FZ_LPR - 1);
        qx = __shfl_sync(0xffffffffu, qx, lead); qy = __shfl_sync(0xffffffffu, qy, lead);
        ax = __shfl_sync(0xffffffffu, ax, lead); ay = __shfl_sync(0xffffffffu, ay, lead);
        if (live)
            for (int g = sub; g < push.world; g += FZ_LPR) {
                push.pos[g][n] = make_float2(qx, qy);
                push.vel[g][n] = make_float2(ax, ay);
            }
        // one system-scope fence per CTA (fences are cumulative over the CTA barrier): the stores are visible to the
        // peers before the step barrier that follows the kernel
        __syncthreads();
        if (threadIdx.x == 0) __threadfence_system();
    }
}

constexpr int ML_MAX_SPLIT = 64;

static int pick_rows_per_thread(int64_t nrows) { return nrows >= 4096 ? 4 : (nrows >= 1024 ? 2 : 1); }

// Row pairs per thread of the packed kernel (2*RP rows per thread).
// 4 rows per thread (RP = 2) is ~0.5 % faster per pair, but halves the CTA count: with few row blocks (an 8-way
// shard of 100k agents has 25) the last wave is mostly idle, so small grids use 2 rows per thread.  Does not change
// any result (each row's column order is the same).
static int pick_row_pairs(int64_t nrows, int64_t N) {
    const int64_t tiles = (N + 511) / 512;
    const int64_t nsplit = tiles < 64 ? tiles : 64;
    const int64_t ctas2 = ((nrows + 511) / 512) * nsplit;
    return (nrows >= 8192 && ctas2 >= 10LL * 4 * sm_count()) ? 2 : 1;   // measured: 12.5k..50k rows of 100k prefer 1
}

// Column splits: enough CTAs for ~8 waves over the SMs, each split a multiple of `tile` columns.
static void pick_split(int64_t nrows, int64_t N, int rows_per_cta, int tile, int *nsplit, int *cols_per_split) {
    const int64_t row_blocks = (nrows + rows_per_cta - 1) / rows_per_cta;
    const int64_t want_ctas = 8LL * 6 * sm_count();
    int64_t s = (want_ctas + row_blocks - 1) / row_blocks;
    const int64_t max_by_cols = (N + tile - 1) / tile;
    if (s > max_by_cols) s = max_by_cols;
    if (s > ML_MAX_SPLIT) s = ML_MAX_SPLIT;
    if (s < 1) s = 1;
    int64_t cps = (N + s - 1) / s;
    cps = (cps + tile - 1) / tile * tile;
    *cols_per_split = static_cast<int>(cps);
    *nsplit = static_cast<int>((N + cps - 1) / cps);
}

// Column splits for the packed kernel: enough CTAs for >= 8 waves over the resident slots (occupancy x SMs), each
// split a whole number of tiles.  Measured at N = 100k (profiles/r01_mlapm_experiments.md): 27-36 splits (6-8 waves)
// are ~2% faster than the exact 2- or 4-wave fits (9, 18), 4 splits (1 wave) 18% slower.  PIML_MLAPM_SPLIT overrides.
static int pick_split2(const void *kernel, int64_t nrows, int64_t N, int rows_per_cta, int *nsplit,
                       int *cols_per_split) {
    static const void *cached_kernel = nullptr;
    static int cached_occ = 0;
    if (kernel != cached_kernel) {
        int occ = 0;
        PIML_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, M2_THREADS, 0));
        cached_occ = occ < 1 ? 1 : occ;
        cached_kernel = kernel;
    }
    (void)nrows; (void)rows_per_cta;
    // The column partition depends on N ONLY: a row's sum is then evaluated in the same order whichever row range
    // (rank) it is computed in, so an agent-sharded crowd is bit-identical to the unsharded one.
    // As many splits as allowed also gives every rank of an 8-way shard enough CTAs to fill its GPU.
    const int64_t tiles = (N + M2_TILE - 1) / M2_TILE;
    int64_t s = tiles < ML_MAX_SPLIT ? tiles : ML_MAX_SPLIT;
    if (s < 1) s = 1;
    if (const char *e = getenv("PIML_MLAPM_SPLIT")) {
        const int64_t f = atoll(e);
        if (f >= 1 && f <= ML_MAX_SPLIT && f <= tiles) s = f;
    }
    const int64_t per = (tiles + s - 1) / s;
    *cols_per_split = static_cast<int>(per * M2_TILE);
    *nsplit = static_cast<int>((tiles + per - 1) / per);
    return PIML_OK;
}

template <int VERSION, bool EXACT>
static void launch_pairs(int R, dim3 grid, cudaStream_t st, const float2 *pos, const float2 *vel, const float2 *dest,
                         int N, int row0, int row1, int cps, const MlConst &k, float2 *partial) {
    if (R == 4)
        mlapm_pairs_kernel<VERSION, 4, EXACT><<<grid, ML_THREADS, 0, st>>>(pos, vel, dest, N, row0, row1, cps, k, partial);
    else if (R == 2)
        mlapm_pairs_kernel<VERSION, 2, EXACT><<<grid, ML_THREADS, 0, st>>>(pos, vel, dest, N, row0, row1, cps, k, partial);
    else
        mlapm_pairs_kernel<VERSION, 1, EXACT><<<grid, ML_THREADS, 0, st>>>(pos, vel, dest, N, row0, row1, cps, k, partial);
}

}  // namespace piml

using namespace piml;

// ---- symmetric evaluation: selection and workspace ------------------------------------------------------------------
static int g_mlapm_algorithm = 0;                 // 0 automatic, 1 ordered pairs (v2), 2 symmetric (v3)
constexpr int64_t MS_AUTO_MIN_AGENTS = 16384;     // below this the ordered-pair kernel fills the GPU better

static int64_t sym_blocks(int64_t N) { return (N + MS_BLOCK - 1) / MS_BLOCK; }
// Block pairs per CTA for nI row blocks of a crowd of T blocks: >= 64 CTAs per SM (measured best at N = 100k on one
// GPU: 2; an 8-way shard of the same crowd gets 1).
static int sym_per(int64_t nI, int64_t T) {
    if (const char *e = getenv("PIML_MLAPM_SYM_PER")) {
        const int f = atoi(e);
        if (f >= 1) return f;
    }
    const int64_t pairs = nI * (T / 2 + 1), want = 64LL * sm_count();
    const int64_t per = pairs / want;
    return per < 1 ? 1 : static_cast<int>(per);
}
// CTAs per block pair (1, 2 or 4 = the 4 column stages of a pair split over that many CTAs): a shard with few row
// blocks (8 ranks at N = 100k: 25 blocks x 99 pairs) would otherwise run ~2 waves of CTAs with a nearly empty last one.
static int sym_sub(int64_t nI, int64_t T, int per) {
    if (const char *e = getenv("PIML_MLAPM_SYM_SUB")) {
        const int f = atoi(e);
        if (f == 1 || f == 2 || f == 4) return f;
    }
    const int64_t ctas = nI * ((T / 2 + 1 + per - 1) / per), wave = 8LL * sm_count();
    int sub = 1;
    while (sub < MS_BLOCK / MS_CT && ctas * sub < 6 * wave) sub *= 2;
    return sub;
}
// Row-partial splits of a launch over nI row blocks: (groups of `per` block pairs) x sub.
static int sym_splits(int64_t nI, int64_t T, int *per, int *sub) {
    *per = sym_per(nI, T);
    *sub = sym_sub(nI, T, *per);
    return static_cast<int>((T / 2 + 1 + *per - 1) / *per) * *sub;
}

// records (32 B per agent) + row-direction partials (16 B per agent and split) + column-direction fixed-point
// accumulators (4 x int64 per agent): O(N) -- 0.15 GB at N = 10^6 (the per-pair column partials this replaces: 15.6 GB)
static int64_t sym_workspace_bytes(int64_t N) {
    const int64_t T = sym_blocks(N), npad = T * MS_BLOCK;
    int per_, sub_;
    const int64_t S = sym_splits(T, T, &per_, &sub_);
    return npad * MS_RECF * sizeof(float) + S * npad * 4 * sizeof(float) + npad * 4 * sizeof(long long) + 256;
}

// Fixed-point scale 2^s of the column accumulators: one pair contributes |w r| = exp(arg) <= exp(|C|) (B r + D r cos <= 0
// is required by sym_params_ok), a column at most N of them; keep two bits of headroom below 2^63.
static int sym_colscale_log2(int64_t N, const piml_mlapm_params *prm) {
    const double c = prm->version == 0 ? 0.0 : fabs(static_cast<double>(prm->C));
    const double bound = static_cast<double>(N) * exp(c);
    return 61 - static_cast<int>(ceil(log2(bound > 1.0 ? bound : 1.0)));
}

extern "C" int piml_set_mlapm_algorithm(int algo) {
    PIML_REQUIRE(algo >= 0 && algo <= 2,
                 "piml_set_mlapm_algorithm: 0 = automatic, 1 = ordered pairs, 2 = symmetric (unordered pairs)");
    g_mlapm_algorithm = algo;
    return PIML_OK;
}

extern "C" int64_t piml_mlapm_workspace_bytes_sym(int64_t N) {
    if (N <= 0) return 0;
    const int64_t a = piml_mlapm_workspace_bytes(N), b = sym_workspace_bytes(N);
    return a > b ? a : b;
}

extern "C" int64_t piml_mlapm_workspace_bytes(int64_t N) {
    if (N < 0) return 0;
    // column records (32 B per agent, padded to a tile) + per-split partial sums (16 B per row and split)
    const int64_t npad = (N + M2_TILE - 1) / M2_TILE * M2_TILE;
    return npad * M2_COLF * sizeof(float) + static_cast<int64_t>(ML_MAX_SPLIT) * N * 4 * sizeof(float) + 256;
}

static MlConst mlapm_consts(const piml_mlapm_params *prm) {
    MlConst k;
    k.version = prm->version;
    k.A = prm->A; k.B = prm->B; k.C = prm->C; k.D = prm->D;
    const double log2e = 1.4426950408889634;
    k.Bl = static_cast<float>(prm->B * log2e); k.Cl = static_cast<float>(prm->C * log2e);
    k.Dl = static_cast<float>(prm->D * log2e);
    // theta tensor as the reference builds it in fp32: ((+-1 * theta) / 180) * pi   (mlapm.py:33)
    const float th = (prm->theta_deg / 180.0f) * 3.14159274101257324f;
    k.cos_t = static_cast<float>(cos(static_cast<double>(th)));
    k.sin_t = static_cast<float>(sin(static_cast<double>(th)));
    k.tau = prm->tau; k.inv_tau = 1.0f / prm->tau;
    return k;
}

// The padding agents of the symmetric kernel need a weight that underflows to 0 at large r.
static bool sym_params_ok(const piml_mlapm_params *prm, const MlConst &k, int64_t N) {
    return !prm->exact_math && (prm->version == 0 ? k.Bl < 0.f : k.Bl + fabsf(k.Dl) < 0.f) &&
           sym_colscale_log2(N, prm) >= 30;                        // 2^-30 resolution at the very least (|C| < ~15)
}

// prep + symmetric pair kernel for the row blocks [I0, I0 + nI) of a crowd of T blocks.
static int launch_sym_pairs(int version, const float2 *p2, const float2 *v2, const float2 *d2, int N, int64_t T,
                            int64_t I0, int64_t nI, int per, int sub, int S, const MlConst &k, float4 *rec, float4 *partialR,
                            long long *colsum, float colscale, cudaStream_t st) {
    const int64_t D = T / 2, npad = T * MS_BLOCK;
    const int threads = 256;
    mlapm_prep_sym_kernel<<<static_cast<unsigned>((npad + threads - 1) / threads), threads, 0, st>>>(
        p2, v2, d2, N, static_cast<int>(npad), rec, reinterpret_cast<ulonglong2 *>(colsum));
    count_launch();
    int rc = check_launch("mlapm_prep_sym_kernel");
    if (rc) return rc;
    M2Const k2{k.Bl, k.Cl, k.Dl};
    dim3 grid(static_cast<unsigned>(nI), static_cast<unsigned>(S));
    int cfg = 0;                                               // tuning: PIML_MLAPM_SYM_CFG = 0..3
    if (const char *e = getenv("PIML_MLAPM_SYM_CFG")) cfg = atoi(e);
#define PIML_LAUNCH_SYM(V, CT, BATCH, RPS)                                                                         \
    do {                                                                                                           \
        using Smem = MsSmem<CT, BATCH, MS_BLOCK / (2 * RPS)>;                                                      \
        if (sizeof(Smem) > 48 * 1024)                    /* per device: a process may drive several GPUs */        \
            PIML_CUDA(cudaFuncSetAttribute(mlapm_sym_kernel<V, CT, BATCH, RPS>,                                    \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,                            \
                                           static_cast<int>(sizeof(Smem))));                                       \
        mlapm_sym_kernel<V, CT, BATCH, RPS><<<grid, MS_BLOCK / (2 * RPS), sizeof(Smem), st>>>(                     \
            rec, static_cast<int>(T), static_cast<int>(D), static_cast<int>(I0), per, sub, k2, partialR,           \
            static_cast<int>(nI * MS_BLOCK), reinterpret_cast<unsigned long long *>(colsum), colscale);            \
    } while (0)
    // measured at N = 100k (ms per step): 8 rows per thread (64 threads) 6.81, 4 rows (128 threads) 6.99, 2 rows 7.16
    if (version == 0) PIML_LAUNCH_SYM(0, MS_CT, MS_BATCH, 2);
    else if (cfg == 2) PIML_LAUNCH_SYM(1, 128, 8, 2);
    else if (cfg == 4) PIML_LAUNCH_SYM(1, 128, 8, 1);
    else PIML_LAUNCH_SYM(1, MS_CT, MS_BATCH, 4);
#undef PIML_LAUNCH_SYM
    count_launch();
    return check_launch("mlapm_sym_kernel");
}

static int mlapm_advance_impl(const float *pos, const float *vel, const float *desired_speed, int ds_dim,
                              const float *dest, int64_t N, int64_t row0, int64_t row1, const piml_mlapm_params *prm,
                              float dt, float radius, float *action, float *pos_new, uint8_t *arrived, void *workspace,
                              int64_t workspace_bytes, const PeerPush &push, void *stream) {
    if (N == 0) return PIML_OK;                                    // empty crowd: nothing to do, pointers may be null
    PIML_REQUIRE(pos && vel && desired_speed && dest && prm && (action || push.world > 0) && workspace,
                 "piml_mlapm: null pointer");
    PIML_REQUIRE(ds_dim == 1 || ds_dim == 2, "piml_mlapm: desired_speed must be (N,1) or (N,2), got ds_dim=%d", ds_dim);
    PIML_REQUIRE(N >= 0 && N < (1LL << 31), "piml_mlapm: N=%lld out of range", static_cast<long long>(N));
    PIML_REQUIRE(0 <= row0 && row0 <= row1 && row1 <= N, "piml_mlapm: bad row range [%lld,%lld) for N=%lld",
                 static_cast<long long>(row0), static_cast<long long>(row1), static_cast<long long>(N));
    PIML_REQUIRE(prm->version == 0 || prm->version == 1,
                 "piml_mlapm: version %d unsupported (0='raw', 1='GC'; 'UCY' is not runnable in the reference)",
                 prm->version);
    PIML_REQUIRE((reinterpret_cast<uintptr_t>(pos) & 7u) == 0 && (reinterpret_cast<uintptr_t>(vel) & 7u) == 0 &&
                     (reinterpret_cast<uintptr_t>(dest) & 7u) == 0 && (reinterpret_cast<uintptr_t>(action) & 7u) == 0,
                 "piml_mlapm: pointers must be 8-byte aligned");
    const int64_t nrows = row1 - row0;
    if (nrows == 0) return PIML_OK;

    const MlConst k = mlapm_consts(prm);

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float2 *p2 = reinterpret_cast<const float2 *>(pos), *v2 = reinterpret_cast<const float2 *>(vel);
    const float2 *d2 = reinterpret_cast<const float2 *>(dest);
    const int iN = static_cast<int>(N), r0 = static_cast<int>(row0), r1 = static_cast<int>(row1);
    int rc;
    if (prm->exact_math) {
        PIML_REQUIRE(push.world == 0, "piml_mlapm: the fused exchange is only available on the production math path");
        // validation variant: IEEE arithmetic in the reference's operation order
        const int R = pick_rows_per_thread(nrows);
        int nsplit, cps;
        pick_split(nrows, N, ML_THREADS * R, ML_TILE, &nsplit, &cps);
        dim3 grid(static_cast<unsigned>((nrows + ML_THREADS * R - 1) / (ML_THREADS * R)),
                  static_cast<unsigned>(nsplit));
        float2 *partial = reinterpret_cast<float2 *>(workspace);
        if (prm->version == 0) launch_pairs<0, true>(R, grid, st, p2, v2, d2, iN, r0, r1, cps, k, partial);
        else launch_pairs<1, true>(R, grid, st, p2, v2, d2, iN, r0, r1, cps, k, partial);
        count_launch();
        rc = check_launch("mlapm_pairs_kernel");
        if (rc) return rc;
        const int threads = 256;
        mlapm_finalize_kernel<<<static_cast<unsigned>((nrows + threads - 1) / threads), threads, 0, st>>>(
            p2, v2, desired_speed, ds_dim, d2, r0, r1, nsplit, partial, prm->tau, dt, radius,
            reinterpret_cast<float2 *>(action), reinterpret_cast<float2 *>(pos_new), arrived);
        count_launch();
        return check_launch("mlapm_finalize_kernel");
    }
    // production, whole crowd: symmetric evaluation (every unordered pair once) when the caller's workspace holds
    // the column-direction sums; row ranges (agent-sharded ranks) and small crowds use the ordered-pair kernel.
    SymPartials symp{nullptr, 0.f, 0, nullptr, 0, 0};
    const bool sym_ok = row0 == 0 && row1 == N && workspace_bytes >= sym_workspace_bytes(N) && sym_params_ok(prm, k, N);
    if (sym_ok && (g_mlapm_algorithm == 2 || (g_mlapm_algorithm == 0 && N >= MS_AUTO_MIN_AGENTS))) {
        const int64_t T = sym_blocks(N), D = T / 2, npad = T * MS_BLOCK;
        int per, sub;
        const int S = sym_splits(T, T, &per, &sub);
        float4 *rec = reinterpret_cast<float4 *>(workspace);
        float4 *partialR = rec + npad * 2;
        long long *colsum = reinterpret_cast<long long *>(partialR + static_cast<int64_t>(S) * npad);
        const int sl = sym_colscale_log2(N, prm);
        rc = launch_sym_pairs(prm->version, p2, v2, d2, iN, T, 0, T, per, sub, S, k, rec, partialR, colsum,
                              ldexpf(1.0f, sl), st);
        if (rc) return rc;
        const int threads = 256;
        (void)D;
        symp = SymPartials{colsum, ldexpf(1.0f, -sl), static_cast<int>(npad), nullptr, 0, 0};
        mlapm_finalize2_kernel<<<static_cast<unsigned>((nrows * FZ_LPR + threads - 1) / threads), threads, 0, st>>>(
            p2, v2, desired_speed, ds_dim, d2, r0, r1, S, partialR, prm->A, k.cos_t, k.sin_t, prm->version, prm->tau,
            dt, radius, reinterpret_cast<float2 *>(action), reinterpret_cast<float2 *>(pos_new), arrived, push, symp);
        count_launch();
        return check_launch("mlapm_finalize2_kernel");
    }
    // ordered pairs: packed-FP32 kernel on 16 B column records
    const int64_t npad = (N + M2_TILE - 1) / M2_TILE * M2_TILE;
    float4 *col8 = reinterpret_cast<float4 *>(workspace);
    float4 *partial4 = col8 + npad;
    {
        const int threads = 256;
        mlapm_prep_kernel<<<static_cast<unsigned>((npad + threads - 1) / threads), threads, 0, st>>>(
            p2, v2, iN, static_cast<int>(npad), col8);
        count_launch();
        rc = check_launch("mlapm_prep_kernel");
        if (rc) return rc;
    }
    int RP = pick_row_pairs(nrows, N), unroll = 4;
    if (const char *e = getenv("PIML_MLAPM_EXP")) sscanf(e, "%d,%d", &RP, &unroll);            // tuning experiments
    M2Const k2{k.Bl, k.Cl, k.Dl};
    dim3 grid;
    int nsplit = 1, cps = 0, rc2 = PIML_OK;
#define PIML_LAUNCH_PAIRS2(V, RPV, UN)                                                                             \
    do {                                                                                                           \
        rc2 = pick_split2(reinterpret_cast<const void *>(&mlapm_pairs2_kernel<V, RPV, UN>), nrows, N,              \
                          M2_THREADS * 2 * RPV, &nsplit, &cps);                                                    \
        grid = dim3(static_cast<unsigned>((nrows + M2_THREADS * 2 * RPV - 1) / (M2_THREADS * 2 * RPV)),            \
                    static_cast<unsigned>(nsplit));                                                                \
        if (rc2 == PIML_OK)                                                                                        \
            mlapm_pairs2_kernel<V, RPV, UN><<<grid, M2_THREADS, 0, st>>>(p2, v2, d2, col8, iN, r0, r1, cps, k2,    \
                                                                         partial4);                                \
    } while (0)
    if (prm->version == 0) {
        if (RP == 2) PIML_LAUNCH_PAIRS2(0, 2, 4); else PIML_LAUNCH_PAIRS2(0, 1, 4);
    } else if (RP == 4) {
        PIML_LAUNCH_PAIRS2(1, 4, 2);
    } else if (RP == 2) {
        if (unroll == 2) PIML_LAUNCH_PAIRS2(1, 2, 2); else PIML_LAUNCH_PAIRS2(1, 2, 4);
    } else {
        if (unroll == 2) PIML_LAUNCH_PAIRS2(1, 1, 2); else PIML_LAUNCH_PAIRS2(1, 1, 4);
    }
    if (rc2) return rc2;
#undef PIML_LAUNCH_PAIRS2
    count_launch();
    rc = check_launch("mlapm_pairs2_kernel");
    if (rc) return rc;
    const int threads = 256;
    mlapm_finalize2_kernel<<<static_cast<unsigned>((nrows * FZ_LPR + threads - 1) / threads), threads, 0, st>>>(
        p2, v2, desired_speed, ds_dim, d2, r0, r1, nsplit, partial4, prm->A, k.cos_t, k.sin_t, prm->version, prm->tau,
        dt, radius, reinterpret_cast<float2 *>(action), reinterpret_cast<float2 *>(pos_new), arrived, push, symp);
    count_launch();
    return check_launch("mlapm_finalize2_kernel");
}

extern "C" int piml_mlapm_advance_f32(const float *pos, const float *vel, const float *desired_speed, int ds_dim,
                                      const float *dest, int64_t N, int64_t row0, int64_t row1,
                                      const piml_mlapm_params *prm, float dt, float radius, float *action,
                                      float *pos_new, uint8_t *arrived, void *workspace, void *stream) {
    PeerPush none;
    none.world = 0;
    return mlapm_advance_impl(pos, vel, desired_speed, ds_dim, dest, N, row0, row1, prm, dt, radius, action, pos_new,
                              arrived, workspace, 0, none, stream);
}

extern "C" int piml_mlapm_advance_ws_f32(const float *pos, const float *vel, const float *desired_speed, int ds_dim,
                                         const float *dest, int64_t N, int64_t row0, int64_t row1,
                                         const piml_mlapm_params *prm, float dt, float radius, float *action,
                                         float *pos_new, uint8_t *arrived, void *workspace, int64_t workspace_bytes,
                                         void *stream) {
    if (N == 0) return PIML_OK;
    PIML_REQUIRE(workspace_bytes >= piml_mlapm_workspace_bytes(N),
                 "piml_mlapm_advance_ws_f32: workspace of %lld bytes, need >= %lld",
                 static_cast<long long>(workspace_bytes), static_cast<long long>(piml_mlapm_workspace_bytes(N)));
    PeerPush none;
    none.world = 0;
    return mlapm_advance_impl(pos, vel, desired_speed, ds_dim, dest, N, row0, row1, prm, dt, radius, action, pos_new,
                              arrived, workspace, workspace_bytes, none, stream);
}

extern "C" int piml_mlapm_advance_push_f32(const float *pos, const float *vel, const float *desired_speed, int ds_dim,
                                           const float *dest, int64_t N, int64_t row0, int64_t row1,
                                           const piml_mlapm_params *prm, float dt, float radius, int world,
                                           const uint64_t *peer_pos_next_host, const uint64_t *peer_vel_next_host,
                                           uint8_t *arrived, void *workspace, void *stream) {
    PIML_REQUIRE(world >= 1 && world <= ML_MAX_PEERS, "piml_mlapm_advance_push_f32: world=%d not in [1,%d]", world,
                 ML_MAX_PEERS);
    PIML_REQUIRE(peer_pos_next_host && peer_vel_next_host, "piml_mlapm_advance_push_f32: null peer pointer table");
    PeerPush push;
    push.world = world;
    for (int g = 0; g < ML_MAX_PEERS; ++g) {
        push.pos[g] = g < world ? reinterpret_cast<float2 *>(static_cast<uintptr_t>(peer_pos_next_host[g])) : nullptr;
        push.vel[g] = g < world ? reinterpret_cast<float2 *>(static_cast<uintptr_t>(peer_vel_next_host[g])) : nullptr;
        PIML_REQUIRE(g >= world || (push.pos[g] && push.vel[g] && (peer_pos_next_host[g] & 7u) == 0 &&
                                    (peer_vel_next_host[g] & 7u) == 0),
                     "piml_mlapm_advance_push_f32: bad peer pointer for rank %d", g);
    }
    return mlapm_advance_impl(pos, vel, desired_speed, ds_dim, dest, N, row0, row1, prm, dt, radius, nullptr, nullptr,
                              arrived, workspace, 0, push, stream);
}

extern "C" int piml_mlapm_step_f32(const float *pos, const float *vel, const float *desired_speed, int ds_dim,
                                   const float *dest, int64_t N, int64_t row0, int64_t row1,
                                   const piml_mlapm_params *prm, float dt, float *action, void *workspace,
                                   void *stream) {
    return piml_mlapm_advance_f32(pos, vel, desired_speed, ds_dim, dest, N, row0, row1, prm, dt, 0.f, action,
                                  nullptr, nullptr, workspace, stream);
}

// ---- agent-sharded symmetric evaluation ------------------------------------------------------------------------------
// Rank g owns the 512-agent blocks [g T / G, (g+1) T / G) and evaluates the block pairs (I, I + d mod T) of its own
// row blocks.  Row-direction sums stay local; the column-direction sums of OTHER ranks' agents are reduced per rank and
// stored into the owner's inbox over NVLink peer memory (phase A).  After one barrier the owner's finalize kernel adds
// the ranks' shares in rank order and pushes the new state into every rank's next-state arrays (phase B); a second
// barrier ends the step.  Traffic per step and rank: 16 B per agent (shares) + 16 B per agent (state) to each peer.

static void sym_block_bounds(int64_t T, int world, int *Ib) {
    for (int g = 0; g <= world; ++g) Ib[g] = static_cast<int>(T * g / world);
}

extern "C" int piml_mlapm_sym_shard_rows(int64_t N, int world, int rank, int64_t *row0, int64_t *row1) {
    PIML_REQUIRE(N > 0 && world >= 1 && world <= ML_MAX_PEERS && rank >= 0 && rank < world && row0 && row1,
                 "piml_mlapm_sym_shard_rows: bad arguments (N=%lld, world=%d, rank=%d)", static_cast<long long>(N),
                 world, rank);
    const int64_t T = sym_blocks(N);
    PIML_REQUIRE(T >= world, "piml_mlapm_sym_shard_rows: %lld blocks of %d agents cannot be split over %d ranks",
                 static_cast<long long>(T), MS_BLOCK, world);
    int Ib[ML_MAX_PEERS + 1];
    sym_block_bounds(T, world, Ib);
    *row0 = static_cast<int64_t>(Ib[rank]) * MS_BLOCK;
    *row1 = static_cast<int64_t>(Ib[rank + 1]) * MS_BLOCK;
    if (*row1 > N) *row1 = N;
    return PIML_OK;
}

// Floats4 per (rank, row) slot of an inbox: the largest shard, so every rank's inbox has the same shape.
static int64_t sym_inbox_stride(int64_t N, int world) {
    const int64_t T = sym_blocks(N);
    return ((T + world - 1) / world) * MS_BLOCK;
}

extern "C" int64_t piml_mlapm_sym_inbox_bytes(int64_t N, int world) {
    if (N <= 0 || world < 1 || world > ML_MAX_PEERS) return 0;
    return sym_inbox_stride(N, world) * world * 4 * sizeof(float);
}

extern "C" int64_t piml_mlapm_sym_shard_workspace_bytes(int64_t N, int world) {
    if (N <= 0 || world < 1 || world > ML_MAX_PEERS) return 0;
    const int64_t T = sym_blocks(N), npad = T * MS_BLOCK;
    const int64_t nI = (T + world - 1) / world;
    int per_, sub_;
    const int64_t S = sym_splits(nI, T, &per_, &sub_);
    return npad * MS_RECF * sizeof(float) + S * nI * MS_BLOCK * 4 * sizeof(float) + npad * 4 * sizeof(long long) + 256;
}

struct SymShardPlan { int64_t T, D, npad, I0, nI; int per, sub, S; float4 *rec, *partialR; long long *colsum; };

static int sym_shard_plan(int64_t N, int world, int rank, void *workspace, int64_t workspace_bytes, SymShardPlan *pl) {
    PIML_REQUIRE(N > 0 && N < (1LL << 31) && world >= 1 && world <= ML_MAX_PEERS && rank >= 0 && rank < world,
                 "piml_mlapm_sym: bad shard (N=%lld, world=%d, rank=%d)", static_cast<long long>(N), world, rank);
    PIML_REQUIRE(workspace && workspace_bytes >= piml_mlapm_sym_shard_workspace_bytes(N, world),
                 "piml_mlapm_sym: workspace of %lld bytes, need >= %lld", static_cast<long long>(workspace_bytes),
                 static_cast<long long>(piml_mlapm_sym_shard_workspace_bytes(N, world)));
    pl->T = sym_blocks(N);
    PIML_REQUIRE(pl->T >= world, "piml_mlapm_sym: fewer blocks than ranks");
    pl->D = pl->T / 2;
    pl->npad = pl->T * MS_BLOCK;
    int Ib[ML_MAX_PEERS + 1];
    sym_block_bounds(pl->T, world, Ib);
    pl->I0 = Ib[rank];
    pl->nI = Ib[rank + 1] - Ib[rank];
    pl->S = sym_splits((pl->T + world - 1) / world, pl->T, &pl->per, &pl->sub);
    pl->rec = reinterpret_cast<float4 *>(workspace);
    pl->partialR = pl->rec + pl->npad * 2;
    // the largest shard's row partials, so that every rank's layout is the same
    pl->colsum = reinterpret_cast<long long *>(pl->partialR + static_cast<int64_t>(pl->S) * ((pl->T + world - 1) / world) * MS_BLOCK);
    return PIML_OK;
}

extern "C" int piml_mlapm_sym_pairs_push_f32(const float *pos, const float *vel, const float *dest, int64_t N,
                                             int world, int rank, const piml_mlapm_params *prm,
                                             const uint64_t *peer_inbox_host, void *workspace,
                                             int64_t workspace_bytes, void *stream) {
    PIML_REQUIRE(pos && vel && dest && prm && peer_inbox_host, "piml_mlapm_sym_pairs_push_f32: null pointer");
    PIML_REQUIRE(prm->version == 0 || prm->version == 1, "piml_mlapm_sym_pairs_push_f32: version %d unsupported",
                 prm->version);
    const MlConst k = mlapm_consts(prm);
    PIML_REQUIRE(sym_params_ok(prm, k, N), "piml_mlapm_sym_pairs_push_f32: parameters outside the symmetric kernel's "
                                            "domain (exact_math, or a weight that does not decay with distance)");
    SymShardPlan pl;
    int rc = sym_shard_plan(N, world, rank, workspace, workspace_bytes, &pl);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = launch_sym_pairs(prm->version, reinterpret_cast<const float2 *>(pos), reinterpret_cast<const float2 *>(vel),
                          reinterpret_cast<const float2 *>(dest), static_cast<int>(N), pl.T, pl.I0, pl.nI, pl.per, pl.sub, pl.S,
                          k, pl.rec, pl.partialR, pl.colsum, ldexpf(1.0f, sym_colscale_log2(N, prm)), st);
    if (rc) return rc;
    PeerInbox peers;
    peers.world = world;
    peers.rank = rank;
    sym_block_bounds(pl.T, world, peers.Ib);
    for (int g = world + 1; g <= ML_MAX_PEERS; ++g) peers.Ib[g] = peers.Ib[world];
    peers.stride = sym_inbox_stride(N, world);
    for (int g = 0; g < ML_MAX_PEERS; ++g) {
        peers.inbox[g] = g < world ? reinterpret_cast<float4 *>(static_cast<uintptr_t>(peer_inbox_host[g])) : nullptr;
        PIML_REQUIRE(g >= world || (peers.inbox[g] && (peer_inbox_host[g] & 15u) == 0),
                     "piml_mlapm_sym_pairs_push_f32: bad inbox pointer for rank %d", g);
    }
    const int threads = 256;
    mlapm_sym_colpush_kernel<<<static_cast<unsigned>((N + threads - 1) / threads), threads, 0, st>>>(
        pl.colsum, ldexpf(1.0f, -sym_colscale_log2(N, prm)), static_cast<int>(N), peers);
    count_launch();
    return check_launch("mlapm_sym_colpush_kernel");
}

extern "C" int piml_mlapm_sym_finalize_push_f32(const float *pos, const float *vel, const float *desired_speed,
                                                int ds_dim, const float *dest, int64_t N, int world, int rank,
                                                const piml_mlapm_params *prm, float dt, float radius,
                                                const float *inbox_local, const uint64_t *peer_pos_next_host,
                                                const uint64_t *peer_vel_next_host, uint8_t *arrived, void *workspace,
                                                int64_t workspace_bytes, void *stream) {
    PIML_REQUIRE(pos && vel && desired_speed && dest && prm && inbox_local && peer_pos_next_host && peer_vel_next_host,
                 "piml_mlapm_sym_finalize_push_f32: null pointer");
    PIML_REQUIRE(ds_dim == 1 || ds_dim == 2, "piml_mlapm_sym_finalize_push_f32: ds_dim=%d", ds_dim);
    SymShardPlan pl;
    int rc = sym_shard_plan(N, world, rank, workspace, workspace_bytes, &pl);
    if (rc) return rc;
    const MlConst k = mlapm_consts(prm);
    PeerPush push;
    push.world = world;
    for (int g = 0; g < ML_MAX_PEERS; ++g) {
        push.pos[g] = g < world ? reinterpret_cast<float2 *>(static_cast<uintptr_t>(peer_pos_next_host[g])) : nullptr;
        push.vel[g] = g < world ? reinterpret_cast<float2 *>(static_cast<uintptr_t>(peer_vel_next_host[g])) : nullptr;
        PIML_REQUIRE(g >= world || (push.pos[g] && push.vel[g]), "piml_mlapm_sym_finalize_push_f32: bad peer pointer");
    }
    const int64_t row0 = pl.I0 * MS_BLOCK;
    int64_t row1 = (pl.I0 + pl.nI) * MS_BLOCK;
    row1 = row1 > N ? N : row1;
    const int64_t nrows = row1 - row0;
    if (nrows <= 0) return PIML_OK;
    SymPartials symp{nullptr, 0.f, static_cast<int>(pl.nI * MS_BLOCK), reinterpret_cast<const float4 *>(inbox_local),
                     world, sym_inbox_stride(N, world)};
    const int threads = 256;
    mlapm_finalize2_kernel<<<static_cast<unsigned>((nrows * FZ_LPR + threads - 1) / threads), threads, 0,
                             static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2 *>(pos), reinterpret_cast<const float2 *>(vel), desired_speed, ds_dim,
        reinterpret_cast<const float2 *>(dest), static_cast<int>(row0), static_cast<int>(row1), pl.S, pl.partialR,
        prm->A, k.cos_t, k.sin_t, prm->version, prm->tau, dt, radius, nullptr, nullptr, arrived, push, symp);
    count_launch();
    return check_launch("mlapm_finalize2_kernel");
}

// ---- host-buffer entry of an agent-sharded step ------------------------------------------------------------------------
// Every rank uploads ONLY its own rows (1/G of the state over PCIe); this kernel stores them into every rank's
// current-state arrays over NVLink peer memory in ONE launch (the caller follows with a cross-rank barrier).
namespace piml {
struct ScatterPeers { int world; float2 *pos[ML_MAX_PEERS]; float2 *vel[ML_MAX_PEERS]; float2 *dest[ML_MAX_PEERS]; };

__global__ void scatter_rows_push_kernel(const float2 *__restrict__ pos_rows, const float2 *__restrict__ vel_rows,
                                         const float2 *__restrict__ dest_rows, int64_t row0, int64_t nrows,
                                         const __grid_constant__ ScatterPeers peers) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < nrows) {
        const float2 p = pos_rows[i], v = vel_rows[i];
        const float2 d = dest_rows ? dest_rows[i] : make_float2(0.f, 0.f);
        for (int g = 0; g < peers.world; ++g) {
            peers.pos[g][row0 + i] = p;
            peers.vel[g][row0 + i] = v;
            if (dest_rows) peers.dest[g][row0 + i] = d;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
}
}  // namespace piml

extern "C" int piml_scatter_rows_push_f32(const float *pos_rows, const float *vel_rows, const float *dest_rows,
                                          int64_t row0, int64_t nrows, int world, const uint64_t *peer_pos_host,
                                          const uint64_t *peer_vel_host, const uint64_t *peer_dest_host, void *stream) {
    PIML_REQUIRE(world >= 1 && world <= ML_MAX_PEERS && row0 >= 0 && nrows >= 0, "piml_scatter_rows_push_f32: bad shard");
    if (nrows == 0) return PIML_OK;
    PIML_REQUIRE(pos_rows && vel_rows && peer_pos_host && peer_vel_host && (!dest_rows || peer_dest_host),
                 "piml_scatter_rows_push_f32: null pointer");
    ScatterPeers peers;
    peers.world = world;
    for (int g = 0; g < ML_MAX_PEERS; ++g) {
        peers.pos[g] = g < world ? reinterpret_cast<float2 *>(static_cast<uintptr_t>(peer_pos_host[g])) : nullptr;
        peers.vel[g] = g < world ? reinterpret_cast<float2 *>(static_cast<uintptr_t>(peer_vel_host[g])) : nullptr;
        peers.dest[g] = (g < world && dest_rows) ? reinterpret_cast<float2 *>(static_cast<uintptr_t>(peer_dest_host[g]))
                                                 : nullptr;
        PIML_REQUIRE(g >= world || (peers.pos[g] && peers.vel[g] && (!dest_rows || peers.dest[g])),
                     "piml_scatter_rows_push_f32: bad peer pointer for rank %d", g);
    }
    const int threads = 256;
    scatter_rows_push_kernel<<<static_cast<unsigned>((nrows + threads - 1) / threads), threads, 0,
                               static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2 *>(pos_rows), reinterpret_cast<const float2 *>(vel_rows),
        reinterpret_cast<const float2 *>(dest_rows), row0, nrows, peers);
    count_launch();
    return check_launch("scatter_rows_push_kernel");
}

