// losses.cu -- the rollout losses of BaseSimulator.test_multiple_rollouts_for_training (reference
// src/models/simulators.py:172-249, called at :795-824 with reduction 'sum') fused into one pass over the (C,T,N,2)
// trajectories, and their backward:
//   mse        sum decay_t (pred - labels)^2                                     multiple_rollout_mse_loss        :172-193
//   collision  sum w_cn decay_t (P pred - P labels)^2,  P x = x - (x . ni) ni,   multiple_rollout_collision_loss  :195-227
//              ni = (labels[T-1] - labels[0]) / (|.| + 1e-6),                    ..._collision_avoidance_loss     :229-249
//              w_cn = [sum_t collisions[c,t,n] > 0] (* abnormal_mask[n])
//   hard       the same with the hard-collision counts.
// The reference evaluates each of them with ~10 eager elementwise launches over (C,T,N,2) temporaries and a reduction.
// One thread per (c, n) walks the T steps; block sums are combined in a fixed order (deterministic).
#include "common.cuh"

namespace piml {

constexpr int LOSS_MAX_T = 256;
constexpr int LOSS_THREADS = 128;

struct LossArgs {
    const float *pred; const float *labels; int64_t label_stride;
    int C, T, N; int reverse;
    const float *coll; const float *hard; const float *abnormal;
    float decay[LOSS_MAX_T];
};

struct LossRow {                                   // what both passes need for one (c, n)
    float nix, niy, w1, w2;
};

__device__ __forceinline__ LossRow loss_row(const LossArgs &a, int c, int n) {
    LossRow r;
    const int64_t base = (static_cast<int64_t>(c) * a.T) * a.N + n;
    const float *l0 = a.labels + base * a.label_stride;
    const float *l1 = a.labels + (base + static_cast<int64_t>(a.T - 1) * a.N) * a.label_stride;
    const float nx = __fsub_rn(l1[0], l0[0]), ny = __fsub_rn(l1[1], l0[1]);      // labels[:, -1] - labels[:, 0]   (:242)
    const float nn = __fadd_rn(norm2_rn(nx, ny), 1e-6f);                         // :243-244
    r.nix = __fdiv_rn(nx, nn); r.niy = __fdiv_rn(ny, nn);                        // :245
    float s1 = 0.f, s2 = 0.f;
    for (int t = 0; t < a.T; ++t) {                                              // torch.sum(collisions, dim=1)   (:209)
        const int64_t e = base + static_cast<int64_t>(t) * a.N;
        if (a.coll) s1 += a.coll[e];
        if (a.hard) s2 += a.hard[e];
    }
    const float ab = a.abnormal ? a.abnormal[n] : 1.0f;                          // :222-224
    r.w1 = (a.coll && s1 > 0.f) ? ab : 0.f;                                      // collisions[collisions > 0] = 1 (:210)
    r.w2 = (a.hard && s2 > 0.f) ? ab : 0.f;
    return r;
}

// e = P pred - P labels for one (c,t,n)
__device__ __forceinline__ float2 loss_residual(const LossRow &r, float px, float py, float lx, float ly) {
    const float dp = __fadd_rn(__fmul_rn(px, r.nix), __fmul_rn(py, r.niy));      // torch.sum(pred * ni, dim=-1)   (:247)
    const float dl = __fadd_rn(__fmul_rn(lx, r.nix), __fmul_rn(ly, r.niy));
    const float ppx = __fsub_rn(px, __fmul_rn(dp, r.nix)), ppy = __fsub_rn(py, __fmul_rn(dp, r.niy));
    const float plx = __fsub_rn(lx, __fmul_rn(dl, r.nix)), ply = __fsub_rn(ly, __fmul_rn(dl, r.niy));
    return make_float2(__fsub_rn(ppx, plx), __fsub_rn(ppy, ply));
}

__device__ __forceinline__ float loss_decay(const LossArgs &a, int t) { return a.decay[t]; }

__global__ void __launch_bounds__(LOSS_THREADS) rollout_losses_kernel(const __grid_constant__ LossArgs a,
                                                                      float *__restrict__ partial) {
    __shared__ float red[3][LOSS_THREADS];
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    float m = 0.f, c1 = 0.f, c2 = 0.f;
    if (i < static_cast<int64_t>(a.C) * a.N) {
        const int c = static_cast<int>(i / a.N), n = static_cast<int>(i % a.N);
        const bool need_row = a.coll || a.hard;
        LossRow r{0.f, 0.f, 0.f, 0.f};
        if (need_row) r = loss_row(a, c, n);
        for (int t = 0; t < a.T; ++t) {
            const int64_t e = (static_cast<int64_t>(c) * a.T + t) * a.N + n;
            const float px = a.pred[e * 2], py = a.pred[e * 2 + 1];
            const float lx = a.labels[e * a.label_stride], ly = a.labels[e * a.label_stride + 1];
            const float dk = loss_decay(a, t);
            const float dx = __fsub_rn(px, lx), dy = __fsub_rn(py, ly);
            m += __fmul_rn(__fmul_rn(dx, dx), dk);                               // (pred-labels)^2 * decay        (:185-192)
            m += __fmul_rn(__fmul_rn(dy, dy), dk);
            if (r.w1 != 0.f || r.w2 != 0.f) {
                const float2 ev = loss_residual(r, px, py, lx, ly);
                const float f = __fmul_rn(__fmul_rn(ev.x, ev.x), dk) + __fmul_rn(__fmul_rn(ev.y, ev.y), dk);
                c1 += r.w1 * f;
                c2 += r.w2 * f;
            }
        }
    }
    red[0][threadIdx.x] = m; red[1][threadIdx.x] = c1; red[2][threadIdx.x] = c2;
    __syncthreads();
    for (int off = LOSS_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off)
            for (int q = 0; q < 3; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x < 3) partial[blockIdx.x * 3 + threadIdx.x] = red[threadIdx.x][0];
}

__global__ void rollout_losses_final_kernel(const float *__restrict__ partial, int nblocks, float *__restrict__ out) {
    if (threadIdx.x < 3) {
        double s = 0.0;                                              // fixed order; fp64 keeps the last bits honest
        for (int b = 0; b < nblocks; ++b) s += static_cast<double>(partial[b * 3 + threadIdx.x]);
        out[threadIdx.x] = static_cast<float>(s);
    }
}

// g_pred = g[0] * 2 decay (pred - labels) + (g[1] w1 + g[2] w2) * 2 decay * P^T (P pred - P labels)
__global__ void __launch_bounds__(LOSS_THREADS) rollout_losses_bwd_kernel(const __grid_constant__ LossArgs a,
                                                                          const float *__restrict__ g_out,
                                                                          float *__restrict__ g_pred) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<int64_t>(a.C) * a.N) return;
    const int c = static_cast<int>(i / a.N), n = static_cast<int>(i % a.N);
    const float g0 = g_out[0], g1 = g_out[1], g2 = g_out[2];
    LossRow r{0.f, 0.f, 0.f, 0.f};
    if (a.coll || a.hard) r = loss_row(a, c, n);
    const float gw = g1 * r.w1 + g2 * r.w2;
    for (int t = 0; t < a.T; ++t) {
        const int64_t e = (static_cast<int64_t>(c) * a.T + t) * a.N + n;
        const float px = a.pred[e * 2], py = a.pred[e * 2 + 1];
        const float lx = a.labels[e * a.label_stride], ly = a.labels[e * a.label_stride + 1];
        const float dk2 = 2.0f * loss_decay(a, t);
        float gx = g0 * dk2 * (px - lx), gy = g0 * dk2 * (py - ly);
        if (gw != 0.f) {
            const float2 ev = loss_residual(r, px, py, lx, ly);
            const float en = ev.x * r.nix + ev.y * r.niy;
            gx += gw * dk2 * (ev.x - en * r.nix);
            gy += gw * dk2 * (ev.y - en * r.niy);
        }
        g_pred[e * 2] = gx; g_pred[e * 2 + 1] = gy;
    }
}

static int fill_args(LossArgs *a, const float *pred, const float *labels, int64_t label_stride, int C, int T, int N,
                     float time_decay, int reverse, const float *coll, const float *hard, const float *abnormal) {
    PIML_REQUIRE(pred && labels, "piml_rollout_losses: null pointer");
    PIML_REQUIRE(C >= 0 && N >= 0 && T >= 1 && T <= LOSS_MAX_T, "piml_rollout_losses: bad dimensions C=%d T=%d N=%d "
                 "(T <= %d)", C, T, N, LOSS_MAX_T);
    PIML_REQUIRE(label_stride >= 2, "piml_rollout_losses: label_stride must be >= 2");
    a->pred = pred; a->labels = labels; a->label_stride = label_stride; a->C = C; a->T = T; a->N = N;
    a->reverse = reverse; a->coll = coll; a->hard = hard; a->abnormal = abnormal;
    // decay = torch.tensor([time_decay ** (T - t - 1) ...]) (or ** t when reversed): python floats -> fp32   (:186-190)
    for (int t = 0; t < T; ++t)
        a->decay[t] = static_cast<float>(pow(static_cast<double>(time_decay), reverse ? t : T - t - 1));
    return PIML_OK;
}

}  // namespace piml

using namespace piml;

extern "C" int64_t piml_rollout_losses_workspace_floats(int C, int N) {
    const int64_t rows = static_cast<int64_t>(C) * N;
    return ((rows + LOSS_THREADS - 1) / LOSS_THREADS) * 3 + 4;
}

extern "C" int piml_rollout_losses_f32(const float *pred, const float *labels, int64_t label_stride, int C, int T,
                                       int N, float time_decay, int reverse, const float *collisions,
                                       const float *hard_collisions, const float *abnormal_mask, float *out,
                                       float *workspace, void *stream) {
    LossArgs a;
    int rc = fill_args(&a, pred, labels, label_stride, C, T, N, time_decay, reverse, collisions, hard_collisions,
                       abnormal_mask);
    if (rc) return rc;
    PIML_REQUIRE(out && workspace, "piml_rollout_losses_f32: null output / workspace");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t rows = static_cast<int64_t>(C) * N;
    const int nblocks = static_cast<int>((rows + LOSS_THREADS - 1) / LOSS_THREADS);
    if (nblocks == 0) {
        PIML_CUDA(cudaMemsetAsync(out, 0, 3 * sizeof(float), st));
        return PIML_OK;
    }
    rollout_losses_kernel<<<nblocks, LOSS_THREADS, 0, st>>>(a, workspace);
    count_launch();
    rc = check_launch("rollout_losses_kernel");
    if (rc) return rc;
    rollout_losses_final_kernel<<<1, 32, 0, st>>>(workspace, nblocks, out);
    count_launch();
    return check_launch("rollout_losses_final_kernel");
}

extern "C" int piml_rollout_losses_backward_f32(const float *pred, const float *labels, int64_t label_stride, int C,
                                                int T, int N, float time_decay, int reverse, const float *collisions,
                                                const float *hard_collisions, const float *abnormal_mask,
                                                const float *g_out, float *g_pred, void *stream) {
    LossArgs a;
    int rc = fill_args(&a, pred, labels, label_stride, C, T, N, time_decay, reverse, collisions, hard_collisions,
                       abnormal_mask);
    if (rc) return rc;
    PIML_REQUIRE(g_out && g_pred, "piml_rollout_losses_backward_f32: null gradient pointer");
    const int64_t rows = static_cast<int64_t>(C) * N;
    if (rows == 0) return PIML_OK;
    rollout_losses_bwd_kernel<<<static_cast<unsigned>((rows + LOSS_THREADS - 1) / LOSS_THREADS), LOSS_THREADS, 0,
                                static_cast<cudaStream_t>(stream)>>>(a, g_out, g_pred);
    count_launch();
    return check_launch("rollout_losses_bwd_kernel");
}

// ---- the two small scalar losses of the training rollout -----------------------------------------------------------
// l1_reg_loss (simulators.py:169-170, 'sum'): weight * sum |x|, and the collision-prediction head's
// F.binary_cross_entropy(pred, target, reduction='sum') with its accuracy count (:826-830).  Both reduce a few hundred
// thousand elements: ONE CTA, fixed summation order (deterministic), no workspace.
namespace piml {

constexpr int SL_THREADS = 1024;

__device__ __forceinline__ double block_sum_d(double v, double *red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int off = SL_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    const double s = red[0];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(SL_THREADS) l1_sum_kernel(const float *__restrict__ x, int64_t n, float weight,
                                                            float *__restrict__ out) {
    __shared__ double red[SL_THREADS];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += SL_THREADS) s += static_cast<double>(__fmul_rn(weight, fabsf(x[i])));
    s = block_sum_d(s, red);
    if (threadIdx.x == 0) out[0] = static_cast<float>(s);
}

__global__ void l1_sum_bwd_kernel(const float *__restrict__ x, int64_t n, float weight, const float *__restrict__ g,
                                  float *__restrict__ gx) {
    const float gw = __fmul_rn(g[0], weight);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float v = x[i];
        gx[i] = v > 0.f ? gw : (v < 0.f ? -gw : 0.f);              // d|x|/dx = sign(x), 0 at 0 like torch.abs
    }
}

// binary_cross_entropy: -(t max(log p, -100) + (1 - t) max(log(1 - p), -100)), ATen's clamp; out[0] = sum,
// out[1] = number of elements with round(p) == t (torch.round: half to even)
__global__ void __launch_bounds__(SL_THREADS) bce_sum_kernel(const float *__restrict__ p, const float *__restrict__ t,
                                                             int64_t n, float *__restrict__ out) {
    __shared__ double red[SL_THREADS];
    double s = 0.0, hit = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += SL_THREADS) {
        const float pi = p[i], ti = t[i];
        const float lp = fmaxf(logf(pi), -100.f), lq = fmaxf(log1pf(-pi), -100.f);
        s += static_cast<double>(-(__fadd_rn(__fmul_rn(ti, lp), __fmul_rn(__fsub_rn(1.f, ti), lq))));
        hit += (rintf(pi) == ti) ? 1.0 : 0.0;
    }
    s = block_sum_d(s, red);
    hit = block_sum_d(hit, red);
    if (threadIdx.x == 0) { out[0] = static_cast<float>(s); out[1] = static_cast<float>(hit); }
}

__global__ void bce_sum_bwd_kernel(const float *__restrict__ p, const float *__restrict__ t, int64_t n,
                                   const float *__restrict__ g, float *__restrict__ gp) {
    const float g0 = g[0];
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float pi = p[i];
        gp[i] = g0 * (pi - t[i]) / fmaxf((1.f - pi) * pi, 1e-12f);  // ATen binary_cross_entropy_backward
    }
}

}  // namespace piml

static unsigned grid_for(int64_t n) {
    const int64_t b = (n + 255) / 256;
    return static_cast<unsigned>(b < 148 * 8 ? (b > 0 ? b : 1) : 148 * 8);
}

extern "C" int piml_l1_sum_f32(const float *x, int64_t n, float weight, float *out, void *stream) {
    PIML_REQUIRE(out && (x || n == 0) && n >= 0, "piml_l1_sum_f32: bad argument");
    l1_sum_kernel<<<1, SL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, n, weight, out);
    count_launch();
    return check_launch("l1_sum_kernel");
}

extern "C" int piml_l1_sum_backward_f32(const float *x, int64_t n, float weight, const float *g_out, float *g_x,
                                        void *stream) {
    PIML_REQUIRE(g_out && (n == 0 || (x && g_x)) && n >= 0, "piml_l1_sum_backward_f32: bad argument");
    if (n == 0) return PIML_OK;
    l1_sum_bwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, weight, g_out, g_x);
    count_launch();
    return check_launch("l1_sum_bwd_kernel");
}

extern "C" int piml_bce_sum_f32(const float *pred, const float *target, int64_t n, float *out, void *stream) {
    PIML_REQUIRE(out && (n == 0 || (pred && target)) && n >= 0, "piml_bce_sum_f32: bad argument");
    bce_sum_kernel<<<1, SL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(pred, target, n, out);
    count_launch();
    return check_launch("bce_sum_kernel");
}

extern "C" int piml_bce_sum_backward_f32(const float *pred, const float *target, int64_t n, const float *g_out,
                                         float *g_pred, void *stream) {
    PIML_REQUIRE(g_out && (n == 0 || (pred && target && g_pred)) && n >= 0, "piml_bce_sum_backward_f32: bad argument");
    if (n == 0) return PIML_OK;
    bce_sum_bwd_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, target, n, g_out, g_pred);
    count_launch();
    return check_launch("bce_sum_bwd_kernel");
}

