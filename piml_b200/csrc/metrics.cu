// metrics.cu -- the per-frame evaluation metrics of BaseSimulator.test_multiple_rollouts (reference
// src/models/simulators.py:520-527 -> src/functions/metrics.py): for every frame t of a rollout, over the agents with
// mask[t] == 1,
//   mae  sum ||p - q||_2                                            mae_with_time_mask      metrics.py:29-42
//   ot   entropic optimal-transport cost between the two point sets  ot_with_time_mask       :45-67, SinkhornDistance :108-199
//   mmd  multi-bandwidth Gaussian-kernel maximum mean discrepancy    mmd_with_time_mask      :70-91, MaximumMeanDiscrepancy :207-273
// The reference loops over the T = 750 frames in Python and solves one small Sinkhorn problem (<= 100 log-domain
// iterations with a data-dependent stop) and one Gram matrix per frame: 23 s per clip.  Here one
// CTA owns one frame: the frame's masked points sit in shared memory, cost / Gram entries are recomputed on the fly
// (2-d points), the Sinkhorn loop stops per frame exactly like the reference's, and all frames run concurrently.
#include <math_constants.h>

#include "common.cuh"

namespace piml {

constexpr int MT_THREADS = 128;
constexpr int MT_MAXN = PIML_METRICS_MAX_AGENTS;  // masked agents per frame held in shared memory

struct MetricArgs {
    const float2 *p, *q; const uint8_t *mask; int T, N;
    float eps; int max_iter; float kernel_mul; int kernel_num;
    float *mae, *ot, *mmd; int *count;
};

__device__ __forceinline__ float block_sum(float v, float *red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int off = MT_THREADS / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    const float s = red[0];
    __syncthreads();
    return s;
}

__device__ __forceinline__ float cost2(float2 a, float2 b) {               // sum(|x - y| ** 2, -1)   (metrics.py:194-199)
    const float dx = fabsf(__fsub_rn(a.x, b.x)), dy = fabsf(__fsub_rn(a.y, b.y));
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__global__ void __launch_bounds__(MT_THREADS) metrics_frames_kernel(const __grid_constant__ MetricArgs a) {
    __shared__ float2 xs[MT_MAXN], ys[MT_MAXN];
    __shared__ float u[MT_MAXN], v[MT_MAXN];
    __shared__ float red[MT_THREADS];
    __shared__ int n_sh;
    const int t = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {                                                // compact the frame's masked agents, in slot order
        int n = 0;
        for (int j = 0; j < a.N; ++j)
            if (a.mask[static_cast<int64_t>(t) * a.N + j] == 1) {
                if (n < MT_MAXN) { xs[n] = a.p[static_cast<int64_t>(t) * a.N + j]; ys[n] = a.q[static_cast<int64_t>(t) * a.N + j]; }
                ++n;
            }
        n_sh = n;
    }
    __syncthreads();
    const int n = n_sh;
    if (tid == 0) a.count[t] = n;
    // ---- mae: sum over the masked agents of ||p - q||_2
    {
        float s = 0.f;                                             // straight from global memory: any number of agents
        for (int j = tid; j < a.N; j += MT_THREADS) {
            const int64_t e = static_cast<int64_t>(t) * a.N + j;
            if (a.mask[e] == 1) s += norm2_rn(__fsub_rn(a.p[e].x, a.q[e].x), __fsub_rn(a.p[e].y, a.q[e].y));
        }
        s = block_sum(s, red);
        if (tid == 0) a.mae[t] = s;
    }
    // the reference skips frames with fewer than 2 points; frames with more than MT_MAXN masked agents do not fit the
    // shared-memory point arrays: OT / MMD are NaN there and the host adapter raises (count[t] tells it)
    if (n <= 1 || n > MT_MAXN) {
        if (tid == 0) {
            if (a.ot) a.ot[t] = CUDART_NAN_F;
            if (a.mmd) a.mmd[t] = CUDART_NAN_F;
        }
        return;
    }
    // ---- ot: log-domain Sinkhorn (metrics.py:131-184), mu = nu = 1/n, u = v = 0, stop when sum |u - u_prev| < 0.1
    if (a.ot) {
        const float lmu = logf(__fadd_rn(1.0f / static_cast<float>(n), 1e-8f));
        for (int i = tid; i < n; i += MT_THREADS) { u[i] = 0.f; v[i] = 0.f; }
        __syncthreads();
        for (int it = 0; it < a.max_iter; ++it) {
            float errp = 0.f;
            for (int i = tid; i < n; i += MT_THREADS) {            // u = eps (log mu - logsumexp_j M_ij) + u
                const float ui = u[i];
                float mx = -CUDART_INF_F;
                for (int j = 0; j < n; ++j) mx = fmaxf(mx, __fdiv_rn(__fadd_rn(__fadd_rn(-cost2(xs[i], ys[j]), ui), v[j]), a.eps));
                float s = 0.f;
                for (int j = 0; j < n; ++j) s += expf(__fsub_rn(__fdiv_rn(__fadd_rn(__fadd_rn(-cost2(xs[i], ys[j]), ui), v[j]), a.eps), mx));
                const float un = __fadd_rn(__fmul_rn(a.eps, __fsub_rn(lmu, __fadd_rn(logf(s), mx))), ui);
                errp += fabsf(__fsub_rn(un, ui));
                u[i] = un;
            }
            __syncthreads();
            for (int j = tid; j < n; j += MT_THREADS) {            // v = eps (log nu - logsumexp_i M_ij) + v
                const float vj = v[j];
                float mx = -CUDART_INF_F;
                for (int i = 0; i < n; ++i) mx = fmaxf(mx, __fdiv_rn(__fadd_rn(__fadd_rn(-cost2(xs[i], ys[j]), u[i]), vj), a.eps));
                float s = 0.f;
                for (int i = 0; i < n; ++i) s += expf(__fsub_rn(__fdiv_rn(__fadd_rn(__fadd_rn(-cost2(xs[i], ys[j]), u[i]), vj), a.eps), mx));
                v[j] = __fadd_rn(__fmul_rn(a.eps, __fsub_rn(lmu, __fadd_rn(logf(s), mx))), vj);
            }
            const float err = block_sum(errp, red);                // also orders the v writes before the next reads
            if (err < 1e-1f) break;                                // uniform: every thread sees the same sum
        }
        float c = 0.f;                                             // cost = sum exp(M) * C   (:176-178)
        for (int i = tid; i < n; i += MT_THREADS)
            for (int j = 0; j < n; ++j) {
                const float cij = cost2(xs[i], ys[j]);
                c += __fmul_rn(expf(__fdiv_rn(__fadd_rn(__fadd_rn(-cij, u[i]), v[j]), a.eps)), cij);
            }
        c = block_sum(c, red);
        if (tid == 0) a.ot[t] = c;
    }
    // ---- mmd (metrics.py:213-273): total = [source; target], 2n points
    if (a.mmd) {
        const int ns = 2 * n;
        auto pt = [&](int k) { return k < n ? xs[k] : ys[k - n]; };
        auto l2 = [&](float2 c0, float2 c1) {                       // ((total0 - total1) ** 2).sum(2)
            const float dx = __fsub_rn(c0.x, c1.x), dy = __fsub_rn(c0.y, c1.y);
            return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        };
        float s = 0.f;
        for (int r = tid; r < ns; r += MT_THREADS) {
            const float2 pr = pt(r);
            for (int c = 0; c < ns; ++c) s += l2(pt(c), pr);
        }
        s = block_sum(s, red);
        float bw = __fdiv_rn(s, static_cast<float>(ns * ns - ns));              // :236
        bw = __fdiv_rn(bw, powf(a.kernel_mul, static_cast<float>(a.kernel_num / 2)));   // :237
        const float fn = static_cast<float>(n);
        const float dxx = fn * fn, dxy = -fn * fn;                              // n == m here (same mask on both sides)
        float acc_x = 0.f, acc_y = 0.f;                                         // (XX + XY).sum(), (YX + YY).sum()   (:271)
        for (int r = tid; r < ns; r += MT_THREADS) {
            const float2 pr = pt(r);
            float same = 0.f, cross = 0.f;
            for (int c = 0; c < ns; ++c) {
                const float d = l2(pt(c), pr);
                float k = 0.f;
                float mul = 1.0f;
                for (int q = 0; q < a.kernel_num; ++q) {                        // sum_i exp(-L2 / (bw * mul^i))   (:238-245)
                    k += expf(__fdiv_rn(-d, __fmul_rn(bw, mul)));
                    mul *= a.kernel_mul;
                }
                if ((c < n) == (r < n)) same += __fdiv_rn(k, dxx); else cross += __fdiv_rn(k, dxy);
            }
            if (r < n) acc_x += same + cross; else acc_y += cross + same;
        }
        const float sx = block_sum(acc_x, red);
        const float sy = block_sum(acc_y, red);
        if (tid == 0) a.mmd[t] = sx + sy;
    }
}

}  // namespace piml

using namespace piml;

extern "C" int piml_metrics_frames_f32(const float *p, const float *q, const uint8_t *mask, int T, int N, float eps,
                                       int max_iter, float kernel_mul, int kernel_num, float *out_mae, float *out_ot,
                                       float *out_mmd, int *out_count, void *stream) {
    if (T == 0) return PIML_OK;
    PIML_REQUIRE(p && q && mask && out_mae && out_count, "piml_metrics_frames_f32: null pointer");
    PIML_REQUIRE(T >= 0 && N >= 0, "piml_metrics_frames_f32: negative dimension");
    PIML_REQUIRE(!out_ot || (eps > 0.f && max_iter >= 1), "piml_metrics_frames_f32: eps must be > 0, max_iter >= 1");
    PIML_REQUIRE(!out_mmd || (kernel_num >= 1 && kernel_num <= 16 && kernel_mul > 0.f),
                 "piml_metrics_frames_f32: bad kernel_num / kernel_mul");
    if (T == 0) return PIML_OK;
    MetricArgs a{reinterpret_cast<const float2 *>(p), reinterpret_cast<const float2 *>(q), mask, T, N, eps, max_iter,
                 kernel_mul, kernel_num, out_mae, out_ot, out_mmd, out_count};
    metrics_frames_kernel<<<T, MT_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a);
    count_launch();
    return check_launch("metrics_frames_kernel");
}
