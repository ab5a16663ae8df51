// features.cu -- neighbour selection + relative features for sm_100a.
//
// Replaces Pedestrians.get_heading_direction / get_relative_quantity / get_nearby_obj_in_sight /
// get_filtered_features / get_relative_features (reference src/data/data.py:351-512).  The reference materialises
// (T,N,N,2)x3 + (T,N,N,6) tensors and fully sorts every row; here one kernel streams candidate tiles through
// shared memory (TMA bulk copies, double buffered), keeps a running top-k per row in registers, merges the lanes
// that share a row with warp shuffles, and writes the k gathered 6-d features directly.
//
// Bit-exactness: the field-of-view predicate and the distances use the exact fp32 operation order of the
// reference's CPU kernels through explicit _rn intrinsics; ordering is (distance, index)
// ascending, which is what torch.sort produces for distinct finite distances (ties/inf are unspecified there).
#include <math.h>
#include <math_constants.h>

#include <atomic>

#include "features_common.cuh"

namespace piml {

// Scan `M` candidates (float2 array `cand`, one frame) for the rows owned by this CTA and build, per row, the
// ascending list of the best keys.  RADIUS: only candidates with distance <= thr are kept and a cheap squared
// distance prefilter skips the exact arithmetic for everything else.  All threads of the CTA must call this.
template <int KMAX, int G, bool RADIUS>
__device__ __forceinline__ void scan_candidates(TopK<KMAX> &best, const float2 *__restrict__ cand, int M, float px,
                                                float py, float hx, float hy, float cos_thr, float thr, float pre2,
                                                int g, float2 (*tile)[FEAT_TILE], uint64_t *bars,
                                                uint32_t &phase_bits) {
    best.init();
    const int ntiles = (M + FEAT_TILE - 1) / FEAT_TILE;
    if (ntiles == 0) return;
    bool waits[2];
    waits[0] = stage_float2(tile[0], cand, min(M, FEAT_TILE), &bars[0]);
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) {
            const int m1 = (t + 1) * FEAT_TILE;
            waits[buf ^ 1] = stage_float2(tile[buf ^ 1], cand + m1, min(M - m1, FEAT_TILE), &bars[buf ^ 1]);
        }
        if (waits[buf]) {
            mbar_wait(&bars[buf], (phase_bits >> buf) & 1u);
            phase_bits ^= (1u << buf);
        }
        const int m0 = t * FEAT_TILE;
        const int tn = min(M - m0, FEAT_TILE);
        const float2 *tl = tile[buf];
#pragma unroll 4
        for (int j = g; j < tn; j += G) {
            const float2 o = tl[j];
            const float rx = __fsub_rn(o.x, px);                 // relative = B - A            (:412)
            const float ry = __fsub_rn(o.y, py);
            bool cand_ok = true;
            if (RADIUS) cand_ok = __fmaf_rn(ry, ry, __fmul_rn(rx, rx)) <= pre2;
            if (cand_ok) {
                const float d = gated_distance(rx, ry, hx, hy, cos_thr);
                if (!RADIUS || d <= thr) best.insert(make_key(d, m0 + j));
            }
        }
        __syncthreads();        // tile[buf] is free for the copy issued at the top of iteration t+1
    }
}


// grid = (ceil(N / (FEAT_THREADS/G)), B).  G lanes cooperate on one row.
template <int KP, int KO, int G>
__global__ void __launch_bounds__(FEAT_THREADS) relative_features_kernel(FeatArgs a) {
    __shared__ __align__(16) float2 tile[2][FEAT_TILE];
    __shared__ __align__(8) uint64_t bars[2];
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase_bits = 0;

    constexpr int ROWS = FEAT_THREADS / G;
    const int b = blockIdx.y;
    const int g = threadIdx.x % G;
    const int n = blockIdx.x * ROWS + threadIdx.x / G;
    const bool live = n < a.N;
    const int nn = live ? n : a.N - 1;                          // dead lanes shadow the last row, never write
    const int64_t row = static_cast<int64_t>(b) * a.N + nn;

    const float2 p = reinterpret_cast<const float2 *>(a.pos)[row];
    float2 v = reinterpret_cast<const float2 *>(a.vel)[row];
    float2 ac = reinterpret_cast<const float2 *>(a.acc)[row];
    // acceleration[isnan] = 0 ; velocity[isnan] = 0, in place on the caller's tensors (data.py:483-484)
    if (live && g == 0) {
        if (v.x != v.x || v.y != v.y)
            reinterpret_cast<float2 *>(a.vel)[row] = make_float2(nan_to_zero(v.x), nan_to_zero(v.y));
        if (ac.x != ac.x || ac.y != ac.y)
            reinterpret_cast<float2 *>(a.acc)[row] = make_float2(nan_to_zero(ac.x), nan_to_zero(ac.y));
    }
    v = make_float2(nan_to_zero(v.x), nan_to_zero(v.y));
    ac = make_float2(nan_to_zero(ac.x), nan_to_zero(ac.y));

    float2 h;
    if (a.head) {
        h = reinterpret_cast<const float2 *>(a.head)[row];
    } else {                                                     // T == 1: heading = v / ||v||, 0 -> /0.1 (:391-394)
        float nv = norm2_rn(v.x, v.y);
        if (nv == 0.0f) nv = 0.1f;
        h = make_float2(__fdiv_rn(v.x, nv), __fdiv_rn(v.y, nv));
    }
    {                                                            // cosine_similarity's own normalisation of x2
        const float nh = fmaxf(norm2_rn(h.x, h.y), 1e-8f);
        h = make_float2(__fdiv_rn(h.x, nh), __fdiv_rn(h.y, nh));
    }

    // ---- pedestrian - pedestrian (data.py:489-494) ----
    {
        TopK<KP> best;
        const float2 *cand = reinterpret_cast<const float2 *>(a.pos) + static_cast<int64_t>(b) * a.N;
        scan_candidates<KP, G, true>(best, cand, a.N, p.x, p.y, h.x, h.y, a.cos_p, a.thr_p, a.pre2_p, g, tile,
                                     bars, phase_bits);
        const float2 *fv = reinterpret_cast<const float2 *>(a.vel) + static_cast<int64_t>(b) * a.N;
        const float2 *fa = reinterpret_cast<const float2 *>(a.acc) + static_cast<int64_t>(b) * a.N;
        for (int j = 0; j < a.kp; ++j) {
            const uint64_t w = group_min<G>(best.key[0]);
            if (w != EMPTY_KEY && best.key[0] == w) best.pop_front();
            if (live && g == (j % G)) {
                float2 f0 = make_float2(0.f, 0.f), f1 = f0, f2 = f0;
                if (w != EMPTY_KEY) {
                    const int m = key_idx(w);
                    const float2 pm = cand[m];
                    const float2 vm = fv[m];
                    const float2 am = fa[m];
                    f0 = make_float2(__fsub_rn(pm.x, p.x), __fsub_rn(pm.y, p.y));
                    f1 = make_float2(__fsub_rn(nan_to_zero(vm.x), v.x), __fsub_rn(nan_to_zero(vm.y), v.y));
                    f2 = make_float2(__fsub_rn(nan_to_zero(am.x), ac.x), __fsub_rn(nan_to_zero(am.y), ac.y));
                }
                float2 *out = reinterpret_cast<float2 *>(a.ped_f) + (row * a.kp + j) * 3;
                out[0] = f0; out[1] = f1; out[2] = f2;
                if (a.ped_idx) a.ped_idx[row * a.kp + j] = (w != EMPTY_KEY) ? key_idx(w) : -1;
                if (a.ped_dist) a.ped_dist[row * a.kp + j] = (w != EMPTY_KEY) ? key_dist(w) : CUDART_INF_F;
            }
        }
    }

    // ---- destination (data.py:496-497) ----
    if (live && g == 0) {
        const float2 d = reinterpret_cast<const float2 *>(a.dest)[row];
        const float2 df = make_float2(nan_to_zero(__fsub_rn(d.x, p.x)), nan_to_zero(__fsub_rn(d.y, p.y)));
        reinterpret_cast<float2 *>(a.dest_f)[row] = df;
        if (a.self_f) {
            const float2 hv = reinterpret_cast<const float2 *>(a.hist_v)[row];
            float *sf = a.self_f + row * 7;
            sf[0] = df.x; sf[1] = df.y; sf[2] = hv.x; sf[3] = hv.y; sf[4] = ac.x; sf[5] = ac.y;
            sf[6] = a.desired_speed[row];
        }
    }

    // ---- pedestrian - obstacle (data.py:499-510): obs = (o, 0, 0) ----
    if (a.M > 0) {
        TopK<KO> best;
        const int oframe = a.obs_channel_T > 0 ? b / a.obs_channel_T : b;
        const float2 *cand = reinterpret_cast<const float2 *>(a.obs + static_cast<int64_t>(oframe) * a.obs_frame_stride);
        scan_candidates<KO, G, true>(best, cand, a.M, p.x, p.y, h.x, h.y, a.cos_o, a.thr_o, a.pre2_o, g, tile, bars,
                                     phase_bits);
        for (int j = 0; j < a.ko; ++j) {
            const uint64_t w = group_min<G>(best.key[0]);
            if (w != EMPTY_KEY && best.key[0] == w) best.pop_front();
            if (live && g == (j % G)) {
                float2 f0 = make_float2(0.f, 0.f), f1 = f0, f2 = f0;
                if (w != EMPTY_KEY) {
                    const float2 om = cand[key_idx(w)];
                    f0 = make_float2(__fsub_rn(om.x, p.x), __fsub_rn(om.y, p.y));
                    f1 = make_float2(__fsub_rn(0.f, v.x), __fsub_rn(0.f, v.y));
                    f2 = make_float2(__fsub_rn(0.f, ac.x), __fsub_rn(0.f, ac.y));
                }
                float2 *out = reinterpret_cast<float2 *>(a.obs_f) + (row * a.ko + j) * 3;
                out[0] = f0; out[1] = f1; out[2] = f2;
                if (a.obs_idx) a.obs_idx[row * a.ko + j] = (w != EMPTY_KEY) ? key_idx(w) : -1;
                if (a.obs_dist) a.obs_dist[row * a.ko + j] = (w != EMPTY_KEY) ? key_dist(w) : CUDART_INF_F;
            }
        }
    }
}

// get_nearby_obj_in_sight proper: no radius, every object competes (inf-distance ones ordered by index).
struct SelArgs {
    const float *pos; const float *obj; int64_t obj_frame_stride; const float *head;
    int B, N, M, kk; float cos_thr; float *out_dist; int64_t *out_idx;
};

template <int KMAX, int G>
__global__ void __launch_bounds__(FEAT_THREADS) select_kernel(SelArgs a) {
    __shared__ __align__(16) float2 tile[2][FEAT_TILE];
    __shared__ __align__(8) uint64_t bars[2];
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase_bits = 0;
    constexpr int ROWS = FEAT_THREADS / G;
    const int b = blockIdx.y;
    const int g = threadIdx.x % G;
    const int n = blockIdx.x * ROWS + threadIdx.x / G;
    const bool live = n < a.N;
    const int64_t row = static_cast<int64_t>(b) * a.N + (live ? n : a.N - 1);
    const float2 p = reinterpret_cast<const float2 *>(a.pos)[row];
    float2 h = reinterpret_cast<const float2 *>(a.head)[row];
    const float nh = fmaxf(norm2_rn(h.x, h.y), 1e-8f);
    h = make_float2(__fdiv_rn(h.x, nh), __fdiv_rn(h.y, nh));
    TopK<KMAX> best;
    const float2 *cand = reinterpret_cast<const float2 *>(a.obj + static_cast<int64_t>(b) * a.obj_frame_stride);
    scan_candidates<KMAX, G, false>(best, cand, a.M, p.x, p.y, h.x, h.y, a.cos_thr, 0.f, 0.f, g, tile, bars,
                                    phase_bits);
    for (int j = 0; j < a.kk; ++j) {
        const uint64_t w = group_min<G>(best.key[0]);
        if (best.key[0] == w) best.pop_front();
        if (live && g == (j % G)) {
            a.out_dist[row * a.kk + j] = key_dist(w);
            a.out_idx[row * a.kk + j] = key_idx(w);
        }
    }
}

// get_heading_direction for T > 1 (data.py:362-395): one thread per (channel, pedestrian) walks the time axis.
__global__ void heading_kernel(const float2 *__restrict__ vel, int C, int T, int N, float2 *__restrict__ head) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * N) return;
    const int c = i / N, n = i % N;
    const float2 *v = vel + static_cast<int64_t>(c) * T * N + n;
    float2 *h = head + static_cast<int64_t>(c) * T * N + n;
    float2 tmp = make_float2(0.f, 0.f);
    for (int t = T - 1; t >= 0; --t) {                           // nearest LATER non-zero velocity (:366-370)
        float2 x = v[static_cast<int64_t>(t) * N];
        x = make_float2(nan_to_zero(x.x), nan_to_zero(x.y));
        if (norm2_rn(x.x, x.y) == 0.0f) x = tmp; else tmp = x;
        h[static_cast<int64_t>(t) * N] = x;
    }
    for (int t = 0; t < T; ++t) {                                // else nearest EARLIER one (:371-375), normalise
        float2 x = h[static_cast<int64_t>(t) * N];
        if (norm2_rn(x.x, x.y) == 0.0f) x = tmp; else tmp = x;
        float nv = norm2_rn(x.x, x.y);
        if (nv == 0.0f) nv = 0.1f;
        h[static_cast<int64_t>(t) * N] = make_float2(__fdiv_rn(x.x, nv), __fdiv_rn(x.y, nv));
    }
}

// calculate_collision_label (data.py:515-535)
__global__ void collision_label_kernel(const float *__restrict__ ped_f, int64_t S, float *__restrict__ out) {
    const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float2 rp = reinterpret_cast<const float2 *>(ped_f)[s * 3];
    const float2 rv = reinterpret_cast<const float2 *>(ped_f)[s * 3 + 1];
    float hit = 0.f;
#pragma unroll
    for (int q = 0; q < 10; ++q) {
        const float tq = __fmul_rn(static_cast<float>(q), 0.1f);
        const float d = norm2_rn(__fadd_rn(rp.x, __fmul_rn(rv.x, tq)), __fadd_rn(rp.y, __fmul_rn(rv.y, tq)));
        if (d < 0.5f && d != 0.f) hit = 1.f;
    }
    out[s] = hit;
}


// ---- backward of the relative features (differentiable rollout, simulators.py:772-778) -----------------------------
struct FeatBwdArgs {
    const float2 *pos; const float2 *dest; const int64_t *ped_idx; const int64_t *obs_idx; int64_t rows; int N, kp, ko;
    const float2 *g_ped_f; const float2 *g_obs_f; const float2 *g_dest_f;
    float *g_pos; float *g_vel; float *g_acc; float2 *g_dest;
};

__device__ __forceinline__ void atomic_add2(float *base, int64_t i, float2 g) {
    atomicAdd(base + 2 * i, g.x);
    atomicAdd(base + 2 * i + 1, g.y);
}

// one thread per (frame, agent): its own (negative) share is summed in registers, the neighbours' (positive) shares
// are scattered.  ped_f = (p_m - p_n, v_m - v_n, a_m - a_n), obs_f = (o - p_n, -v_n, -a_n), dest_f = dest - p_n.
__global__ void relative_features_bwd_kernel(FeatBwdArgs a) {
    const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= a.rows) return;
    const int64_t frame0 = (row / a.N) * a.N;
    float2 gp = make_float2(0.f, 0.f), gv = gp, ga = gp;
    for (int j = 0; j < a.kp; ++j) {
        const int64_t m = a.ped_idx[row * a.kp + j];
        if (m < 0) continue;                                      // zero-padded slot: constant (data.py:459-462)
        const float2 *g = a.g_ped_f + (row * a.kp + j) * 3;
        const float2 g0 = g[0], g1 = g[1], g2 = g[2];
        gp.x -= g0.x; gp.y -= g0.y; gv.x -= g1.x; gv.y -= g1.y; ga.x -= g2.x; ga.y -= g2.y;
        atomic_add2(a.g_pos, frame0 + m, g0);
        atomic_add2(a.g_vel, frame0 + m, g1);
        atomic_add2(a.g_acc, frame0 + m, g2);
    }
    for (int j = 0; j < a.ko; ++j) {
        if (a.obs_idx[row * a.ko + j] < 0) continue;
        const float2 *g = a.g_obs_f + (row * a.ko + j) * 3;
        const float2 g0 = g[0], g1 = g[1], g2 = g[2];
        gp.x -= g0.x; gp.y -= g0.y; gv.x -= g1.x; gv.y -= g1.y; ga.x -= g2.x; ga.y -= g2.y;
    }
    float2 gd = make_float2(0.f, 0.f);
    if (a.g_dest_f) {                                             // dest_features[isnan] = 0 cuts the gradient (:497)
        const float2 p = a.pos[row], d = a.dest[row], g = a.g_dest_f[row];
        const float dx = d.x - p.x, dy = d.y - p.y;
        if (!(dx != dx)) { gd.x = g.x; gp.x -= g.x; }
        if (!(dy != dy)) { gd.y = g.y; gp.y -= g.y; }
    }
    a.g_dest[row] = gd;
    atomic_add2(a.g_pos, row, gp);
    atomic_add2(a.g_vel, row, gv);
    atomic_add2(a.g_acc, row, ga);
}

// ---- Pedestrians.collision_detection (data.py:538-601) ---------------------------------------------------------
struct CollArgs {
    const float2 *pos; const float2 *real; int C, T, N; float thr; int mode; float *full; float *rowsum;
};

__device__ __forceinline__ bool touching(const float2 *frame, int n, int m, float thr) {
    const float2 pn = frame[n], pm = frame[m];
    const float d = norm2_rn(__fsub_rn(pm.x, pn.x), __fsub_rn(pm.y, pn.y));      // NaN compares false -> 0 (:550)
    return d < thr;
}

// one thread per (channel, n, m); walks the time axis twice (friends, then output).
__global__ void collision_detection_kernel(CollArgs a) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t NN = static_cast<int64_t>(a.N) * a.N;
    if (i >= NN * a.C) return;
    const int c = static_cast<int>(i / NN);
    const int n = static_cast<int>((i % NN) / a.N), m = static_cast<int>(i % a.N);
    const float2 *base = a.pos + static_cast<int64_t>(c) * a.T * a.N;
    bool friends;
    if (a.mode == 3) {
        // friends: more than 25 touching frames in real_position (diagonal included there, :571-579) or in position
        int cnt = 0;
        if (a.real) { for (int t = 0; t < a.T; ++t) cnt += touching(a.real + static_cast<int64_t>(t) * a.N, n, m, a.thr); }
        else if (n != m) { for (int t = 0; t < a.T; ++t) cnt += touching(base + static_cast<int64_t>(t) * a.N, n, m, a.thr); }
        friends = cnt > 25;
    } else {
        // training: pairs touching in any of the first 4 frames are friends (:588-593)
        friends = false;
        if (n != m)
            for (int t = 0; t < min(4, a.T); ++t) friends |= touching(base + static_cast<int64_t>(t) * a.N, n, m, a.thr);
    }
    for (int t = 0; t < a.T; ++t) {
        const bool hit = n != m && !friends && touching(base + static_cast<int64_t>(t) * a.N, n, m, a.thr);
        const int64_t fr = static_cast<int64_t>(c) * a.T + t;
        if (a.full) a.full[(fr * a.N + n) * a.N + m] = hit ? 1.f : 0.f;
        if (a.rowsum && hit) atomicAdd(a.rowsum + fr * a.N + n, 1.f);       // integer-valued: order independent
    }
}

static std::atomic<int> g_feature_algo{0};     // 0: automatic, 1: all pairs, 2: cell list
constexpr int CELLS_MIN_AGENTS = 4096;

static int pick_group(int64_t B, int N) {
    // enough CTAs to fill 148 SMs several times over with one thread per row?  otherwise spread a row over lanes
    const int64_t target = 4LL * sm_count();
    if (B * ((N + FEAT_THREADS - 1) / FEAT_THREADS) >= target) return 1;
    if (B * ((N + FEAT_THREADS / 8 - 1) / (FEAT_THREADS / 8)) >= target) return 8;
    return 32;
}


template <int KP, int KO>
static void launch_features(const FeatArgs &a, int G, cudaStream_t st) {
    if (G == 1) {
        dim3 grid((a.N + FEAT_THREADS - 1) / FEAT_THREADS, a.B);
        relative_features_kernel<KP, KO, 1><<<grid, FEAT_THREADS, 0, st>>>(a);
    } else if (G == 8) {
        dim3 grid((a.N + FEAT_THREADS / 8 - 1) / (FEAT_THREADS / 8), a.B);
        relative_features_kernel<KP, KO, 8><<<grid, FEAT_THREADS, 0, st>>>(a);
    } else {
        dim3 grid((a.N + FEAT_THREADS / 32 - 1) / (FEAT_THREADS / 32), a.B);
        relative_features_kernel<KP, KO, 32><<<grid, FEAT_THREADS, 0, st>>>(a);
    }
}

template <int KMAX>
static void launch_select(const SelArgs &a, int G, cudaStream_t st) {
    if (G == 1) {
        dim3 grid((a.N + FEAT_THREADS - 1) / FEAT_THREADS, a.B);
        select_kernel<KMAX, 1><<<grid, FEAT_THREADS, 0, st>>>(a);
    } else if (G == 8) {
        dim3 grid((a.N + FEAT_THREADS / 8 - 1) / (FEAT_THREADS / 8), a.B);
        select_kernel<KMAX, 8><<<grid, FEAT_THREADS, 0, st>>>(a);
    } else {
        dim3 grid((a.N + FEAT_THREADS / 32 - 1) / (FEAT_THREADS / 32), a.B);
        select_kernel<KMAX, 32><<<grid, FEAT_THREADS, 0, st>>>(a);
    }
}

}  // namespace piml

using namespace piml;

extern "C" int piml_heading_f32(const float *vel, int C, int T, int N, float *head, void *stream) {
    PIML_REQUIRE(vel && head, "piml_heading_f32: null pointer");
    PIML_REQUIRE(C >= 0 && T >= 0 && N >= 0, "piml_heading_f32: negative dimension");
    if (static_cast<int64_t>(C) * T * N == 0) return PIML_OK;
    const int threads = 128;
    heading_kernel<<<(C * N + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2 *>(vel), C, T, N, reinterpret_cast<float2 *>(head));
    count_launch();
    return check_launch("heading_kernel");
}

// TimeIndexedPedData.make_dataset, data.py:797-806: per pedestrian, the mean speed over the first `skip` frames
// starting at its first frame with non-zero velocity (frame 0 when it never moves).  One thread per pedestrian; the
// (T,N,2) reads of a warp are coalesced across pedestrians.
__global__ void desired_speed_kernel(const float2 *__restrict__ vel, int T, int N, int skip, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int start = 0;
    for (int j = 0; j < T; ++j) {
        const float2 v = vel[static_cast<size_t>(j) * N + i];
        if (piml::norm2_rn(v.x, v.y) > 0.0f) { start = j; break; }
    }
    const int end = min(start + skip, T);
    double s = 0.0;                              // torch.mean's fp32 cascade order is an ATen detail: <= 1 ulp apart
    for (int j = start; j < end; ++j) {
        const float2 v = vel[static_cast<size_t>(j) * N + i];
        s += static_cast<double>(piml::norm2_rn(v.x, v.y));
    }
    out[i] = static_cast<float>(s / static_cast<double>(end - start));      // empty slice -> NaN like torch.mean
}

extern "C" int piml_desired_speed_f32(const float *vel, int T, int N, int skip_frames, float *out, void *stream) {
    PIML_REQUIRE(vel && out, "piml_desired_speed_f32: null pointer");
    PIML_REQUIRE(T >= 0 && N >= 0 && skip_frames >= 0, "piml_desired_speed_f32: negative dimension");
    if (N == 0) return PIML_OK;
    const int threads = 128;
    desired_speed_kernel<<<(N + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2 *>(vel), T, N, skip_frames, out);
    count_launch();
    return check_launch("desired_speed_kernel");
}

// Pedestrians.get_relative_quantity, data.py:398-414: out[b,n,m,:] = B[b,m,:] - A[b,n,:].  The hot path never
// materialises this tensor (it is fused into the selection / feature kernels); this entry serves the reference's
// polar / symbolic-regression callers that ask for it explicitly.
__global__ void relative_quantity_kernel(const float *__restrict__ A, const float *__restrict__ B, int64_t total, int N,
                                         int M, int d, float *__restrict__ out) {
    for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % d);
        const int64_t r = e / d;
        const int m = static_cast<int>(r % M);
        const int64_t bn = r / M;
        const int64_t b = bn / N;
        out[e] = __fsub_rn(B[(b * M + m) * d + c], A[bn * d + c]);
    }
}

extern "C" int piml_relative_quantity_f32(const float *A, const float *B, int64_t frames, int N, int M, int d,
                                          float *out, void *stream) {
    PIML_REQUIRE(A && B && out, "piml_relative_quantity_f32: null pointer");
    PIML_REQUIRE(frames >= 0 && N >= 0 && M >= 0 && d >= 1, "piml_relative_quantity_f32: bad dimension");
    const int64_t total = frames * N * M * d;
    if (total == 0) return PIML_OK;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    relative_quantity_kernel<<<static_cast<unsigned>(blocks < 148 * 32 ? blocks : 148 * 32), threads, 0,
                               static_cast<cudaStream_t>(stream)>>>(A, B, total, N, M, d, out);
    count_launch();
    return check_launch("relative_quantity_kernel");
}

// Pedestrians.get_filtered_features, data.py:449-464: gather the k selected columns of (rows, M, d) and zero every
// slot whose distance exceeds the threshold (NaN distances keep their slot, like `nearby_dist > thr`).
__global__ void filtered_features_kernel(const float *__restrict__ feat, const int64_t *__restrict__ idx,
                                         const float *__restrict__ dist, int64_t total, int M, int k, int d, float thr,
                                         float *__restrict__ out) {
    for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(e % d);
        const int64_t slot = e / d;
        const int64_t row = slot / k;
        out[e] = (dist[slot] > thr) ? 0.0f : feat[(row * M + idx[slot]) * d + c];
    }
}

extern "C" int piml_filtered_features_f32(const float *features, const int64_t *idx, const float *dist, int64_t rows,
                                          int M, int k, int d, float dist_threshold, float *out, void *stream) {
    PIML_REQUIRE(features && idx && dist && out, "piml_filtered_features_f32: null pointer");
    PIML_REQUIRE(rows >= 0 && M >= 1 && k >= 0 && d >= 1, "piml_filtered_features_f32: bad dimension");
    const int64_t total = rows * k * d;
    if (total == 0) return PIML_OK;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    filtered_features_kernel<<<static_cast<unsigned>(blocks < 148 * 32 ? blocks : 148 * 32), threads, 0,
                               static_cast<cudaStream_t>(stream)>>>(features, idx, dist, total, M, k, d, dist_threshold,
                                                                    out);
    count_launch();
    return check_launch("filtered_features_kernel");
}

extern "C" int piml_select_neighbors_f32(const float *pos, const float *obj, int64_t obj_frame_stride,
                                         const float *head, int B, int N, int M, int k, float cos_thr,
                                         float *out_dist, int64_t *out_idx, void *stream) {
    PIML_REQUIRE(pos && obj && head && out_dist && out_idx, "piml_select_neighbors_f32: null pointer");
    PIML_REQUIRE(B >= 0 && N >= 0 && M >= 0 && k >= 0, "piml_select_neighbors_f32: negative dimension");
    const int kk = k < M ? k : M;
    PIML_REQUIRE(kk <= 32, "piml_select_neighbors_f32: k=%d > 32 is not supported", k);
    PIML_REQUIRE(B <= 65535, "piml_select_neighbors_f32: more than 65535 frames per call");
    if (static_cast<int64_t>(B) * N == 0 || kk == 0) return PIML_OK;
    SelArgs a{pos, obj, obj_frame_stride, head, B, N, M, kk, cos_thr, out_dist, out_idx};
    const int G = pick_group(B, N);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (kk <= 8) launch_select<8>(a, G, st);
    else if (kk <= 16) launch_select<16>(a, G, st);
    else launch_select<32>(a, G, st);
    count_launch();
    return check_launch("select_kernel");
}

static int relative_features_impl(const float *pos, float *vel, float *acc, const float *dest,
                                          const float *head, const float *obs, int obs_per_channel, int C, int T,
                                          int N, int M, int kp, float cos_thr_ped, float dist_thr_ped, int ko,
                                          float cos_thr_obs, float dist_thr_obs, float *ped_f, float *obs_f,
                                          float *dest_f, int64_t *ped_idx, float *ped_dist, int64_t *obs_idx,
                                          float *obs_dist, const float *hist_v, const float *desired_speed,
                                          float *self_f, void *stream, int64_t row0 = 0, int64_t row1 = 0) {
    PIML_REQUIRE(C >= 0 && T >= 0 && N >= 0 && M >= 0 && kp >= 0 && ko >= 0,
                 "piml_relative_features_f32: negative dimension");
    if (static_cast<int64_t>(C) * T * N == 0) return PIML_OK;         // empty batch: pointers may be null
    PIML_REQUIRE(pos && vel && acc && dest && ped_f && dest_f, "piml_relative_features_f32: null pointer");
    PIML_REQUIRE(M == 0 || (obs && obs_f), "piml_relative_features_f32: obstacles given but obs/obs_f is null");
    PIML_REQUIRE(head || T == 1, "piml_relative_features_f32: head may only be NULL when T == 1 (got T=%d)", T);
    const int kpp = kp < N ? kp : N;
    const int kop = ko < M ? ko : M;
    PIML_REQUIRE(kpp <= 32 && kop <= 32, "piml_relative_features_f32: topk (%d,%d) > 32 is not supported", kp, ko);
    const int64_t B = static_cast<int64_t>(C) * T;
    PIML_REQUIRE(B <= 65535, "piml_relative_features_f32: more than 65535 frames per call (split the batch)");
    if (B * N == 0) return PIML_OK;
    FeatArgs a;
    a.pos = pos; a.vel = vel; a.acc = acc; a.dest = dest; a.head = head; a.obs = obs;
    a.obs_frame_stride = obs_per_channel ? static_cast<int64_t>(M) * 2 : 0;
    a.obs_channel_T = obs_per_channel ? T : 0;
    a.B = static_cast<int>(B); a.N = N; a.M = M; a.kp = kpp; a.ko = kop;
    a.cos_p = cos_thr_ped; a.thr_p = dist_thr_ped; a.pre2_p = prefilter_sq(dist_thr_ped);
    a.cos_o = cos_thr_obs; a.thr_o = dist_thr_obs; a.pre2_o = prefilter_sq(dist_thr_obs);
    a.ped_f = ped_f; a.obs_f = obs_f; a.dest_f = dest_f;
    a.ped_idx = ped_idx; a.ped_dist = ped_dist; a.obs_idx = obs_idx; a.obs_dist = obs_dist;
    a.hist_v = hist_v; a.desired_speed = desired_speed; a.self_f = self_f;
    a.row0 = row0; a.row1 = row1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // large crowds: uniform-grid cell list (identical neighbour set, see features_cells.cu); small scenes: all pairs
    const int algo = g_feature_algo.load(std::memory_order_relaxed);
    const float thr_max = fmaxf(dist_thr_ped, M > 0 ? dist_thr_obs : 0.f);
    const bool cells_ok = thr_max > 0.f && thr_max < 1e18f;
    if (row1 > 0) {
        PIML_REQUIRE(0 <= row0 && row0 < row1 && row1 <= B * N, "piml_state_features_rows_f32: bad row range");
        PIML_REQUIRE(cells_ok, "piml_state_features_rows_f32: row ranges need a finite distance threshold (cell list)");
        return relative_features_cells(a, obs_per_channel ? C : 1, st);
    }
    if (cells_ok && (algo == 2 || (algo == 0 && N >= CELLS_MIN_AGENTS)))
        return relative_features_cells(a, obs_per_channel ? C : 1, st);
    const int G = pick_group(B, N);
    if (kpp <= 8 && kop <= 16) launch_features<8, 16>(a, G, st);
    else if (kpp <= 16 && kop <= 16) launch_features<16, 16>(a, G, st);
    else launch_features<32, 32>(a, G, st);
    count_launch();
    return check_launch("relative_features_kernel");
}

extern "C" int piml_relative_features_f32(const float *pos, float *vel, float *acc, const float *dest,
                                          const float *head, const float *obs, int obs_per_channel, int C, int T,
                                          int N, int M, int kp, float cos_thr_ped, float dist_thr_ped, int ko,
                                          float cos_thr_obs, float dist_thr_obs, float *ped_f, float *obs_f,
                                          float *dest_f, int64_t *ped_idx, float *ped_dist, int64_t *obs_idx,
                                          float *obs_dist, void *stream) {
    return relative_features_impl(pos, vel, acc, dest, head, obs, obs_per_channel, C, T, N, M, kp, cos_thr_ped,
                                  dist_thr_ped, ko, cos_thr_obs, dist_thr_obs, ped_f, obs_f, dest_f, ped_idx,
                                  ped_dist, obs_idx, obs_dist, nullptr, nullptr, nullptr, stream);
}

extern "C" int piml_state_features_f32(const float *pos, float *vel, float *acc, const float *dest, const float *obs,
                                       int obs_per_scene, int S, int N, int M, int kp, float cos_thr_ped,
                                       float dist_thr_ped, int ko, float cos_thr_obs, float dist_thr_obs,
                                       const float *hist_v, const float *desired_speed, float *ped_f, float *obs_f,
                                       float *self_f, float *dest_f, void *stream) {
    PIML_REQUIRE(hist_v && desired_speed && self_f, "piml_state_features_f32: null pointer");
    return relative_features_impl(pos, vel, acc, dest, nullptr, obs, obs_per_scene, S, 1, N, M, kp, cos_thr_ped,
                                  dist_thr_ped, ko, cos_thr_obs, dist_thr_obs, ped_f, obs_f, dest_f, nullptr,
                                  nullptr, nullptr, nullptr, hist_v, desired_speed, self_f, stream);
}

extern "C" int piml_state_features_rows_f32(const float *pos, float *vel, float *acc, const float *dest,
                                            const float *obs, int N, int M, int64_t row0, int64_t row1, int kp,
                                            float cos_thr_ped, float dist_thr_ped, int ko, float cos_thr_obs,
                                            float dist_thr_obs, const float *hist_v, const float *desired_speed,
                                            float *ped_f, float *obs_f, float *self_f, float *dest_f, void *stream) {
    PIML_REQUIRE(hist_v && desired_speed && self_f, "piml_state_features_rows_f32: null pointer");
    PIML_REQUIRE(row1 > row0, "piml_state_features_rows_f32: empty row range");
    return relative_features_impl(pos, vel, acc, dest, nullptr, obs, 0, 1, 1, N, M, kp, cos_thr_ped, dist_thr_ped, ko,
                                  cos_thr_obs, dist_thr_obs, ped_f, obs_f, dest_f, nullptr, nullptr, nullptr, nullptr,
                                  hist_v, desired_speed, self_f, stream, row0, row1);
}

extern "C" int piml_collision_label_f32(const float *ped_f, int64_t S, float *out, void *stream) {
    PIML_REQUIRE(ped_f && out, "piml_collision_label_f32: null pointer");
    PIML_REQUIRE(S >= 0, "piml_collision_label_f32: negative size");
    if (S == 0) return PIML_OK;
    const int threads = 256;
    collision_label_kernel<<<static_cast<unsigned>((S + threads - 1) / threads), threads, 0,
                             static_cast<cudaStream_t>(stream)>>>(ped_f, S, out);
    count_launch();
    return check_launch("collision_label_kernel");
}

extern "C" int piml_relative_features_backward_f32(const float *pos, const float *dest, const int64_t *ped_idx,
                                                   const int64_t *obs_idx, int B, int N, int kp, int ko,
                                                   const float *g_ped_f, const float *g_obs_f, const float *g_dest_f,
                                                   float *g_pos, float *g_vel, float *g_acc, float *g_dest,
                                                   void *stream) {
    PIML_REQUIRE(pos && dest && g_pos && g_vel && g_acc && g_dest, "piml_relative_features_backward_f32: null pointer");
    PIML_REQUIRE(B >= 0 && N >= 0 && kp >= 0 && ko >= 0, "piml_relative_features_backward_f32: negative dimension");
    PIML_REQUIRE(kp == 0 || (ped_idx && g_ped_f), "piml_relative_features_backward_f32: ped_idx / g_ped_f is null");
    PIML_REQUIRE(ko == 0 || (obs_idx && g_obs_f), "piml_relative_features_backward_f32: obs_idx / g_obs_f is null");
    const int64_t rows = static_cast<int64_t>(B) * N;
    if (rows == 0) return PIML_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PIML_CUDA(cudaMemsetAsync(g_pos, 0, sizeof(float) * 2 * rows, st));
    PIML_CUDA(cudaMemsetAsync(g_vel, 0, sizeof(float) * 2 * rows, st));
    PIML_CUDA(cudaMemsetAsync(g_acc, 0, sizeof(float) * 2 * rows, st));
    FeatBwdArgs a;
    a.pos = reinterpret_cast<const float2 *>(pos); a.dest = reinterpret_cast<const float2 *>(dest);
    a.ped_idx = ped_idx; a.obs_idx = obs_idx; a.rows = rows; a.N = N; a.kp = kp; a.ko = ko;
    a.g_ped_f = reinterpret_cast<const float2 *>(g_ped_f); a.g_obs_f = reinterpret_cast<const float2 *>(g_obs_f);
    a.g_dest_f = reinterpret_cast<const float2 *>(g_dest_f);
    a.g_pos = g_pos; a.g_vel = g_vel; a.g_acc = g_acc; a.g_dest = reinterpret_cast<float2 *>(g_dest);
    const int threads = 128;
    relative_features_bwd_kernel<<<static_cast<unsigned>((rows + threads - 1) / threads), threads, 0, st>>>(a);
    count_launch();
    return check_launch("relative_features_bwd_kernel");
}

extern "C" int piml_collision_detection_f32(const float *position, const float *real_position, int C, int T, int N,
                                            float threshold, int mode, float *out_full, float *out_rowsum,
                                            void *stream) {
    PIML_REQUIRE(position && (out_full || out_rowsum), "piml_collision_detection_f32: null pointer");
    PIML_REQUIRE(C >= 0 && T >= 0 && N >= 0, "piml_collision_detection_f32: negative dimension");
    PIML_REQUIRE(mode == 3 || mode == 4, "piml_collision_detection_f32: mode must be 3 or 4 (input rank)");
    PIML_REQUIRE(mode == 4 || C == 1, "piml_collision_detection_f32: mode 3 takes one (T,N,2) clip (C = 1)");
    PIML_REQUIRE(!real_position || mode == 3, "piml_collision_detection_f32: real_position only with mode 3");
    const int64_t tot = static_cast<int64_t>(C) * N * N;
    if (tot == 0 || T == 0) return PIML_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (out_rowsum) PIML_CUDA(cudaMemsetAsync(out_rowsum, 0, sizeof(float) * C * T * N, st));
    CollArgs a{reinterpret_cast<const float2 *>(position), reinterpret_cast<const float2 *>(real_position), C, T, N,
               threshold, mode, out_full, out_rowsum};
    const int threads = 256;
    collision_detection_kernel<<<static_cast<unsigned>((tot + threads - 1) / threads), threads, 0, st>>>(a);
    count_launch();
    return check_launch("collision_detection_kernel");
}

extern "C" int piml_set_feature_algorithm(int algo) {
    PIML_REQUIRE(algo >= 0 && algo <= 2, "piml_set_feature_algorithm: 0 = automatic, 1 = all pairs, 2 = cell list");
    g_feature_algo.store(algo, std::memory_order_relaxed);
    return PIML_OK;
}

namespace piml { void tc_scratch_free(); void nn_scratch_free(); }   // mlp_tc.cu, nn_step.cu

extern "C" int piml_free_workspace(void) {
    cell_scratch_free();
    piml::tc_scratch_free();
    piml::nn_scratch_free();
    return PIML_OK;
}
