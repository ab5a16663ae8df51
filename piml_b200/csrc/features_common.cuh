// features_common.cuh -- pieces shared by the all-pairs (features.cu) and the cell-list (features_cells.cu) variants of
// the neighbour-selection / relative-feature kernels: the exact-arithmetic field-of-view distance, the 64-bit
// (distance, index) keys, the register top-k list and the argument block.  Sharing them is what makes the two variants
// return the identical neighbour set: both evaluate the same predicate and order candidates by the same key.
#pragma once
#include <math.h>
#include <math_constants.h>

#include "common.cuh"

namespace piml {

constexpr int FEAT_THREADS = 128;
constexpr int FEAT_TILE = 1024;              // candidates per shared-memory tile (8 KB as float2)
constexpr uint64_t EMPTY_KEY = ~0ull;

// Ascending list of the KMAX smallest 64-bit keys seen so far, held in registers (fully unrolled).
template <int KMAX>
struct TopK {
    uint64_t key[KMAX];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < KMAX; ++i) key[i] = EMPTY_KEY;
    }
    __device__ __forceinline__ void insert(uint64_t c) {
        if (c >= key[KMAX - 1]) return;
#pragma unroll
        for (int i = KMAX - 1; i > 0; --i) {
            const uint64_t lo = key[i - 1];
            key[i] = (c < lo) ? lo : ((c < key[i]) ? c : key[i]);
        }
        key[0] = (c < key[0]) ? c : key[0];
    }
    __device__ __forceinline__ void pop_front() {
#pragma unroll
        for (int i = 0; i < KMAX - 1; ++i) key[i] = key[i + 1];
        key[KMAX - 1] = EMPTY_KEY;
    }
};

__device__ __forceinline__ uint64_t make_key(float dist, int idx) {
    return (static_cast<uint64_t>(__float_as_uint(dist)) << 32) | static_cast<uint32_t>(idx);
}
__device__ __forceinline__ float key_dist(uint64_t k) { return __uint_as_float(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ int key_idx(uint64_t k) { return static_cast<int>(static_cast<uint32_t>(k)); }

template <int G>
__device__ __forceinline__ uint64_t group_min(uint64_t v) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        const uint64_t o = __shfl_xor_sync(0xffffffffu, v, off);
        v = (o < v) ? o : v;
    }
    return v;
}

// Distance from (px,py) to (ox,oy) with the field-of-view gate applied: data.py:432-443.
// (hx,hy) is the heading already divided by max(||heading||, 1e-8) (cosine_similarity normalises each operand).
__device__ __forceinline__ float gated_distance(float rx, float ry, float hx, float hy, float cos_thr) {
    if (rx != rx) rx = CUDART_INF_F;                             // relative_pos[isnan] = inf   (:433)
    if (ry != ry) ry = CUDART_INF_F;
    float d = norm2_rn(rx, ry);                                  // torch.norm                  (:434)
    const float nr = fmaxf(d, 1e-8f);
    float c = __fadd_rn(__fmul_rn(__fdiv_rn(rx, nr), hx), __fmul_rn(__fdiv_rn(ry, nr), hy));   // (:439-440)
    if (c != c) c = -1.0f;                                       // view_field[isnan] = -1      (:441)
    if (c < cos_thr) d = CUDART_INF_F;                           // (:442-443)
    return d;
}

struct FeatArgs {
    const float *pos; float *vel; float *acc; const float *dest; const float *head; const float *obs;
    int64_t obs_frame_stride;    // floats between consecutive frames' obstacle arrays (0: shared)
    int obs_channel_T;           // if > 0: obstacle array index = frame / obs_channel_T (per-channel obstacles)
    int B, N, M, kp, ko;         // kp, ko already clamped to min(k, N|M)
    float cos_p, thr_p, pre2_p, cos_o, thr_o, pre2_o;
    float *ped_f; float *obs_f; float *dest_f;
    int64_t *ped_idx; float *ped_dist; int64_t *obs_idx; float *obs_dist;
    // optional rollout extras: self_f (B,N,7) = [dest_f, hist_v, acceleration, desired_speed]  (simulators.py:651)
    const float *hist_v; const float *desired_speed; float *self_f;
    // row range of the flattened (B*N) rows this call evaluates (agent-sharded ranks): outputs are indexed by
    // row - row0; row1 == 0 means all rows.  Cell-list evaluation only.
    int64_t row0, row1;
};

// slack so that sqrtf(d2) <= thr  =>  d2 <= pre2 for every fp32 d2 (prefilter must be a superset)
static inline float prefilter_sq(float thr) {
    if (!(thr < 1e18f)) return INFINITY;
    const double t = static_cast<double>(thr);
    return static_cast<float>(t * t * (1.0 + 1e-6)) + 1e-30f;
}

// Compact slot rows of the fused NN step (nn_step.cu): only the non-empty slots of each branch, 6 floats per row in
// arrival order (a row's message does not depend on its place), and for every (agent, slot) the row it went to
// (empty, zero-padded slots yield the network's f(0)).  counts[2] must be zero on entry.
struct CompactOut {
    float *rows_ped, *rows_obs;    // [counts[0]][6], [counts[1]][6]
    int *map_ped, *map_obs;        // (B*N, kp), (B*N, ko): written for the LIVE slots only, which are a prefix of an
                                   // agent's slots (ascending keys, empty slots last)
    uint16_t *live;                // (B*N): live pedestrian slots | live obstacle slots << 8
    int *counts;                   // nullptr: no compact output
};

// features_cells.cu
int relative_features_cells(const FeatArgs &a, int obs_frames, cudaStream_t st, const CompactOut *co = nullptr);
void cell_scratch_free();

}  // namespace piml
