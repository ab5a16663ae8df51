// features_cells.cu -- uniform-grid (cell list) variant of the neighbour-selection + relative-feature kernel for large
// crowds (sm_100a).  Same outputs, bit for bit, as relative_features_kernel in features.cu, which scans all N (+M)
// candidates per agent like the reference does (src/data/data.py:416-447 sorts every row of the N x N distance matrix).
//
// Why the neighbour set is identical: get_filtered_features zeroes every slot farther than dist_threshold
// (data.py:459-462), so only candidates within the radius can ever appear in the output; a grid of cells at least as
// wide as the radius guarantees that all of them lie in the 3 x 3 block of cells around the agent, and both variants
// evaluate the same exact-arithmetic predicate (gated_distance) and keep the k smallest (distance, index) keys.
//
// Pipeline per call (no host synchronisation, all on the caller's stream):
//   1. cell_count2_kernel  : cell of every point -> spatial-hash bucket; atomic count per bucket, rank of the point;
//                            agents without a position get their (all-empty) outputs right here
//   2. scan (2 kernels)    : exclusive prefix sum of the bucket counts; the second kernel zeroes the counters it has
//                            read, so the next call needs no memset (3 kernels + memset beyond 4096 scan blocks)
//   3. cell_scatter2_kernel: records {x, y, flat row | obstacle index, packed cell} sorted by bucket (counting sort)
//   4. features_sorted_kernel: G = 4 lanes per agent, agents taken in SORTED (bucket) order, so the lanes of a warp walk
//      the same 9 buckets (coherent trip counts, broadcast loads); each lane keeps the k best of every 4th candidate,
//      the lists are merged with shuffles and the lanes share the slot outputs.  Hash collisions and repeated buckets
//      are filtered by comparing the record's packed cell with the cell being visited.  Optionally it also emits the
//      COMPACT slot rows (non-empty slots only, + a slot -> row map) the fused NN step feeds to the tensor cores
//      (nn_step.cu); agent-sharded ranks of the fused step pass the list of their own agents (own_list_kernel).
//      Row-range calls with DENSE outputs (the three-call sharded route) keep the one-thread-per-agent
//      features_cells_kernel.
// Buckets are a hash of the integer cell coordinates (no bounding box needed => no device->host read of extents);
// cell coordinates are computed in fp64 so that rounding can never move a candidate two cells away.
#include <mutex>

#include "features_common.cuh"

namespace piml {

constexpr int CELL_THREADS = 128;
constexpr int SCAN_PER_BLOCK = 2048;      // elements per block of the prefix sum (256 threads x 8)

struct HashGrid {
    int H;                 // buckets per frame (power of two)
    int frames, n;         // frames x n points
    const int *start;      // frames * H + 1 exclusive prefix sums
    const float4 *rec;     // frames * n records sorted by bucket: {x, y, as_float(index), as_float(packed cell)}
};

__device__ __forceinline__ int cell_coord(float x, double inv_cs) {
    double t = floor(static_cast<double>(x) * inv_cs);
    t = fmin(fmax(t, -32768.0), 32767.0);                          // monotone clamp keeps "at most one cell apart"
    return static_cast<int>(t);
}
__device__ __forceinline__ uint32_t pack_cell(int cx, int cy) {
    return (static_cast<uint32_t>(cx) & 0xffffu) | (static_cast<uint32_t>(cy) << 16);
}
__device__ __forceinline__ uint32_t bucket_of(int cx, int cy, int H) {
    const uint32_t h = static_cast<uint32_t>(cx) * 73856093u ^ static_cast<uint32_t>(cy) * 19349663u;
    return (h ^ (h >> 15)) & static_cast<uint32_t>(H - 1);
}

// exclusive scan, phase 1: per-block totals
__global__ void __launch_bounds__(256) scan_block_sums_kernel(const int *__restrict__ in, int64_t n,
                                                              int *__restrict__ block_sums) {
    __shared__ int red[256];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_PER_BLOCK;
    int s = 0;
    for (int e = threadIdx.x; e < SCAN_PER_BLOCK; e += 256)
        if (base + e < n) s += in[base + e];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = red[0];
}

// phase 2: one block turns the block totals into exclusive offsets (sequential over chunks of 1024)
__global__ void __launch_bounds__(1024) scan_offsets_kernel(int *__restrict__ block_sums, int nblocks) {
    __shared__ int buf[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < nblocks; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const int v = i < nblocks ? block_sums[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const int t = threadIdx.x >= off ? buf[threadIdx.x - off] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblocks) block_sums[i] = carry + buf[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += buf[1023];
        __syncthreads();
    }
}

// phase 3: exclusive scan inside each block + block offset; out has n + 1 entries (out[n] = grand total)
__global__ void __launch_bounds__(256) scan_apply_kernel(const int *__restrict__ in, int64_t n,
                                                         const int *__restrict__ block_offs, int *__restrict__ out) {
    __shared__ int part[256];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_PER_BLOCK + threadIdx.x * 8;
    int v[8], s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { v[q] = (base + q < n) ? in[base + q] : 0; s += v[q]; }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const int t = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    int run = block_offs[blockIdx.x] + part[threadIdx.x] - s;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (base + q < n) out[base + q] = run;
        run += v[q];
    }
    if (base <= n - 1 && n - 1 < base + 8) out[n] = run;           // the thread owning the last element
}

// phases 2 + 3 in one launch for up to SCAN_FUSED_BLOCKS blocks: every block adds up the totals of the blocks before
// it by itself (a few hundred ints from L2) instead of waiting for a one-block offsets kernel.  It also ZEROES the
// counters it has read (and the own-list counter behind them), so the next call's count kernel needs no memset:
// the grid build is 4 launches (count, block sums, this, scatter).
constexpr int SCAN_FUSED_BLOCKS = 4096;
__global__ void __launch_bounds__(256) scan_apply_fused_kernel(int *__restrict__ in, int64_t n,
                                                               const int *__restrict__ block_sums,
                                                               int *__restrict__ out) {
    __shared__ int part[256];
    __shared__ int red[8];
    int pre = 0;
    for (int b = threadIdx.x; b < static_cast<int>(blockIdx.x); b += 256) pre += block_sums[b];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = pre;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_PER_BLOCK + threadIdx.x * 8;
    int v[8], s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        v[q] = (base + q < n) ? in[base + q] : 0;
        s += v[q];
        if (base + q < n) in[base + q] = 0;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const int t = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
#pragma unroll
    for (int w = 0; w < 8; ++w) run += red[w];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (base + q < n) out[base + q] = run;
        run += v[q];
    }
    if (base <= n - 1 && n - 1 < base + 8) { out[n] = run; in[n] = 0; in[n + 1] = 0; }   // the thread owning the last element
}

// Both point sets of a feature call (agents and obstacles) in ONE counting-sort chain: the cell arrays are
// concatenated (agents' buckets first), so one count / scan / scatter sequence builds both grids.
// set 0: points [0, total0), n0 per frame, H0 buckets per frame, cells from 0;
// set 1: the following total1 points, cells from cells0 on.  The exclusive scan runs over the concatenation, so the
// obstacle grid's start values already include the number of agent records and both grids index the SAME record array.
struct TwoSets {
    const float2 *pts0, *pts1; int64_t total0, total1; int n0, n1, H0, H1; int64_t cells0;
};

// The agents without a position (NaN: absent from the frame) are in no bucket and need no search: everything the
// feature call owes them -- all-zero slots (a live-slot count of 0 in the compact form), dest_f / self_f, the in-place
// NaN -> 0 of their velocity and acceleration (data.py:483-484) -- is written right here by the count kernel, which
// sees every agent anyway; the sorted-order feature kernel then takes the present agents only.  (A list of the absent
// agents for that kernel's tail was tried first: 4096 mostly empty scenes append 400k entries per step through ONE
// counter, 25 us of same-address atomics even when aggregated per warp.)
struct AbsentOut {
    bool enabled;                  // false: the one-thread-per-agent kernel handles absent rows itself (row-range calls)
    FeatArgs a;                    // output pointers, dims, row range (row1 > 0: only rows [row0, row1) are written)
    uint16_t *live;                // compact form: live-slot counts, or nullptr
};

__global__ void cell_count2_kernel(const TwoSets t, double inv_cs, int *__restrict__ counts, int *__restrict__ rank,
                                   const AbsentOut ab) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= t.total0 + t.total1) return;
    const bool second = i >= t.total0;
    const int64_t j = second ? i - t.total0 : i;
    const float2 p = (second ? t.pts1 : t.pts0)[j];
    if (p.x != p.x || p.y != p.y) {
        rank[i] = -1;
        const FeatArgs &a = ab.a;
        if (second || !ab.enabled || (a.row1 > 0 && (j < a.row0 || j >= a.row1))) return;
        const int64_t row = j;
        float2 v = reinterpret_cast<const float2 *>(a.vel)[row], ac = reinterpret_cast<const float2 *>(a.acc)[row];
        if (v.x != v.x || v.y != v.y)
            reinterpret_cast<float2 *>(a.vel)[row] = make_float2(nan_to_zero(v.x), nan_to_zero(v.y));
        if (ac.x != ac.x || ac.y != ac.y)
            reinterpret_cast<float2 *>(a.acc)[row] = make_float2(nan_to_zero(ac.x), nan_to_zero(ac.y));
        ac = make_float2(nan_to_zero(ac.x), nan_to_zero(ac.y));
        const float2 z = make_float2(0.f, 0.f);
        if (a.ped_f)
            for (int q = 0; q < a.kp; ++q) {
                float2 *out = reinterpret_cast<float2 *>(a.ped_f) + (row * a.kp + q) * 3;
                out[0] = z; out[1] = z; out[2] = z;
                if (a.ped_idx) a.ped_idx[row * a.kp + q] = -1;
                if (a.ped_dist) a.ped_dist[row * a.kp + q] = CUDART_INF_F;
            }
        if (a.obs_f && a.M > 0)
            for (int q = 0; q < a.ko; ++q) {
                float2 *out = reinterpret_cast<float2 *>(a.obs_f) + (row * a.ko + q) * 3;
                out[0] = z; out[1] = z; out[2] = z;
                if (a.obs_idx) a.obs_idx[row * a.ko + q] = -1;
                if (a.obs_dist) a.obs_dist[row * a.ko + q] = CUDART_INF_F;
            }
        if (a.dest_f) {
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
        if (ab.live) ab.live[row] = 0;
        return;
    }
    const int n = second ? t.n1 : t.n0, H = second ? t.H1 : t.H0;
    const int64_t frame = j / n;
    const uint32_t b = bucket_of(cell_coord(p.x, inv_cs), cell_coord(p.y, inv_cs), H);
    rank[i] = atomicAdd(&counts[(second ? t.cells0 : 0) + frame * H + b], 1);
}

__global__ void cell_scatter2_kernel(const TwoSets t, double inv_cs, const int *__restrict__ start,
                                     const int *__restrict__ rank, float4 *__restrict__ rec) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= t.total0 + t.total1) return;
    const int r = rank[i];
    if (r < 0) return;
    const bool second = i >= t.total0;
    const int64_t j = second ? i - t.total0 : i;
    const float2 p = (second ? t.pts1 : t.pts0)[j];
    const int n = second ? t.n1 : t.n0, H = second ? t.H1 : t.H0;
    const int64_t frame = j / n;
    const int cx = cell_coord(p.x, inv_cs), cy = cell_coord(p.y, inv_cs);
    const uint32_t b = bucket_of(cx, cy, H);
    // agents carry their FLAT row (frame * n + index: the sorted-order kernel finds the agent from the record),
    // obstacles their index inside the frame
    rec[start[(second ? t.cells0 : 0) + frame * H + b] + r] =
        make_float4(p.x, p.y, __int_as_float(static_cast<int>(second ? j - frame * n : j)), __uint_as_float(pack_cell(cx, cy)));
}

// Keep the k best gated candidates of the 3 x 3 cell block around (cx, cy).
template <int KMAX>
__device__ __forceinline__ void scan_cells(TopK<KMAX> &best, const HashGrid &g, int64_t gframe, int cx, int cy, float px,
                                           float py, float hx, float hy, float cos_thr, float thr, float pre2,
                                           int idx_base) {
    best.init();
    const int *st = g.start + gframe * g.H;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int ccx = cx + dx, ccy = cy + dy;
            const uint32_t want = pack_cell(ccx, ccy);
            const uint32_t b = bucket_of(ccx, ccy, g.H);
            const int e1 = st[b + 1];
            for (int e = st[b]; e < e1; ++e) {
                const float4 r = g.rec[e];
                if (__float_as_uint(r.w) != want) continue;       // hash collision or a wrapped neighbour cell
                const float rx = __fsub_rn(r.x, px), ry = __fsub_rn(r.y, py);
                if (!(__fmaf_rn(ry, ry, __fmul_rn(rx, rx)) <= pre2)) continue;
                const float d = gated_distance(rx, ry, hx, hy, cos_thr);
                if (d <= thr) best.insert(make_key(d, __float_as_int(r.z) - idx_base));
            }
        }
}

template <int KP, int KO>
__global__ void __launch_bounds__(CELL_THREADS) features_cells_kernel(FeatArgs a, HashGrid gp, HashGrid go,
                                                                      double inv_cs) {
    const int64_t all_rows = static_cast<int64_t>(a.B) * a.N;
    const int64_t first = a.row1 > 0 ? a.row0 : 0, last = a.row1 > 0 ? a.row1 : all_rows;
    if (a.row1 > 0) {
        // a row range leaves the other rows' velocities / accelerations to other ranks, but every rank keeps the whole
        // state: apply the in-place NaN -> 0 of data.py:483-484 to ALL rows so that the replicas stay identical
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < all_rows;
             i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
            if (i >= first && i < last) continue;
            const float2 vi = reinterpret_cast<const float2 *>(a.vel)[i], ai = reinterpret_cast<const float2 *>(a.acc)[i];
            if (vi.x != vi.x || vi.y != vi.y)
                reinterpret_cast<float2 *>(a.vel)[i] = make_float2(nan_to_zero(vi.x), nan_to_zero(vi.y));
            if (ai.x != ai.x || ai.y != ai.y)
                reinterpret_cast<float2 *>(a.acc)[i] = make_float2(nan_to_zero(ai.x), nan_to_zero(ai.y));
        }
    }
    const int64_t row = first + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= last) return;
    const int64_t orow = row - first;                               // where this row's outputs go
    const int b = static_cast<int>(row / a.N);
    const float2 p = reinterpret_cast<const float2 *>(a.pos)[row];
    float2 v = reinterpret_cast<const float2 *>(a.vel)[row];
    float2 ac = reinterpret_cast<const float2 *>(a.acc)[row];
    if (v.x != v.x || v.y != v.y)                                  // in-place NaN -> 0 (data.py:483-484)
        reinterpret_cast<float2 *>(a.vel)[row] = make_float2(nan_to_zero(v.x), nan_to_zero(v.y));
    if (ac.x != ac.x || ac.y != ac.y)
        reinterpret_cast<float2 *>(a.acc)[row] = make_float2(nan_to_zero(ac.x), nan_to_zero(ac.y));
    v = make_float2(nan_to_zero(v.x), nan_to_zero(v.y));
    ac = make_float2(nan_to_zero(ac.x), nan_to_zero(ac.y));
    float2 h;
    if (a.head) {
        h = reinterpret_cast<const float2 *>(a.head)[row];
    } else {
        float nv = norm2_rn(v.x, v.y);
        if (nv == 0.0f) nv = 0.1f;
        h = make_float2(__fdiv_rn(v.x, nv), __fdiv_rn(v.y, nv));
    }
    {
        const float nh = fmaxf(norm2_rn(h.x, h.y), 1e-8f);
        h = make_float2(__fdiv_rn(h.x, nh), __fdiv_rn(h.y, nh));
    }
    const bool present = !(p.x != p.x || p.y != p.y);
    const int cx = present ? cell_coord(p.x, inv_cs) : 0, cy = present ? cell_coord(p.y, inv_cs) : 0;

    // ---- pedestrian - pedestrian ----
    {
        TopK<KP> best;
        best.init();
        if (present) scan_cells<KP>(best, gp, b, cx, cy, p.x, p.y, h.x, h.y, a.cos_p, a.thr_p, a.pre2_p, b * a.N);
        const float2 *fp = reinterpret_cast<const float2 *>(a.pos) + static_cast<int64_t>(b) * a.N;
        const float2 *fv = reinterpret_cast<const float2 *>(a.vel) + static_cast<int64_t>(b) * a.N;
        const float2 *fa = reinterpret_cast<const float2 *>(a.acc) + static_cast<int64_t>(b) * a.N;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            if (j >= a.kp) break;
            const uint64_t w = best.key[j];
            float2 f0 = make_float2(0.f, 0.f), f1 = f0, f2 = f0;
            if (w != EMPTY_KEY) {
                const int m = key_idx(w);
                const float2 pm = fp[m], vm = fv[m], am = fa[m];
                f0 = make_float2(__fsub_rn(pm.x, p.x), __fsub_rn(pm.y, p.y));
                f1 = make_float2(__fsub_rn(nan_to_zero(vm.x), v.x), __fsub_rn(nan_to_zero(vm.y), v.y));
                f2 = make_float2(__fsub_rn(nan_to_zero(am.x), ac.x), __fsub_rn(nan_to_zero(am.y), ac.y));
            }
            float2 *out = reinterpret_cast<float2 *>(a.ped_f) + (orow * a.kp + j) * 3;
            out[0] = f0; out[1] = f1; out[2] = f2;
            if (a.ped_idx) a.ped_idx[orow * a.kp + j] = (w != EMPTY_KEY) ? key_idx(w) : -1;
            if (a.ped_dist) a.ped_dist[orow * a.kp + j] = (w != EMPTY_KEY) ? key_dist(w) : CUDART_INF_F;
        }
    }
    // ---- destination ----
    {
        const float2 d = reinterpret_cast<const float2 *>(a.dest)[row];
        const float2 df = make_float2(nan_to_zero(__fsub_rn(d.x, p.x)), nan_to_zero(__fsub_rn(d.y, p.y)));
        reinterpret_cast<float2 *>(a.dest_f)[orow] = df;
        if (a.self_f) {
            const float2 hv = reinterpret_cast<const float2 *>(a.hist_v)[row];
            float *sf = a.self_f + orow * 7;
            sf[0] = df.x; sf[1] = df.y; sf[2] = hv.x; sf[3] = hv.y; sf[4] = ac.x; sf[5] = ac.y;
            sf[6] = a.desired_speed[row];
        }
    }
    // ---- pedestrian - obstacle ----
    if (a.M > 0) {
        TopK<KO> best;
        best.init();
        const int oframe = a.obs_frame_stride == 0 ? 0 : (a.obs_channel_T > 0 ? b / a.obs_channel_T : b);
        if (present) scan_cells<KO>(best, go, oframe, cx, cy, p.x, p.y, h.x, h.y, a.cos_o, a.thr_o, a.pre2_o, 0);
        const float2 *cand = reinterpret_cast<const float2 *>(a.obs + static_cast<int64_t>(oframe) * a.obs_frame_stride);
#pragma unroll
        for (int j = 0; j < KO; ++j) {
            if (j >= a.ko) break;
            const uint64_t w = best.key[j];
            float2 f0 = make_float2(0.f, 0.f), f1 = f0, f2 = f0;
            if (w != EMPTY_KEY) {
                const float2 om = cand[key_idx(w)];
                f0 = make_float2(__fsub_rn(om.x, p.x), __fsub_rn(om.y, p.y));
                f1 = make_float2(__fsub_rn(0.f, v.x), __fsub_rn(0.f, v.y));
                f2 = make_float2(__fsub_rn(0.f, ac.x), __fsub_rn(0.f, ac.y));
            }
            float2 *out = reinterpret_cast<float2 *>(a.obs_f) + (orow * a.ko + j) * 3;
            out[0] = f0; out[1] = f1; out[2] = f2;
            if (a.obs_idx) a.obs_idx[orow * a.ko + j] = (w != EMPTY_KEY) ? key_idx(w) : -1;
            if (a.obs_dist) a.obs_dist[orow * a.ko + j] = (w != EMPTY_KEY) ? key_dist(w) : CUDART_INF_F;
        }
    }
}

// ---- sorted-order kernel: G lanes per agent ------------------------------------------------------------------------
constexpr int SORT_G = 4;

// Lane `lane` of the agent's group keeps the k best gated candidates among every G-th record of the 9 buckets.
//
// Two passes, because a warp executes whatever ANY of its lanes needs: the exact gate (IEEE sqrt and two divisions) and
// the 64-bit top-k insertion are ~100 instructions, but only ~1 candidate in 8 survives.  Pass 1 walks the records with
// a cheap CONSERVATIVE test -- radius prefilter, then the cosine of the bearing from one approximate rsqrt:
//     c' = (r . h) * rsqrt(|r|^2),  |c' - c| < 2e-6 for |r|^2 >= 1e-12 (each exact operation of gated_distance rounds
//     a term bounded by 1 + eps; the approximate path adds the rsqrt's 2^-22 and two roundings)
// and parks the survivors (relative position, index) in a per-lane shared-memory list; one within GATE_MARGIN = 1e-4 of
// the threshold (or with a NaN / degenerate c') is flagged.  Pass 2 drains the list densely: exact distance (the sort
// key) for all, the exact gate only for the flagged ones.  The decisions are those of gated_distance, bit for bit.
constexpr float GATE_MARGIN = 1e-4f;
constexpr int BND_WORDS = 27;                  // per group: first record, end, packed cell of the 9 buckets

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// pending survivors of pass 1, pooled per GROUP in shared memory so that pass 2 is balanced over the group's lanes:
// [field][group][entry], group stride POOL_STRIDE words (bank = 4 * group + lane for the pass-2 reads: conflict-free)
constexpr int POOL_CAP = 32, POOL_STRIDE = POOL_CAP + 4, POOL_GROUPS = CELL_THREADS / SORT_G;
struct Pending {
    float *rx, *ry; uint32_t *id;              // id: candidate index | (gate undecided << 31)
    __device__ __forceinline__ Pending(float *base) {
        rx = base + (threadIdx.x / SORT_G) * POOL_STRIDE; ry = rx + POOL_GROUPS * POOL_STRIDE;
        id = reinterpret_cast<uint32_t *>(ry + POOL_GROUPS * POOL_STRIDE);
    }
};

// One flattened loop over the lane's share (every G-th record) of the 9 buckets, so that the lanes of a warp stay
// converged across bucket boundaries and the code stays small (an unrolled bucket loop with the insertion inlined nine
// times was 150 KB of SASS: the warps starved on instruction fetch).  The group's lanes share the 18 bucket-bound
// lookups through shared memory (`bnd`), and ONE drain site runs pass 2 -- for the whole warp at once, with the
// group's survivors dealt evenly to its lanes -- when a group's pool is nearly full or no lane has a record left.
template <int KMAX, int G>
__device__ __forceinline__ void scan_cells_split(TopK<KMAX> &best, const HashGrid &g, int64_t gframe, bool present,
                                                 int cx, int cy, int lane, float px, float py, float hx, float hy,
                                                 float cos_thr, float thr, float pre2, int idx_base, const Pending &pd,
                                                 int *bnd) {
    // bnd[c]: first record of bucket c MINUS the flat position where the bucket starts, bnd[9 + c]: flat end of bucket c,
    // bnd[18 + c]: its packed cell.  Flat position t = lane, lane + G, ... runs over the concatenation of the 9 ranges.
    const int *st = g.start + gframe * g.H;
    for (int c = lane; c < 9; c += G) {
        const int ccx = cx + (c % 3) - 1, ccy = cy + (c / 3) - 1;
        const uint32_t b = bucket_of(ccx, ccy, g.H);
        const int2 se = present ? make_int2(st[b], st[b + 1]) : make_int2(0, 0);
        bnd[c] = se.x;
        bnd[9 + c] = se.y - se.x;
        bnd[18 + c] = static_cast<int>(pack_cell(ccx, ccy));
    }
    __syncwarp();
    if (lane == 0) {                                               // lengths -> flat ends, starts -> offsets
        int run = 0;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            const int len = bnd[9 + c];
            bnd[c] -= run;
            run += len;
            bnd[9 + c] = run;
        }
    }
    __syncwarp();
    const float c_lo = cos_thr - GATE_MARGIN, c_hi = cos_thr + GATE_MARGIN;
    const int T = bnd[17];
    int c = 0, off = bnd[0], t_end = bnd[9];
    uint32_t want = static_cast<uint32_t>(bnd[18]);
    // The warp stays converged: every lane runs until no lane has a record left, and all lanes drain together.
    const int wl = threadIdx.x & 31;
    const unsigned gbits = ((1u << G) - 1u) << (wl & ~(G - 1)), below = (1u << wl) - 1u;
    int gcnt = 0;                                                  // the group's pending entries (same on its lanes)
    for (int t = lane;; t += G) {
        bool push = false;
        float rx = 0.f, ry = 0.f;
        uint32_t id = 0;
        if (t < T) {
            while (t >= t_end) {                                   // next non-empty bucket (c < 9 because t < T)
                ++c;
                off = bnd[c]; t_end = bnd[9 + c]; want = static_cast<uint32_t>(bnd[18 + c]);
            }
            const float4 r = g.rec[t + off];
            rx = __fsub_rn(r.x, px); ry = __fsub_rn(r.y, py);
            const float d2 = __fmaf_rn(ry, ry, __fmul_rn(rx, rx));
            const float ca = fmaf(rx, hx, ry * hy) * rsqrt_approx(d2);
            // same cell (else: hash collision or a wrapped neighbour cell), not certainly outside radius / field of view
            push = __float_as_uint(r.w) == want && d2 <= pre2 && !(ca < c_lo);
            id = static_cast<uint32_t>(__float_as_int(r.z) - idx_base) |
                 ((ca > c_hi && d2 >= 1e-12f) ? 0u : 0x80000000u);  // NaN compares false: undecided
        }
        const unsigned mine = __ballot_sync(0xffffffffu, push) & gbits;
        if (push) {
            const int pos = gcnt + __popc(mine & below);
            pd.rx[pos] = rx; pd.ry[pos] = ry; pd.id[pos] = id;
        }
        gcnt += __popc(mine);
        const bool more = __any_sync(0xffffffffu, t + G < T);
        if (!more || __any_sync(0xffffffffu, gcnt > POOL_CAP - G)) {
            __syncwarp();                                          // pass 2: exact distance (and gate where undecided)
            for (int q = lane; q < gcnt; q += G) {
                const float qx = pd.rx[q], qy = pd.ry[q];
                const uint32_t w = pd.id[q];
                const float d = (w >> 31) ? gated_distance(qx, qy, hx, hy, cos_thr) : norm2_rn(qx, qy);
                if (d <= thr) best.insert(make_key(d, static_cast<int>(w & 0x7fffffffu)));
            }
            __syncwarp();
            gcnt = 0;
            if (!more) break;
        }
    }
    __syncwarp();                                                  // `bnd` is reused by the next branch
}

// The group's k smallest keys in ascending order, dealt to the lanes round robin: slot j ends up in
// mine[j / G] of lane j % G.  Keys are unique (one candidate is seen by exactly one lane), EMPTY_KEY pads.
// Returns the number of non-empty keys among the k (the agent's live slots: a prefix of its slots).
template <int KMAX, int G>
__device__ __forceinline__ int merge_deal(TopK<KMAX> &best, int k, int lane, uint64_t (&mine)[(KMAX + G - 1) / G]) {
    int nlive = 0;
#pragma unroll
    for (int q = 0; q < (KMAX + G - 1) / G; ++q) mine[q] = EMPTY_KEY;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        if (j >= k) break;                                         // k is uniform: every lane leaves together
        const uint64_t h = best.key[0];
        const uint64_t w = group_min<G>(h);
        if (__all_sync(0xffffffffu, w == EMPTY_KEY)) break;       // nothing left in any group of the warp (uniform)
        if (h == w && w != EMPTY_KEY) best.pop_front();
        if ((j % G) == lane) mine[j / G] = w;
        nlive += (w != EMPTY_KEY) ? 1 : 0;
    }
    return nlive;
}

// Append the warp's live rows to a branch's compact row list; returns this lane's row (-1: not live).
__device__ __forceinline__ int compact_append(bool live, int *counter) {
    const unsigned mask = __ballot_sync(0xffffffffu, live);
    if (!mask) return -1;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return live ? base + __popc(mask & ((1u << lane) - 1)) : -1;
}

template <int KP, int KO>
__global__ void __launch_bounds__(CELL_THREADS) features_sorted_kernel(FeatArgs a, HashGrid gp, HashGrid go,
                                                                       double inv_cs, CompactOut co,
                                                                       const int *__restrict__ order,
                                                                       const int *__restrict__ n_order) {
    constexpr int G = SORT_G;
    __shared__ float pend_all[3 * POOL_GROUPS * POOL_STRIDE];
    __shared__ int bnd_all[BND_WORDS * (CELL_THREADS / SORT_G)];
    const Pending pend(pend_all);
    int *bnd = bnd_all + BND_WORDS * (threadIdx.x / SORT_G);
    // groups take the PRESENT agents in sorted order: all of them, or (agent-sharded ranks) the `order` list of this
    // rank's own agents (sorted positions, own_list_kernel); absent agents were served by the count kernel
    const int64_t n_present = gp.start[static_cast<int64_t>(a.B) * gp.H];
    const int64_t all_rows = order ? *n_order : n_present;
    const int64_t grp = (static_cast<int64_t>(blockIdx.x) * CELL_THREADS + threadIdx.x) / G;
    if (static_cast<int64_t>(blockIdx.x) * CELL_THREADS / G >= all_rows) return;   // whole CTA beyond the list (uniform)
    const int l = threadIdx.x % G;
    const bool valid = grp < all_rows;                             // no early exit: the group shuffles need every lane
    int64_t row = 0;
    float2 p = make_float2(CUDART_NAN_F, CUDART_NAN_F);
    if (valid) {
        const float4 r = gp.rec[order ? order[grp] : grp];
        row = __float_as_int(r.z);
        p = make_float2(r.x, r.y);
    }
    const int b = static_cast<int>(row / a.N);
    float2 v = make_float2(0.f, 0.f), ac = v;
    if (valid) {
        v = reinterpret_cast<const float2 *>(a.vel)[row];
        ac = reinterpret_cast<const float2 *>(a.acc)[row];
        if (l == 0) {                                              // in-place NaN -> 0 (data.py:483-484)
            if (v.x != v.x || v.y != v.y)
                reinterpret_cast<float2 *>(a.vel)[row] = make_float2(nan_to_zero(v.x), nan_to_zero(v.y));
            if (ac.x != ac.x || ac.y != ac.y)
                reinterpret_cast<float2 *>(a.acc)[row] = make_float2(nan_to_zero(ac.x), nan_to_zero(ac.y));
        }
    }
    v = make_float2(nan_to_zero(v.x), nan_to_zero(v.y));
    ac = make_float2(nan_to_zero(ac.x), nan_to_zero(ac.y));
    float2 h;
    if (a.head && valid) {
        h = reinterpret_cast<const float2 *>(a.head)[row];
    } else {
        float nv = norm2_rn(v.x, v.y);
        if (nv == 0.0f) nv = 0.1f;
        h = make_float2(__fdiv_rn(v.x, nv), __fdiv_rn(v.y, nv));
    }
    {
        const float nh = fmaxf(norm2_rn(h.x, h.y), 1e-8f);
        h = make_float2(__fdiv_rn(h.x, nh), __fdiv_rn(h.y, nh));
    }
    const bool present = valid && !(p.x != p.x || p.y != p.y);
    const int cx = present ? cell_coord(p.x, inv_cs) : 0, cy = present ? cell_coord(p.y, inv_cs) : 0;

    int live_ped = 0, live_obs = 0;
    // ---- pedestrian - pedestrian ----
    {
        TopK<KP> best;
        best.init();
        scan_cells_split<KP, G>(best, gp, b, present, cx, cy, l, p.x, p.y, h.x, h.y, a.cos_p, a.thr_p, a.pre2_p, b * a.N, pend, bnd);
        uint64_t mine[(KP + G - 1) / G];
        live_ped = merge_deal<KP, G>(best, a.kp, l, mine);
        const float2 *fp = reinterpret_cast<const float2 *>(a.pos) + static_cast<int64_t>(b) * a.N;
        const float2 *fv = reinterpret_cast<const float2 *>(a.vel) + static_cast<int64_t>(b) * a.N;
        const float2 *fa = reinterpret_cast<const float2 *>(a.acc) + static_cast<int64_t>(b) * a.N;
#pragma unroll
        for (int jj = 0; jj < (KP + G - 1) / G; ++jj) {
            if (jj * G >= a.kp) break;                             // uniform
            const int j = jj * G + l;
            const uint64_t w = mine[jj];
            const bool in = valid && j < a.kp;
            const bool live = in && w != EMPTY_KEY;
            float2 f0 = make_float2(0.f, 0.f), f1 = f0, f2 = f0;
            if (live) {
                const int m = key_idx(w);
                const float2 pm = fp[m], vm = fv[m], am = fa[m];
                f0 = make_float2(__fsub_rn(pm.x, p.x), __fsub_rn(pm.y, p.y));
                f1 = make_float2(__fsub_rn(nan_to_zero(vm.x), v.x), __fsub_rn(nan_to_zero(vm.y), v.y));
                f2 = make_float2(__fsub_rn(nan_to_zero(am.x), ac.x), __fsub_rn(nan_to_zero(am.y), ac.y));
            }
            if (in && a.ped_f) {
                float2 *out = reinterpret_cast<float2 *>(a.ped_f) + (row * a.kp + j) * 3;
                out[0] = f0; out[1] = f1; out[2] = f2;
                if (a.ped_idx) a.ped_idx[row * a.kp + j] = live ? key_idx(w) : -1;
                if (a.ped_dist) a.ped_dist[row * a.kp + j] = live ? key_dist(w) : CUDART_INF_F;
            }
            if (co.counts) {
                const int crow = compact_append(live, co.counts);
                if (live) {
                    float2 *out = reinterpret_cast<float2 *>(co.rows_ped) + static_cast<int64_t>(crow) * 3;
                    out[0] = f0; out[1] = f1; out[2] = f2;
                }
                if (live) co.map_ped[row * a.kp + j] = crow;
            }
        }
    }
    // ---- destination ----
    if (valid && l == 0 && a.dest_f) {
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
    // ---- pedestrian - obstacle ----
    if (a.M > 0) {
        TopK<KO> best;
        best.init();
        const int oframe = a.obs_frame_stride == 0 ? 0 : (a.obs_channel_T > 0 ? b / a.obs_channel_T : b);
        scan_cells_split<KO, G>(best, go, oframe, present, cx, cy, l, p.x, p.y, h.x, h.y, a.cos_o, a.thr_o, a.pre2_o, 0, pend, bnd);
        uint64_t mine[(KO + G - 1) / G];
        live_obs = merge_deal<KO, G>(best, a.ko, l, mine);
        const float2 *cand = reinterpret_cast<const float2 *>(a.obs + static_cast<int64_t>(oframe) * a.obs_frame_stride);
#pragma unroll
        for (int jj = 0; jj < (KO + G - 1) / G; ++jj) {
            if (jj * G >= a.ko) break;                             // uniform
            const int j = jj * G + l;
            const uint64_t w = mine[jj];
            const bool in = valid && j < a.ko;
            const bool live = in && w != EMPTY_KEY;
            float2 f0 = make_float2(0.f, 0.f), f1 = f0, f2 = f0;
            if (live) {
                const float2 om = cand[key_idx(w)];
                f0 = make_float2(__fsub_rn(om.x, p.x), __fsub_rn(om.y, p.y));
                f1 = make_float2(__fsub_rn(0.f, v.x), __fsub_rn(0.f, v.y));
                f2 = make_float2(__fsub_rn(0.f, ac.x), __fsub_rn(0.f, ac.y));
            }
            if (in && a.obs_f) {
                float2 *out = reinterpret_cast<float2 *>(a.obs_f) + (row * a.ko + j) * 3;
                out[0] = f0; out[1] = f1; out[2] = f2;
                if (a.obs_idx) a.obs_idx[row * a.ko + j] = live ? key_idx(w) : -1;
                if (a.obs_dist) a.obs_dist[row * a.ko + j] = live ? key_dist(w) : CUDART_INF_F;
            }
            if (co.counts) {
                const int crow = compact_append(live, co.counts + 1);
                if (live) {
                    float2 *out = reinterpret_cast<float2 *>(co.rows_obs) + static_cast<int64_t>(crow) * 3;
                    out[0] = f0; out[1] = f1; out[2] = f2;
                }
                if (live) co.map_obs[row * a.ko + j] = crow;
            }
        }
    }
    if (co.counts && valid && l == 0) co.live[row] = static_cast<uint16_t>(live_ped | (live_obs << 8));
}

// Agent-sharded ranks: the sorted positions of the PRESENT agents in rows [row0, row1) (warp-aggregated append: the
// entries a warp appends are neighbours in space, which is all the feature kernel's coherence needs).
__global__ void own_list_kernel(const float4 *__restrict__ rec, const int *__restrict__ start_end, int64_t row0,
                                int64_t row1, int *__restrict__ order, int *__restrict__ counter) {
    const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool own = false;
    if (s < *start_end) {
        const int64_t row = __float_as_int(rec[s].z);
        own = row >= row0 && row < row1;
    }
    const int pos = compact_append(own, counter);
    if (own) order[pos] = static_cast<int>(s);
}

// ---- host side -------------------------------------------------------------------------------------------------
struct ByteScratch { cudaStream_t st; int dev; char *buf; size_t cap; int64_t zeroed; };
static ByteScratch g_cell_scratch[8] = {};
static int g_cell_used = 0;
static std::mutex g_cell_mutex;                  // the cache is process-wide (freed by piml_free_workspace from any thread)

static int cell_scratch_get(cudaStream_t st, size_t bytes, char **out, ByteScratch **slot) {
    std::lock_guard<std::mutex> lock(g_cell_mutex);
    int dev = 0;
    PIML_CUDA(cudaGetDevice(&dev));
    ByteScratch *s = nullptr;
    for (int i = 0; i < g_cell_used; ++i)
        if (g_cell_scratch[i].st == st && g_cell_scratch[i].dev == dev) s = &g_cell_scratch[i];
    if (!s) {
        s = &g_cell_scratch[g_cell_used < 8 ? g_cell_used++ : 7];
        if (s->buf) { cudaSetDevice(s->dev); cudaFree(s->buf); cudaSetDevice(dev); }
        s->st = st; s->dev = dev; s->buf = nullptr; s->cap = 0; s->zeroed = 0;
    }
    if (s->cap < bytes) {
        if (s->buf) {
            PIML_CUDA(cudaStreamSynchronize(st));                  // kernels of an earlier call may still read it
            PIML_CUDA(cudaFree(s->buf));
        }
        s->buf = nullptr; s->cap = 0; s->zeroed = 0;
        const size_t want = bytes + bytes / 4;
        PIML_CUDA(cudaMalloc(&s->buf, want));
        s->cap = want;
    }
    *out = s->buf;
    *slot = s;
    return PIML_OK;
}

void cell_scratch_free() {
    std::lock_guard<std::mutex> lock(g_cell_mutex);
    for (int i = 0; i < g_cell_used; ++i)
        if (g_cell_scratch[i].buf) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaSetDevice(g_cell_scratch[i].dev);
            cudaFree(g_cell_scratch[i].buf);
            cudaSetDevice(dev);
            g_cell_scratch[i].buf = nullptr; g_cell_scratch[i].cap = 0; g_cell_scratch[i].zeroed = 0;
        }
    g_cell_used = 0;
}

static int pow2_at_least(int64_t x) {
    int h = 256;
    while (h < x && h < (1 << 28)) h <<= 1;
    return h;
}

static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// Cell-list evaluation of the features described by `a` (same contract as relative_features_kernel).
// obs_frames: number of distinct obstacle arrays (1 when shared by all frames).  co (optional, whole-row calls only):
// also emit the compact slot rows of the fused NN step; a.ped_f / a.obs_f / a.dest_f may then be null.
int relative_features_cells(const FeatArgs &a, int obs_frames, cudaStream_t st, const CompactOut *co) {
    const float thr = fmaxf(a.thr_p, a.M > 0 ? a.thr_o : 0.f);
    PIML_REQUIRE(thr > 0.f && thr < 1e18f, "cell-list features need a finite positive distance threshold");
    const double inv_cs = 1.0 / (static_cast<double>(thr) * (1.0 + 1e-5));
    PIML_REQUIRE(static_cast<int64_t>(a.B) * a.N < (1LL << 31) && static_cast<int64_t>(obs_frames) * a.M < (1LL << 31),
                 "cell-list features: too many points");
    PIML_REQUIRE(!(co && a.row1 > 0) || a.B == 1, "cell-list features: a row range with compact rows needs one frame");
    // one counting-sort chain for both point sets (see TwoSets)
    const int HP = pow2_at_least(2LL * a.N), HO = a.M > 0 ? pow2_at_least(2LL * a.M) : 0;
    const int64_t cellsP = static_cast<int64_t>(a.B) * HP, cellsO = a.M > 0 ? static_cast<int64_t>(obs_frames) * HO : 0;
    const int64_t cells = cellsP + cellsO;
    const int64_t totalP = static_cast<int64_t>(a.B) * a.N, totalO = a.M > 0 ? static_cast<int64_t>(obs_frames) * a.M : 0;
    const int nblocks = static_cast<int>((cells + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK);
    const size_t bytes = align256(sizeof(int) * (cells + 2)) + align256(sizeof(int) * (cells + 1)) +
                         align256(sizeof(int) * (totalP + totalO)) + align256(sizeof(int) * (nblocks + 1)) +
                         align256(sizeof(int) * totalP) + align256(sizeof(float4) * (totalP + totalO));
    char *base = nullptr;
    ByteScratch *slot = nullptr;
    int rc = cell_scratch_get(st, bytes, &base, &slot);
    if (rc) return rc;
    int *counts = reinterpret_cast<int *>(base); base += align256(sizeof(int) * (cells + 2));   // + own-list counter (+ 1 spare)
    int *start = reinterpret_cast<int *>(base); base += align256(sizeof(int) * (cells + 1));
    int *rank = reinterpret_cast<int *>(base); base += align256(sizeof(int) * (totalP + totalO));
    int *bsums = reinterpret_cast<int *>(base); base += align256(sizeof(int) * (nblocks + 1));
    int *order = reinterpret_cast<int *>(base); base += align256(sizeof(int) * totalP);
    float4 *rec = reinterpret_cast<float4 *>(base);
    TwoSets ts{reinterpret_cast<const float2 *>(a.pos), reinterpret_cast<const float2 *>(a.obs), totalP, totalO, a.N,
               a.M > 0 ? a.M : 1, HP, HO > 0 ? HO : 1, cellsP};
    // the fused scan leaves the counters zeroed for the next call with the same layout
    const bool fused = nblocks <= SCAN_FUSED_BLOCKS;
    if (slot->zeroed != cells + 2) {
        PIML_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (cells + 2), st));
        slot->zeroed = fused ? cells + 2 : 0;
    }
    const unsigned pblocks = static_cast<unsigned>((totalP + totalO + CELL_THREADS - 1) / CELL_THREADS);
    AbsentOut ab;
    ab.enabled = !(a.row1 > 0 && !co);                             // the row-range dense kernel serves its absent rows itself
    ab.a = a;
    ab.live = co ? co->live : nullptr;
    cell_count2_kernel<<<pblocks, CELL_THREADS, 0, st>>>(ts, inv_cs, counts, rank, ab);
    scan_block_sums_kernel<<<nblocks, 256, 0, st>>>(counts, cells, bsums);
    if (fused) {
        scan_apply_fused_kernel<<<nblocks, 256, 0, st>>>(counts, cells, bsums, start);
        count_launch(4);
    } else {
        scan_offsets_kernel<<<1, 1024, 0, st>>>(bsums, nblocks);
        scan_apply_kernel<<<nblocks, 256, 0, st>>>(counts, cells, bsums, start);
        count_launch(5);
    }
    cell_scatter2_kernel<<<pblocks, CELL_THREADS, 0, st>>>(ts, inv_cs, start, rank, rec);
    rc = check_launch("cell-list build");
    if (rc) return rc;
    HashGrid hp{HP, a.B, a.N, start, rec}, ho{0, 0, 0, nullptr, nullptr};
    if (a.M > 0) ho = HashGrid{HO, obs_frames, a.M, start + cellsP, rec};
    if (a.row1 > 0 && !co) {                                       // row range: one thread per agent, index order
        const int64_t rows = a.row1 - a.row0;
        const unsigned blocks = static_cast<unsigned>((rows + CELL_THREADS - 1) / CELL_THREADS);
        if (a.kp <= 8 && a.ko <= 16) features_cells_kernel<8, 16><<<blocks, CELL_THREADS, 0, st>>>(a, hp, ho, inv_cs);
        else if (a.kp <= 16 && a.ko <= 16) features_cells_kernel<16, 16><<<blocks, CELL_THREADS, 0, st>>>(a, hp, ho, inv_cs);
        else features_cells_kernel<32, 32><<<blocks, CELL_THREADS, 0, st>>>(a, hp, ho, inv_cs);
        count_launch();
        return check_launch("features_cells_kernel");
    }
    CompactOut c{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (co) c = *co;
    const int *ord = nullptr;
    int64_t groups = totalP;
    if (a.row1 > 0) {                                              // agent-sharded rank: its own agents, in sorted order
        own_list_kernel<<<static_cast<unsigned>((totalP + 255) / 256), 256, 0, st>>>(rec, start + cellsP, a.row0, a.row1,
                                                                                     order, counts + cells + 1);
        count_launch();
        ord = order;
        groups = a.row1 - a.row0;
    }
    const int *n_ord = counts + cells + 1;                         // own-list length (device)
    const unsigned blocks = static_cast<unsigned>((groups * SORT_G + CELL_THREADS - 1) / CELL_THREADS);
    if (a.kp <= 6 && a.ko <= 10) features_sorted_kernel<6, 10><<<blocks, CELL_THREADS, 0, st>>>(a, hp, ho, inv_cs, c, ord, n_ord);   // the reference's topk
    else if (a.kp <= 8 && a.ko <= 16) features_sorted_kernel<8, 16><<<blocks, CELL_THREADS, 0, st>>>(a, hp, ho, inv_cs, c, ord, n_ord);
    else if (a.kp <= 16 && a.ko <= 16) features_sorted_kernel<16, 16><<<blocks, CELL_THREADS, 0, st>>>(a, hp, ho, inv_cs, c, ord, n_ord);
    else features_sorted_kernel<32, 32><<<blocks, CELL_THREADS, 0, st>>>(a, hp, ho, inv_cs, c, ord, n_ord);
    count_launch();
    return check_launch("features_sorted_kernel");
}

}  // namespace piml
