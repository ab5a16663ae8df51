// mlp.cu -- fused forward of the physics-infused interaction networks for sm_100a.
//
// Replaces MLP / ResBlock / ResDNN (reference src/models/model.py:40-119) and the forwards of PINNSF (:762-792),
// PINNSF_bottleneck (:1104-1135), PINNSF_bottleneck_multitask (:1185-1221) and PINNSF_multitask (:1271-1305).
// The reference runs 9-11 MKL/cuBLAS addmm calls plus elementwise kernels per forward.  Here ONE kernel takes a tile
// of up to 128 neighbour-slot rows (whole agents of one branch: 21 agents x 6 pedestrian slots, or 12 agents x 10
// obstacle slots) through encoder -> processor -> decoder -> predictor (-> collision head):
//   * activations live in shared memory TRANSPOSED ([feature][row]), ping-ponged between two 128x132 buffers;
//   * every Linear is a register-tiled GEMM: 256 threads x (8 rows x 8 outputs), packed FP32 FMAs (FFMA2) with the
//     weight as the broadcast operand, fp32 accumulation (1e-5 parity with the reference rules out TF32);
//   * weights stream from L2 in 32-row chunks by TMA bulk copies (cp.async.bulk, double buffered on mbarriers), the
//     next chunk -- also across layer boundaries -- in flight while the current one is consumed;
//   * weights come pre-permuted (piml_pinnsf_pack_f32) so that each thread's 8 output columns {tc + 16 j} are two
//     conflict-free float4 loads and the transposed stores of a warp hit 32 distinct banks.
// Per-agent message sums go to a small workspace; pinnsf_finish_kernel adds the destination (social-force driving)
// term.  Every agent's sum is produced by exactly one CTA in a fixed order: results are run-to-run deterministic.
//
// Reference quirks reproduced on purpose: ResDNN with >1 layers is exactly 2*x (B-4);
// zero-padded neighbour slots are NOT masked and contribute f(0) (B-5); the destination norm of a channelled
// (C,N,7) input reduces over the agent axis (B-3) -- handled by dest_colnorm_kernel.
#include <math_constants.h>

#include "common.cuh"

namespace piml {

constexpr int FT_THREADS = 256;
constexpr int FT_TR = 128;            // slot rows per tile
constexpr int FT_TRP = 132;           // row stride of the transposed activation buffers (132 % 32 == 4)
constexpr int FT_KC = 32;             // weight rows per TMA chunk
constexpr int FT_MAXW = 128;          // widest supported layer
constexpr int FT_MAXCH = 64;          // chunks per branch
constexpr int FT_SMALL = 3;           // rows of the small output buffer: predictor (2) + collision logit (1)

struct FLayer { int K, OUT, OUTP, NJ, w_off, b_off; };   // offsets in floats, relative to the branch / head base
struct FChunk { int off, bytes; };                        // absolute float offset into the packed vector

struct FPlan {
    int n_enc; FLayer enc[8];
    int proc_mode; FLayer proc;
    int n_dec; FLayer dec[8];
    FLayer pred;
    int n_coll; FLayer coll[2];
    int branch_off[2], coll_off;
    int kind, pw, dw;
    int64_t total;                    // floats in the packed vector
};

struct FTab { int n[2]; FChunk c[2][FT_MAXCH]; };

struct FArgs {
    const float *params; const float *ped; const float *obs; const float *drop_ped; const float *drop_obs;
    int64_t R; int kp, ko, ag_ped, ag_obs; int64_t n_ped_tiles, n_obs_tiles;
    float *sums;                      // (R,4): ped.x ped.y obs.x obs.y
    float *ped_msgs; float *obs_msgs; float *coll;
};

// column of output o in a permuted weight row (see header comment)
__host__ __device__ __forceinline__ int perm_col(int o, int NJ) {
    const int j = o >> 4, tc = o & 15;
    return NJ == 8 ? ((j >> 2) * 64 + tc * 4 + (j & 3)) : (tc * NJ + j);
}

// Weight-chunk pipeline state (uniform across the CTA).
struct WPipe {
    const float *base; const FChunk *tab; int n; int cons; float *wbuf; uint64_t *bars; uint32_t phase;
    __device__ __forceinline__ void issue(int i) {                 // one thread
        const FChunk ch = tab[i];
        mbar_expect_tx(&bars[i & 1], static_cast<uint32_t>(ch.bytes));
        tma_bulk_g2s(wbuf + (i & 1) * FT_KC * FT_MAXW, base + ch.off, static_cast<uint32_t>(ch.bytes), &bars[i & 1]);
    }
    // Wait for chunk `cons`, release the other buffer and refill it with chunk cons+1.  All threads call it.
    __device__ __forceinline__ const float *acquire() {
        const int b = cons & 1;
        mbar_wait(&bars[b], (phase >> b) & 1u);
        phase ^= (1u << b);
        __syncthreads();              // everyone is done with chunk cons-1 (and with the previous layer's input)
        if (threadIdx.x == 0 && cons + 1 < n) issue(cons + 1);
        ++cons;
        return wbuf + b * FT_KC * FT_MAXW;
    }
};

// Ys[o][r] = act(scale * (b[o] + sum_k W[k][o] Xs[k][r])) (+ Xs[o][r]) (* drop) for r < nrows.  All threads call it.
template <int NJ>
__device__ __forceinline__ void dense(WPipe &wp, const FLayer &L, const float *__restrict__ pbase, const float *Xs,
                                      float *Ys, int nrows, bool relu, float scale, bool residual,
                                      const float *__restrict__ drop, int drop_ld) {
    const int tc = threadIdx.x & 15, tr = threadIdx.x >> 4;
    const bool active = (tr >> 1) * 16 < nrows;                    // warp-uniform: this warp's 16 rows hold data
    float2 acc[NJ][4];
    const float *bias = pbase + L.b_off;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int o = tc + 16 * j;
        const float b = o < L.OUT ? bias[o] : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = make_float2(b, b);
    }
    const int nch = (L.K + FT_KC - 1) / FT_KC;
    for (int c = 0; c < nch; ++c) {
        const float *wb = wp.acquire();
        if (!active) continue;
        const int kc = min(FT_KC, L.K - c * FT_KC);
        const float *xr = Xs + (c * FT_KC) * FT_TRP + tr * 8;
        const float *wr = wb + (NJ == 8 ? tc * 4 : tc * NJ);
#pragma unroll 4
        for (int kk = 0; kk < kc; ++kk) {
            const float4 xa = *reinterpret_cast<const float4 *>(xr + kk * FT_TRP);
            const float4 xb = *reinterpret_cast<const float4 *>(xr + kk * FT_TRP + 4);
            float w[NJ];
            if constexpr (NJ == 8) {
                const float4 wa = *reinterpret_cast<const float4 *>(wr + kk * 128);
                const float4 wc = *reinterpret_cast<const float4 *>(wr + kk * 128 + 64);
                w[0] = wa.x; w[1] = wa.y; w[2] = wa.z; w[3] = wa.w;
                w[4] = wc.x; w[5] = wc.y; w[6] = wc.z; w[7] = wc.w;
            } else if constexpr (NJ == 4) {
                const float4 wa = *reinterpret_cast<const float4 *>(wr + kk * 64);
                w[0] = wa.x; w[1] = wa.y; w[2] = wa.z; w[3] = wa.w;
            } else if constexpr (NJ == 2) {
                const float2 wa = *reinterpret_cast<const float2 *>(wr + kk * 32);
                w[0] = wa.x; w[1] = wa.y;
            } else {
                w[0] = wr[kk * 16];
            }
            const float2 x0 = make_float2(xa.x, xa.y), x1 = make_float2(xa.z, xa.w);
            const float2 x2 = make_float2(xb.x, xb.y), x3 = make_float2(xb.z, xb.w);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float2 ww = make_float2(w[j], w[j]);
                acc[j][0] = __ffma2_rn(x0, ww, acc[j][0]);
                acc[j][1] = __ffma2_rn(x1, ww, acc[j][1]);
                acc[j][2] = __ffma2_rn(x2, ww, acc[j][2]);
                acc[j][3] = __ffma2_rn(x3, ww, acc[j][3]);
            }
        }
    }
    if (!active) return;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int o = tc + 16 * j;
        if (o >= L.OUT) continue;
        float y[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) { y[2 * q] = acc[j][q].x * scale; y[2 * q + 1] = acc[j][q].y * scale; }
        if (relu) {
#pragma unroll
            for (int q = 0; q < 8; ++q) y[q] = fmaxf(y[q], 0.f);
        }
        if (residual) {                                            // ResBlock: lin(x) + x   (model.py:78-79)
#pragma unroll
            for (int q = 0; q < 8; ++q) y[q] += Xs[o * FT_TRP + tr * 8 + q];
        }
        if (drop) {                                                // Dropout multipliers in train()  (model.py:118)
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (tr * 8 + q < nrows) y[q] *= drop[static_cast<int64_t>(tr * 8 + q) * drop_ld + o];
        }
        float4 *dst = reinterpret_cast<float4 *>(Ys + o * FT_TRP + tr * 8);
        dst[0] = make_float4(y[0], y[1], y[2], y[3]);
        dst[1] = make_float4(y[4], y[5], y[6], y[7]);
    }
}

__device__ __noinline__ void dense_any(WPipe &wp, const FLayer &L, const float *pbase, const float *Xs, float *Ys,
                                          int nrows, bool relu, float scale, bool residual = false,
                                          const float *drop = nullptr, int drop_ld = 0) {
    switch (L.NJ) {
        case 8: dense<8>(wp, L, pbase, Xs, Ys, nrows, relu, scale, residual, drop, drop_ld); break;
        case 4: dense<4>(wp, L, pbase, Xs, Ys, nrows, relu, scale, residual, drop, drop_ld); break;
        case 2: dense<2>(wp, L, pbase, Xs, Ys, nrows, relu, scale, residual, drop, drop_ld); break;
        default: dense<1>(wp, L, pbase, Xs, Ys, nrows, relu, scale, residual, drop, drop_ld); break;
    }
}

__global__ void __launch_bounds__(FT_THREADS, 1) pinnsf_tile_kernel(const __grid_constant__ FPlan P,
                                                                    const __grid_constant__ FTab T,
                                                                    const __grid_constant__ FArgs a) {
    extern __shared__ __align__(128) float smem[];
    float *bufA = smem;
    float *bufB = bufA + FT_MAXW * FT_TRP;
    float *wbuf = bufB + FT_MAXW * FT_TRP;
    float *small = wbuf + 2 * FT_KC * FT_MAXW;                    // [FT_SMALL][FT_TRP]
    uint64_t *bars = reinterpret_cast<uint64_t *>(small + FT_SMALL * FT_TRP);
    const int tid = threadIdx.x;

    const int64_t tile = blockIdx.x;
    const int br = tile < a.n_ped_tiles ? 0 : 1;
    const int k = br == 0 ? a.kp : a.ko;
    const int AG = br == 0 ? a.ag_ped : a.ag_obs;
    const int64_t agent0 = (br == 0 ? tile : tile - a.n_ped_tiles) * AG;
    const int na = static_cast<int>(min(static_cast<int64_t>(AG), a.R - agent0));
    const int nrows = na * k;
    const int64_t row0 = agent0 * k;
    const float *pbase = a.params + P.branch_off[br];
    const bool want_coll = br == 0 && P.n_coll > 0 && a.coll != nullptr;

    WPipe wp;
    wp.base = a.params; wp.tab = T.c[br]; wp.n = T.n[br]; wp.cons = 0; wp.wbuf = wbuf; wp.bars = bars; wp.phase = 0;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
        wp.issue(0);
    }
    // stage the 6-d features transposed: bufA[c][r]
    {
        const float *feat = (br == 0 ? a.ped : a.obs) + row0 * 6;
        for (int e = tid; e < nrows * 6; e += FT_THREADS) bufA[(e % 6) * FT_TRP + (e / 6)] = feat[e];
    }
    __syncthreads();                                               // barrier init + features visible
    float *cur = bufA, *oth = bufB;
    const float *drop = br == 0 ? a.drop_ped : a.drop_obs;
    if (drop) drop += row0 * P.pw;

    for (int l = 0; l < P.n_enc; ++l) {      // MLP: ReLU between layers, Identity at the end (model.py:54-61)
        const bool last = l == P.n_enc - 1;
        const bool fold = last && P.proc_mode == 0;                // ResDNN == 2x (+ dropout)
        dense_any(wp, P.enc[l], pbase, cur, oth, nrows, !last, fold ? 2.f : 1.f, false, fold ? drop : nullptr, P.pw);
        float *t = cur; cur = oth; oth = t;
    }
    if (P.proc_mode == 1) {                  // single ResBlock: relu(Wx+b) + x, then dropout (model.py:68-79,118)
        dense_any(wp, P.proc, pbase, cur, oth, nrows, true, 1.f, true, drop, P.pw);
        float *t = cur; cur = oth; oth = t;
    }
    float *msgs_out = br == 0 ? a.ped_msgs : a.obs_msgs;

    if (P.kind == 0) {
        // per-slot decoder -> predictor; messages are 2-d (model.py:1190-1194)
        for (int l = 0; l < P.n_dec; ++l) {
            dense_any(wp, P.dec[l], pbase, cur, oth, nrows, l < P.n_dec - 1, 1.f);
            float *t = cur; cur = oth; oth = t;
        }
        dense_any(wp, P.pred, pbase, cur, small, nrows, false, 1.f);
        if (want_coll) {                     // collision head on the decoder output (model.py:1214-1215)
            const float *cb = a.params + P.coll_off;
            const float *h = cur;
            for (int l = 0; l < P.n_coll; ++l) {
                const bool last = l == P.n_coll - 1;
                dense_any(wp, P.coll[l], cb, h, last ? small + 2 * FT_TRP : oth, nrows, !last, 1.f);
                h = oth;
            }
        }
        __syncthreads();
        if (want_coll)
            for (int r = tid; r < nrows; r += FT_THREADS) a.coll[row0 + r] = 1.f / (1.f + expf(-small[2 * FT_TRP + r]));
        if (msgs_out)
            for (int e = tid; e < nrows * 2; e += FT_THREADS) msgs_out[row0 * 2 + e] = small[(e & 1) * FT_TRP + (e >> 1)];
        if (tid < 2 * na) {                  // torch.sum(dim=-2) over the k slots (model.py:1194/1202)
            const int ag = tid >> 1, c = tid & 1;
            float s = 0.f;
            for (int j = 0; j < k; ++j) s += small[c * FT_TRP + ag * k + j];
            a.sums[(agent0 + ag) * 4 + br * 2 + c] = s;
        }
    } else {
        // messages are the processor outputs; sum over slots, then decode per agent (model.py:1276-1279)
        __syncthreads();
        if (msgs_out)
            for (int e = tid; e < nrows * P.pw; e += FT_THREADS) {
                const int r = e / P.pw, i = e % P.pw;
                msgs_out[(row0 + r) * P.pw + i] = cur[i * FT_TRP + r];
            }
        if (want_coll) {                     // collision head on the per-slot messages (model.py:1298-1299)
            const float *cb = a.params + P.coll_off;
            const float *h = cur;
            for (int l = 0; l < P.n_coll; ++l) {
                const bool last = l == P.n_coll - 1;
                dense_any(wp, P.coll[l], cb, h, last ? small + 2 * FT_TRP : oth, nrows, !last, 1.f);
                h = oth;
            }
            __syncthreads();
            for (int r = tid; r < nrows; r += FT_THREADS) a.coll[row0 + r] = 1.f / (1.f + expf(-small[2 * FT_TRP + r]));
        }
        for (int e = tid; e < na * P.pw; e += FT_THREADS) {
            const int i = e / na, ag = e % na;
            float s = 0.f;
            for (int j = 0; j < k; ++j) s += cur[i * FT_TRP + ag * k + j];
            oth[i * FT_TRP + ag] = s;
        }
        { float *t = cur; cur = oth; oth = t; }
        for (int l = 0; l < P.n_dec; ++l) {
            dense_any(wp, P.dec[l], pbase, cur, oth, na, l < P.n_dec - 1, 1.f);
            float *t = cur; cur = oth; oth = t;
        }
        dense_any(wp, P.pred, pbase, cur, small, na, false, 1.f);
        __syncthreads();
        if (tid < 2 * na) {
            const int ag = tid >> 1, c = tid & 1;
            a.sums[(agent0 + ag) * 4 + br * 2 + c] = small[c * FT_TRP + ag];
        }
    }
}

// acc = sum_ped + sum_obs + (v0 * dest/||dest|| - v) / tau      (model.py:1205-1212)
__global__ void pinnsf_finish_kernel(const float *__restrict__ sums, const float *__restrict__ self,
                                     const float *__restrict__ dnorm, int64_t R, int has_obs, float tau,
                                     float *__restrict__ acc) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= 2 * R) return;
    const int64_t ag = i >> 1;
    const int c = static_cast<int>(i & 1);
    const float *s = self + ag * 7;
    float nrm = dnorm ? dnorm[ag * 2 + c] : norm2_rn(s[0], s[1]);
    if (nrm == 0.f) nrm = __fadd_rn(nrm, 0.1f);
    const float dir = __fdiv_rn(s[c], nrm);
    const float dterm = __fdiv_rn(__fsub_rn(__fmul_rn(s[6], dir), s[2 + c]), tau);
    float m = sums[ag * 4 + c];
    if (has_obs) m = __fadd_rn(m, sums[ag * 4 + 2 + c]);
    acc[i] = __fadd_rn(m, dterm);
}

// column norms over the agent axis for channelled (C,N,7) inputs: dnorm[(c*N+n)*2 + q] = ||self[c,:,q]||_2
__global__ void dest_colnorm_kernel(const float *__restrict__ self, int group, float *__restrict__ dnorm) {
    __shared__ float red[2][128];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * group;
    float s0 = 0.f, s1 = 0.f;
    for (int i = threadIdx.x; i < group; i += blockDim.x) {
        const float x = self[(base + i) * 7], y = self[(base + i) * 7 + 1];
        s0 = fmaf(x, x, s0); s1 = fmaf(y, y, s1);
    }
    red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            red[0][threadIdx.x] += red[0][threadIdx.x + off];
            red[1][threadIdx.x] += red[1][threadIdx.x + off];
        }
        __syncthreads();
    }
    const float n0 = sqrtf(red[0][0]), n1 = sqrtf(red[1][0]);
    for (int i = threadIdx.x; i < group; i += blockDim.x) {
        dnorm[(base + i) * 2] = n0;
        dnorm[(base + i) * 2 + 1] = n1;
    }
}

// ---- parameter packing: torch order (per Linear: W (out,in) row-major, b) -> device layout ------------------------
struct PackRec { int K, OUT, OUTP, NJ; int64_t src, dst; };       // dst: Wp [K][OUTP] then bias [OUTP]
struct PackTab { int n; PackRec r[48]; };

__global__ void pinnsf_pack_kernel(const __grid_constant__ PackTab T, const float *__restrict__ src,
                                   float *__restrict__ dst, int64_t total) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int l = 0;
    while (l + 1 < T.n && i >= T.r[l + 1].dst) ++l;
    const PackRec &L = T.r[l];
    const int64_t e = i - L.dst;
    float v = 0.f;
    if (e < static_cast<int64_t>(L.K) * L.OUTP) {
        const int kk = static_cast<int>(e / L.OUTP), cp = static_cast<int>(e % L.OUTP);
        // invert perm_col
        int o;
        if (L.NJ == 8) { const int half = cp >> 6, rem = cp & 63; o = (rem >> 2) + 16 * (half * 4 + (rem & 3)); }
        else { o = (cp / L.NJ) + 16 * (cp % L.NJ); }
        if (o < L.OUT) v = src[L.src + static_cast<int64_t>(o) * L.K + kk];
    } else {
        const int o = static_cast<int>(e - static_cast<int64_t>(L.K) * L.OUTP);
        if (o < L.OUT) v = src[L.src + static_cast<int64_t>(L.OUT) * L.K + o];
    }
    dst[i] = v;
}

static int nj_for(int out) { return out > 64 ? 8 : (out > 32 ? 4 : (out > 16 ? 2 : 1)); }

// Builds the layer plan, the packing table and (optionally) the chunk tables.
static int build_plan(const piml_net_desc *d, FPlan *P, PackTab *PT) {
    PIML_REQUIRE(d->n_enc >= 1 && d->n_enc <= 8 && d->n_dec >= 1 && d->n_dec <= 8 && d->n_coll >= 0 && d->n_coll <= 2,
                 "piml_pinnsf: unsupported layer counts (enc %d, dec %d, coll %d)", d->n_enc, d->n_dec, d->n_coll);
    PIML_REQUIRE(d->kind == 0 || d->kind == 1, "piml_pinnsf: kind must be 0 or 1");
    PIML_REQUIRE(d->proc_mode == 0 || d->proc_mode == 1, "piml_pinnsf: proc_mode must be 0 or 1");
    PIML_REQUIRE(d->enc_dims[0] == 6, "piml_pinnsf: feature dim must be 6, got %d", d->enc_dims[0]);
    int maxw = 6;
    int64_t src = 0;
    int dst = 0;
    PT->n = 0;
    auto rec = [&](int in, int out, int base, bool record) {
        FLayer L;
        L.K = in; L.OUT = out; L.NJ = nj_for(out); L.OUTP = 16 * L.NJ;
        L.w_off = dst - base; L.b_off = L.w_off + in * L.OUTP;
        if (record) {
            PackRec r{in, out, L.OUTP, L.NJ, src, dst};
            PT->r[PT->n++] = r;
        }
        src += static_cast<int64_t>(in) * out + out;
        dst += in * L.OUTP + L.OUTP;
        if (in > maxw) maxw = in;
        if (out > maxw) maxw = out;
        return L;
    };
    for (int l = 0; l <= d->n_enc; ++l) PIML_REQUIRE(d->enc_dims[l] >= 1, "piml_pinnsf: bad encoder width");
    for (int l = 0; l <= d->n_dec; ++l) PIML_REQUIRE(d->dec_dims[l] >= 1, "piml_pinnsf: bad decoder width");
    P->n_enc = d->n_enc; P->n_dec = d->n_dec; P->proc_mode = d->proc_mode; P->n_coll = d->n_coll; P->kind = d->kind;
    P->pw = d->enc_dims[d->n_enc];
    P->dw = d->dec_dims[d->n_dec];
    PIML_REQUIRE(d->dec_dims[0] == P->pw, "piml_pinnsf: decoder input %d != processor width %d", d->dec_dims[0], P->pw);
    for (int br = 0; br < 2; ++br) {
        const int base = dst;
        P->branch_off[br] = base;
        for (int l = 0; l < d->n_enc; ++l) P->enc[l] = rec(d->enc_dims[l], d->enc_dims[l + 1], base, true);
        if (d->proc_mode == 1) P->proc = rec(P->pw, P->pw, base, true);
        for (int l = 0; l < d->n_dec; ++l) P->dec[l] = rec(d->dec_dims[l], d->dec_dims[l + 1], base, true);
        P->pred = rec(P->dw, 2, base, true);
    }
    P->coll_off = dst;
    if (d->n_coll) {
        PIML_REQUIRE(d->coll_dims[d->n_coll] == 1, "piml_pinnsf: collision head must end in width 1");
        PIML_REQUIRE(d->coll_dims[0] == (d->kind == 0 ? P->dw : P->pw),
                     "piml_pinnsf: collision head input width %d does not match", d->coll_dims[0]);
        const int base = dst;
        for (int l = 0; l < d->n_coll; ++l) P->coll[l] = rec(d->coll_dims[l], d->coll_dims[l + 1], base, true);
    }
    P->total = dst;
    PIML_REQUIRE(maxw <= FT_MAXW, "piml_pinnsf: layer width %d > %d is not supported", maxw, FT_MAXW);
    return PIML_OK;
}

// chunk consumption order of one branch -- must mirror pinnsf_tile_kernel exactly
static int build_chunks(const FPlan &P, int br, bool want_coll, FTab *T) {
    int n = 0;
    auto add = [&](const FLayer &L, int base) -> int {
        const int nch = (L.K + FT_KC - 1) / FT_KC;
        for (int c = 0; c < nch; ++c) {
            if (n >= FT_MAXCH) return -1;
            const int kc = L.K - c * FT_KC < FT_KC ? L.K - c * FT_KC : FT_KC;
            T->c[br][n].off = base + L.w_off + c * FT_KC * L.OUTP;
            T->c[br][n].bytes = kc * L.OUTP * static_cast<int>(sizeof(float));
            ++n;
        }
        return 0;
    };
    const int base = P.branch_off[br];
    int bad = 0;
    for (int l = 0; l < P.n_enc; ++l) bad |= add(P.enc[l], base);
    if (P.proc_mode == 1) bad |= add(P.proc, base);
    const bool coll = want_coll && br == 0 && P.n_coll > 0;
    if (P.kind == 0) {
        for (int l = 0; l < P.n_dec; ++l) bad |= add(P.dec[l], base);
        bad |= add(P.pred, base);
        if (coll) for (int l = 0; l < P.n_coll; ++l) bad |= add(P.coll[l], P.coll_off);
    } else {
        if (coll) for (int l = 0; l < P.n_coll; ++l) bad |= add(P.coll[l], P.coll_off);
        for (int l = 0; l < P.n_dec; ++l) bad |= add(P.dec[l], base);
        bad |= add(P.pred, base);
    }
    T->n[br] = n;
    return bad;
}

// Stream-keyed scratch (per-agent sums and channelled destination norms); grows on demand, freed at process exit.
struct Scratch { cudaStream_t st; int dev; float *buf; int64_t cap; };
static int scratch_get(cudaStream_t st, int64_t floats, float **out) {
    static thread_local Scratch slots[8] = {};
    static thread_local int used = 0;
    int dev = 0;
    PIML_CUDA(cudaGetDevice(&dev));
    Scratch *s = nullptr;
    for (int i = 0; i < used; ++i)
        if (slots[i].st == st && slots[i].dev == dev) s = &slots[i];
    if (!s) {
        s = &slots[used < 8 ? used++ : 7];
        if (s->buf) { cudaSetDevice(s->dev); cudaFree(s->buf); cudaSetDevice(dev); }
        s->st = st; s->dev = dev; s->buf = nullptr; s->cap = 0;
    }
    if (s->cap < floats) {
        if (s->buf) PIML_CUDA(cudaFree(s->buf));
        s->buf = nullptr; s->cap = 0;
        PIML_CUDA(cudaMalloc(&s->buf, sizeof(float) * floats));
        s->cap = floats;
    }
    *out = s->buf;
    return PIML_OK;
}

}  // namespace piml

using namespace piml;

extern "C" int64_t piml_pinnsf_packed_floats(const piml_net_desc *desc) {
    if (!desc) return -1;
    FPlan P;
    PackTab PT;
    if (build_plan(desc, &P, &PT)) return -1;
    return P.total;
}

extern "C" int piml_pinnsf_pack_f32(const piml_net_desc *desc, const float *params_torch, float *packed, void *stream) {
    PIML_REQUIRE(desc && params_torch && packed, "piml_pinnsf_pack_f32: null pointer");
    FPlan P;
    PackTab PT;
    int rc = build_plan(desc, &P, &PT);
    if (rc) return rc;
    PIML_REQUIRE(aligned16(packed), "piml_pinnsf_pack_f32: packed must be 16-byte aligned");
    const int threads = 256;
    pinnsf_pack_kernel<<<static_cast<unsigned>((P.total + threads - 1) / threads), threads, 0,
                         static_cast<cudaStream_t>(stream)>>>(PT, params_torch, packed, P.total);
    count_launch();
    return check_launch("pinnsf_pack_kernel");
}

extern "C" int piml_pinnsf_forward_f32(const piml_net_desc *desc, const float *params, int has_obs, float tau,
                                       const float *ped, const float *obs, const float *self, int64_t R, int kp,
                                       int ko, int norm_group, const float *drop_ped, const float *drop_obs,
                                       float *acc, float *ped_msgs, float *obs_msgs, float *coll, void *stream) {
    PIML_REQUIRE(desc && params && ped && self && acc, "piml_pinnsf_forward_f32: null pointer");
    PIML_REQUIRE(!has_obs || obs, "piml_pinnsf_forward_f32: has_obs set but obs is null");
    PIML_REQUIRE(R >= 0 && kp >= 1 && ko >= 0, "piml_pinnsf_forward_f32: bad dimensions R=%lld kp=%d ko=%d",
                 static_cast<long long>(R), kp, ko);
    PIML_REQUIRE(kp <= FT_TR && ko <= FT_TR, "piml_pinnsf_forward_f32: more than %d slots per agent", FT_TR);
    PIML_REQUIRE(aligned16(params), "piml_pinnsf_forward_f32: params must be 16-byte aligned (piml_pinnsf_pack_f32)");
    if (!has_obs || ko == 0) { has_obs = 0; ko = 0; }
    FPlan P;
    PackTab PT;
    int rc = build_plan(desc, &P, &PT);
    if (rc) return rc;
    if (R == 0) return PIML_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    FTab T;
    T.n[1] = 0;
    PIML_REQUIRE(build_chunks(P, 0, coll != nullptr, &T) == 0 && (!has_obs || build_chunks(P, 1, false, &T) == 0),
                 "piml_pinnsf_forward_f32: network needs more than %d weight chunks per branch", FT_MAXCH);

    float *scratch = nullptr;
    rc = scratch_get(st, R * 4 + (norm_group > 0 ? R * 2 : 0), &scratch);
    if (rc) return rc;
    float *sums = scratch;
    const float *dnorm = nullptr;
    if (norm_group > 0) {
        PIML_REQUIRE(R % norm_group == 0, "piml_pinnsf_forward_f32: R=%lld not a multiple of norm_group=%d",
                     static_cast<long long>(R), norm_group);
        float *dn = scratch + R * 4;
        dest_colnorm_kernel<<<static_cast<unsigned>(R / norm_group), 128, 0, st>>>(self, norm_group, dn);
        count_launch();
        rc = check_launch("dest_colnorm_kernel");
        if (rc) return rc;
        dnorm = dn;
    }

    FArgs a;
    a.params = params; a.ped = ped; a.obs = obs; a.drop_ped = drop_ped; a.drop_obs = drop_obs;
    a.R = R; a.kp = kp; a.ko = ko;
    a.ag_ped = FT_TR / kp; a.ag_obs = ko ? FT_TR / ko : 1;
    a.n_ped_tiles = (R + a.ag_ped - 1) / a.ag_ped;
    a.n_obs_tiles = has_obs ? (R + a.ag_obs - 1) / a.ag_obs : 0;
    a.sums = sums; a.ped_msgs = ped_msgs; a.obs_msgs = has_obs ? obs_msgs : nullptr; a.coll = coll;
    const size_t smem = sizeof(float) * (2 * FT_MAXW * FT_TRP + 2 * FT_KC * FT_MAXW + FT_SMALL * FT_TRP) + 16;
    static thread_local bool attr_set = false;
    if (!attr_set) {
        PIML_CUDA(cudaFuncSetAttribute(pinnsf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
        attr_set = true;
    }
    const int64_t tiles = a.n_ped_tiles + a.n_obs_tiles;
    PIML_REQUIRE(tiles < (1LL << 31), "piml_pinnsf_forward_f32: too many tiles");
    pinnsf_tile_kernel<<<static_cast<unsigned>(tiles), FT_THREADS, smem, st>>>(P, T, a);
    count_launch();
    rc = check_launch("pinnsf_tile_kernel");
    if (rc) return rc;
    const int threads = 256;
    pinnsf_finish_kernel<<<static_cast<unsigned>((2 * R + threads - 1) / threads), threads, 0, st>>>(
        sums, self, dnorm, R, has_obs, tau, acc);
    count_launch();
    return check_launch("pinnsf_finish_kernel");
}
