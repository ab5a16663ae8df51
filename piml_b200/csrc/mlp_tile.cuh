// mlp_tile.cuh -- shared machinery of the fused interaction-network kernels (forward: mlp.cu, backward: mlp_bwd.cu):
// the layer plan, the TMA weight-chunk pipeline and the register-tiled dense layer on transposed shared-memory tiles.
#pragma once
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

struct FLayer { int K, OUT, OUTP, NJ, w_off, b_off; int64_t t_off; };   // w/b offsets in floats relative to the branch /
                                                                        // head base; t_off: offset in the torch-layout vector
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
    int64_t t_branch_off[2], t_coll_off, t_total;   // the same in the torch-layout vector (pack_state_dict order)
};

struct FTab { int n[2]; FChunk c[2][FT_MAXCH]; };

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

// Ys[o][r] = act(scale * (b[o] + sum_k W[k][o] Xs[k][r])) (+ addsrc[o][r]) (* drop) for r < nrows.  All threads call it.
// addsrc may alias Ys (each element is read and written by the same thread).
template <int NJ>
__device__ __forceinline__ void dense(WPipe &wp, const FLayer &L, const float *__restrict__ pbase, const float *Xs,
                                      float *Ys, int nrows, bool relu, float scale, const float *addsrc,
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
        if (addsrc) {                                              // ResBlock: lin(x) + x   (model.py:78-79)
#pragma unroll
            for (int q = 0; q < 8; ++q) y[q] += addsrc[o * FT_TRP + tr * 8 + q];
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

static __device__ __noinline__ void dense_any(WPipe &wp, const FLayer &L, const float *pbase, const float *Xs, float *Ys,
                                          int nrows, bool relu, float scale, const float *addsrc = nullptr,
                                          const float *drop = nullptr, int drop_ld = 0) {
    switch (L.NJ) {
        case 8: dense<8>(wp, L, pbase, Xs, Ys, nrows, relu, scale, addsrc, drop, drop_ld); break;
        case 4: dense<4>(wp, L, pbase, Xs, Ys, nrows, relu, scale, addsrc, drop, drop_ld); break;
        case 2: dense<2>(wp, L, pbase, Xs, Ys, nrows, relu, scale, addsrc, drop, drop_ld); break;
        default: dense<1>(wp, L, pbase, Xs, Ys, nrows, relu, scale, addsrc, drop, drop_ld); break;
    }
}

// Elementwise visit of a transposed [width][nrows] shared-memory tile, f(i, r).  A warp covers 8 features x 4 rows per
// step: the 32 shared-memory addresses i*FT_TRP + r fall into 32 distinct banks, and a row-major global access
// [(row0 + r) * width + i] touches four full 32-byte sectors.
template <class F>
__device__ __forceinline__ void tile_pass(int width, int nrows, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fi = lane & 7, ri = lane >> 3;
    const int nfb = (width + 7) >> 3, nrb = (nrows + 3) >> 2;
    for (int b = warp; b < nfb * nrb; b += FT_THREADS / 32) {
        const int i = (b % nfb) * 8 + fi, r = (b / nfb) * 4 + ri;
        if (i < width && r < nrows) f(i, r);
    }
}

// dst[(row0 + r) * width + i] = tile[i][r]
__device__ __forceinline__ void store_tile(const float *tile, float *dst, int64_t row0, int nrows, int width) {
    tile_pass(width, nrows, [&](int i, int r) { dst[(row0 + r) * width + i] = tile[i * FT_TRP + r]; });
}

// Float offsets of the per-layer row-major matrices kept between the training-mode forward and the backward:
// the STASH holds every Linear's output after its activation (the next Linear's input); the G workspace of the
// backward holds the gradient w.r.t. every Linear's pre-activation output, with the same shapes.
struct SPlan {
    int64_t enc[2][8];   // encoder layer l of branch br: (rows_br, width_l); the last one after the 2x / dropout fold
    int64_t dec[2][8];   // decoder layer l: kind 0 (rows_br, width_l), kind 1 (R, width_l)
    int64_t proc_h[2];   // single-block processor only: relu(W e + b) (rows_br, pw)
    int64_t proc[2];     // single-block processor only: its output (relu(..) + e) * dropout (rows_br, pw)
    int64_t sum[2];      // kind 1 only: per-agent sum of the slot embeddings (R, pw)
    int64_t pred[2];     // G only: predictor output gradient, kind 0 (rows_br, 2), kind 1 (R, 2)
    int64_t collh;       // collision head hidden layer (rows_ped, hidden)
    int64_t prob;        // stash: sigmoid output (rows_ped); G: gradient of the logit (rows_ped)
    int64_t total;
};

static SPlan make_splan(const FPlan &P, bool has_obs, int64_t R, int kp, int ko, bool want_coll) {
    SPlan S;
    int64_t off = 0;
    auto take = [&](int64_t rows, int width) { const int64_t o = off; off += ((rows * width + 3) / 4) * 4; return o; };
    for (int br = 0; br < 2; ++br) {
        const int64_t rows = (br == 0) ? R * kp : (has_obs ? R * ko : 0);
        for (int l = 0; l < 8; ++l) {
            S.enc[br][l] = l < P.n_enc ? take(rows, P.enc[l].OUT) : 0;
            S.dec[br][l] = 0;
        }
        S.proc_h[br] = P.proc_mode == 1 ? take(rows, P.pw) : 0;
        S.proc[br] = P.proc_mode == 1 ? take(rows, P.pw) : 0;
        for (int l = 0; l < P.n_dec; ++l) S.dec[br][l] = take(P.kind == 0 ? rows : (rows ? R : 0), P.dec[l].OUT);
        S.sum[br] = P.kind == 1 ? take(rows ? R : 0, P.pw) : 0;
        S.pred[br] = take(P.kind == 0 ? rows : (rows ? R : 0), 2);
    }
    const bool coll = want_coll && P.n_coll > 0;
    S.collh = (coll && P.n_coll > 1) ? take(R * kp, P.coll[0].OUT) : 0;
    S.prob = coll ? take(R * kp, 1) : 0;
    S.total = off;
    return S;
}

// ---- parameter packing: torch order (per Linear: W (out,in) row-major, b) -> device layout ------------------------
// dst: Wp [K][OUTP] then bias [OUTP];  Wp[kk][perm(o)] = src[src_w + o*so + kk*sk];  bias[o] = src[src_b + o] (src_b < 0: 0)
struct PackRec { int K, OUT, OUTP, NJ; int64_t src_w, src_b, dst; int so, sk; };
struct PackTab { int n; PackRec r[48]; };

static __global__ void pinnsf_pack_kernel(const __grid_constant__ PackTab T, const float *__restrict__ src,
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
        if (o < L.OUT) v = src[L.src_w + static_cast<int64_t>(o) * L.so + static_cast<int64_t>(kk) * L.sk];
    } else {
        const int o = static_cast<int>(e - static_cast<int64_t>(L.K) * L.OUTP);
        if (o < L.OUT && L.src_b >= 0) v = src[L.src_b + o];
    }
    dst[i] = v;
}

static int nj_for(int out) { return out > 64 ? 8 : (out > 32 ? 4 : (out > 16 ? 2 : 1)); }

// Builds the layer plan and the packing table.  transposed = false: the forward layout (W^T, K = in rows of OUT = out
// columns, bias).  transposed = true: the backward layout for dX = dY W: per Linear K = out rows of OUT = in columns
// (torch's own (out,in) matrix, columns permuted for the tile kernel) and a ZERO bias.
static int build_plan(const piml_net_desc *d, FPlan *P, PackTab *PT, bool transposed = false) {
    PIML_REQUIRE(d->n_enc >= 1 && d->n_enc <= 8 && d->n_dec >= 1 && d->n_dec <= 8 && d->n_coll >= 0 && d->n_coll <= 2,
                 "piml_pinnsf: unsupported layer counts (enc %d, dec %d, coll %d)", d->n_enc, d->n_dec, d->n_coll);
    PIML_REQUIRE(d->kind == 0 || d->kind == 1, "piml_pinnsf: kind must be 0 or 1");
    PIML_REQUIRE(d->proc_mode == 0 || d->proc_mode == 1, "piml_pinnsf: proc_mode must be 0 or 1");
    PIML_REQUIRE(d->enc_dims[0] == 6, "piml_pinnsf: feature dim must be 6, got %d", d->enc_dims[0]);
    int maxw = 6;
    int64_t src = 0;
    int dst = 0;
    PT->n = 0;
    int64_t tbase = 0;
    auto rec = [&](int in, int out, int base, bool record) {
        FLayer L;
        L.K = transposed ? out : in; L.OUT = transposed ? in : out; L.NJ = nj_for(L.OUT); L.OUTP = 16 * L.NJ;
        L.w_off = dst - base; L.b_off = L.w_off + L.K * L.OUTP;
        L.t_off = src - tbase;
        if (record) {
            PackRec r{L.K, L.OUT, L.OUTP, L.NJ, src, transposed ? -1 : src + static_cast<int64_t>(in) * out, dst,
                      transposed ? 1 : in, transposed ? in : 1};
            PT->r[PT->n++] = r;
        }
        src += static_cast<int64_t>(in) * out + out;
        dst += L.K * L.OUTP + L.OUTP;
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
        tbase = src;
        P->t_branch_off[br] = src;
        for (int l = 0; l < d->n_enc; ++l) P->enc[l] = rec(d->enc_dims[l], d->enc_dims[l + 1], base, true);
        if (d->proc_mode == 1) P->proc = rec(P->pw, P->pw, base, true);
        for (int l = 0; l < d->n_dec; ++l) P->dec[l] = rec(d->dec_dims[l], d->dec_dims[l + 1], base, true);
        P->pred = rec(P->dw, 2, base, true);
    }
    P->coll_off = dst;
    tbase = src;
    P->t_coll_off = src;
    if (d->n_coll) {
        PIML_REQUIRE(d->coll_dims[d->n_coll] == 1, "piml_pinnsf: collision head must end in width 1");
        PIML_REQUIRE(d->coll_dims[0] == (d->kind == 0 ? P->dw : P->pw),
                     "piml_pinnsf: collision head input width %d does not match", d->coll_dims[0]);
        const int base = dst;
        for (int l = 0; l < d->n_coll; ++l) P->coll[l] = rec(d->coll_dims[l], d->coll_dims[l + 1], base, true);
    }
    P->total = dst;
    P->t_total = src;
    PIML_REQUIRE(maxw <= FT_MAXW, "piml_pinnsf: layer width %d > %d is not supported", maxw, FT_MAXW);
    return PIML_OK;
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
