// mlp_tc.cu -- interaction-network forward on the 5th-generation tensor cores (tcgen05, TMEM) with 3xTF32 operand
// splitting, for sm_100a.   (work in progress: the single-layer self test)
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace piml {

// Self test of the tensor-core path: Y (128,N) = X (128,K) W^T, W (N,K) row-major like torch's Linear.weight, evaluated
// as 3 tf32 MMAs per K step (x_lo w_hi + x_hi w_lo + x_hi w_hi), A in TMEM, B in shared memory, D in TMEM.
__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const float *__restrict__ x, const float *__restrict__ w, int K,
                                                          int N, int terms, int lbo_o, int sbo_o, int cold, int colah, int colal, float *__restrict__ y,
                                                          float *__restrict__ dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *w_hi = reinterpret_cast<uint32_t *>(smem_raw);                  // [K/4][N][4]
    uint32_t *w_lo = w_hi + K * N;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int e = tid; e < N * K; e += 128) {
        const int n = e / K, k = e % K;
        uint32_t hi, lo;
        tc::split_tf32(w[e], hi, lo);
        const int cell = ((k >> 2) * N + n) * 4 + (k & 3);
        w_hi[cell] = hi; w_lo[cell] = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");              // generic-proxy writes -> tensor core reads
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t lane_base = tbase + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t COL_D = cold, COL_AH = colah, COL_AL = colal;
    // A: thread = row
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) tc::split_tf32(x[tid * K + k0 + q], hi[q], lo[q]);
        tc::st8(lane_base + COL_AH + k0, hi);
        tc::st8(lane_base + COL_AL + k0, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::idesc_tf32(N);
        const uint32_t lbo = lbo_o ? lbo_o : N * 16, sbo = sbo_o ? sbo_o : 128;
        const uint32_t kstep = 2 * N * 16;
        bool acc = false;
        for (int j = 0; j < K / 8; ++j) {
            const uint64_t bh = tc::smem_desc(tc::smem_addr(w_hi) + j * kstep, lbo, sbo);
            const uint64_t bl = tc::smem_desc(tc::smem_addr(w_lo) + j * kstep, lbo, sbo);
            if (terms >= 3) { tc::mma_tf32_ts(tbase + COL_D, tbase + COL_AL + j * 8, bh, idesc, acc); acc = true; }
            if (terms >= 2) { tc::mma_tf32_ts(tbase + COL_D, tbase + COL_AH + j * 8, bl, idesc, acc); acc = true; }
            tc::mma_tf32_ts(tbase + COL_D, tbase + COL_AH + j * 8, bh, idesc, acc);
            acc = true;
        }
        tc::commit(&bar);
    }
    const bool done = mbar_wait_bounded(&bar, 0, 1u << 22);
    tc::fence_after_sync();
    if (!done) {                                                              // never hang the box on a bad descriptor
        for (int n = 0; n < N; ++n) y[tid * N + n] = __int_as_float(0x7fc00000);
        __syncthreads();
        if (warp == 0) tc::tmem_dealloc(tbase, 512);
        return;
    }
    for (int n0 = 0; n0 < N; n0 += 32) {
        uint32_t r[32];
        tc::ld32(lane_base + COL_D + n0, r);
        tc::wait_ld();
#pragma unroll
        for (int q = 0; q < 32; ++q)
            if (n0 + q < N) y[tid * N + n0 + q] = __uint_as_float(r[q]);
    }
    if (dbg)
        for (int k0 = 0; k0 < K; k0 += 32) {                                  // what does the A region hold afterwards?
            uint32_t r[32];
            tc::ld32(lane_base + COL_AH + k0, r);
            tc::wait_ld();
#pragma unroll
            for (int q = 0; q < 32; ++q)
                if (k0 + q < K) dbg[tid * K + k0 + q] = __uint_as_float(r[q]);
        }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

}  // namespace piml

using namespace piml;

extern "C" int piml_tc_selftest_f32(const float *x, const float *w, int K, int N, int terms, float *y, void *stream) {
    int lbo_o = 0, sbo_o = 0;
    if (const char *e = getenv("PIML_TC_LBO")) lbo_o = atoi(e);
    if (const char *e = getenv("PIML_TC_SBO")) sbo_o = atoi(e);
    int cold = 0, colah = 128, colal = 256;
    if (const char *e = getenv("PIML_TC_COLD")) cold = atoi(e);
    if (const char *e = getenv("PIML_TC_COLAH")) colah = atoi(e);
    if (const char *e = getenv("PIML_TC_COLAL")) colal = atoi(e);
    float *dbg = nullptr;
    if (getenv("PIML_TC_DBG")) dbg = y + 128 * N;                             // caller allocated 128*(N+K) floats
    PIML_REQUIRE(x && w && y, "piml_tc_selftest_f32: null pointer");
    PIML_REQUIRE(K >= 8 && K <= 128 && K % 8 == 0 && N >= 16 && N <= 128 && N % 16 == 0,
                 "piml_tc_selftest_f32: need K in [8,128] multiple of 8 and N in [16,128] multiple of 16");
    PIML_REQUIRE(terms >= 1 && terms <= 3, "piml_tc_selftest_f32: terms must be 1 (plain tf32), 2 or 3 (3xTF32)");
    const size_t smem = sizeof(uint32_t) * 2 * K * N;
    PIML_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    tc_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(x, w, K, N, terms, lbo_o, sbo_o, cold, colah, colal, y, dbg);
    count_launch();
    return check_launch("tc_probe_kernel");
}
