// api.cu -- library-wide C ABI helpers (version, errors, launch accounting, device info).
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace piml {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PIML_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return PIML_OK;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;          // B200; not cached so a later call can still succeed
    }
    return cached;
}

// ---- pipe-throughput probes: measured denominators for the FP32 / MUFU roofline of the all-pairs kernels ----------
__global__ void __launch_bounds__(256) fp32_probe_kernel(float *out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = static_cast<float>(threadIdx.x + q) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = fmaf(x[q], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += x[q];
    out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) mufu_probe_kernel(float *out, int iters) {
    float x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = static_cast<float>(threadIdx.x + q) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int q = 0; q < 8; ++q) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[q]) : "f"(-x[q]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += x[q];
    out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}


// Packed FP32 (Blackwell fma.rn.f32x2 -> SASS FFMA2): 8 independent 64-bit chains, 2 FMAs per instruction.
__global__ void __launch_bounds__(256) fp32x2_probe_kernel(float *out, int iters, float a, float b) {
    unsigned long long x[8], av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float f = static_cast<float>(threadIdx.x + q) * 1e-3f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(x[q]) : "f"(f));
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int q = 0; q < 8; ++q) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[q]) : "l"(av), "l"(bv));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[q]));
        s += lo + hi;
    }
    out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

// Co-issue probe: per unit 8 FFMA2 (16 FMAs) + NM MUFU.EX2 + NA FSEL-class ALU ops on independent chains.  Tells
// whether MUFU / ALU-pipe work hides behind packed FP32 work (time == max of the pipes) or adds to it.
template <int NM, int NA>
__global__ void __launch_bounds__(256) mix_probe_kernel(float *out, int iters, float a, float b) {
    unsigned long long x[8], av, bv;
    float m[4], c[4];
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float f = static_cast<float>(threadIdx.x + q) * 1e-3f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(x[q]) : "f"(f));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { m[q] = static_cast<float>(threadIdx.x + q) * 1e-3f; c[q] = m[q] + 1.f; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int q = 0; q < 8; ++q) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[q]) : "l"(av), "l"(bv));
#pragma unroll
            for (int q = 0; q < NM; ++q) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(m[q & 3]) : "f"(-m[q & 3]));
#pragma unroll
            for (int q = 0; q < NA; ++q)
                asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %1, %0, p;}" : "+f"(c[q & 3]) : "f"(a));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[q]));
        s += lo + hi;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) s += m[q] + c[q];
    out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

// Operand-pattern probes for the packed FP32 pipe: 8 chains whose source operands are all distinct registers (no
// loop-invariant operand for the reuse cache).  MODE 0: x=fma2(x,y,z)  1: x=mul2(x,y)  2: x=add2(x,y)
// 3: x=fma2(y,y,x)  4: scalar x=fma(x,y,z)  5: x=fma2(y_q, z_{q+1}, x)  (accumulate form)
template <int MODE>
__global__ void __launch_bounds__(256) operand_probe_kernel(float *out, int iters) {
    unsigned long long x[8], y[8], z[8];
    float xs[8], ys[8], zs[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float f = static_cast<float>(threadIdx.x + q) * 1e-3f;
        xs[q] = f; ys[q] = 0.999f + f * 1e-6f; zs[q] = 1e-3f + f * 1e-6f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x[q]) : "f"(xs[q]), "f"(xs[q] + 0.5f));
        asm("mov.b64 %0, {%1, %2};" : "=l"(y[q]) : "f"(ys[q]), "f"(ys[q] - 1e-4f));
        asm("mov.b64 %0, {%1, %2};" : "=l"(z[q]) : "f"(zs[q]), "f"(zs[q] + 1e-4f));
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[q]) : "l"(y[q]), "l"(z[q]));
                if (MODE == 1) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[q]) : "l"(y[q]));
                if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[q]) : "l"(z[q]));
                if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(x[q]) : "l"(z[q]));
                if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(xs[q]) : "f"(ys[q]), "f"(zs[q]));
                if (MODE == 5) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x[q]) : "l"(y[q]), "l"(z[(q + 1) & 7]));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[q]));
        s += lo + hi + xs[q];
    }
    out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

}  // namespace piml

extern "C" int piml_pipe_probe(int which, int ctas, int iters, float *out, void *stream) {
    PIML_REQUIRE(out && ctas > 0 && iters > 0, "piml_pipe_probe: bad arguments");
    PIML_REQUIRE(which >= 0 && which <= 12, "piml_pipe_probe: which must be 0..12");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (which) {
        case 0: piml::fp32_probe_kernel<<<ctas, 256, 0, st>>>(out, iters, 0.999f, 1e-3f); break;
        case 1: piml::mufu_probe_kernel<<<ctas, 256, 0, st>>>(out, iters); break;
        case 2: piml::fp32x2_probe_kernel<<<ctas, 256, 0, st>>>(out, iters, 0.999f, 1e-3f); break;
        case 3: piml::mix_probe_kernel<2, 0><<<ctas, 256, 0, st>>>(out, iters, 0.999f, 1e-3f); break;
        case 4: piml::mix_probe_kernel<4, 0><<<ctas, 256, 0, st>>>(out, iters, 0.999f, 1e-3f); break;
        case 5: piml::mix_probe_kernel<0, 4><<<ctas, 256, 0, st>>>(out, iters, 0.999f, 1e-3f); break;
        case 6: piml::mix_probe_kernel<4, 4><<<ctas, 256, 0, st>>>(out, iters, 0.999f, 1e-3f); break;
        case 7: piml::operand_probe_kernel<0><<<ctas, 256, 0, st>>>(out, iters); break;
        case 8: piml::operand_probe_kernel<1><<<ctas, 256, 0, st>>>(out, iters); break;
        case 9: piml::operand_probe_kernel<2><<<ctas, 256, 0, st>>>(out, iters); break;
        case 10: piml::operand_probe_kernel<3><<<ctas, 256, 0, st>>>(out, iters); break;
        case 11: piml::operand_probe_kernel<4><<<ctas, 256, 0, st>>>(out, iters); break;
        default: piml::operand_probe_kernel<5><<<ctas, 256, 0, st>>>(out, iters); break;
    }
    piml::count_launch();
    return piml::check_launch("pipe_probe_kernel");
}

extern "C" int piml_version(void) { return 100; }

extern "C" const char *piml_last_error(void) { return piml::g_err; }

extern "C" int64_t piml_launch_count(void) { return piml::g_launches.load(std::memory_order_relaxed); }

extern "C" int piml_device_info(int *sm_count, int *cc) {
    int dev = 0, n = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -piml::fail(PIML_ERR_CUDA, "no CUDA device");
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (sm_count) *sm_count = n;
    if (cc) *cc = major * 10 + minor;
    return 0;
}
