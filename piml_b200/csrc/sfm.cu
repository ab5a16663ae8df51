// sfm.cu -- social-force pairwise repulsion per neighbour slot for sm_100a.
// Replaces UTILS.calc_acceleration (reference src/utils/utils.py:31-100).
#include "common.cuh"

namespace piml {

// v1/v2 use the relative POSITION as dv (reference quirk utils.py:67,84), so cos = (dr.dr)/(r+eps)/(r+eps).
__global__ void calc_acceleration_kernel(const float *__restrict__ rel, int64_t S, int stride, int version, float A,
                                         float B, float C, float D, float ct, float st, float eps,
                                         float2 *__restrict__ out) {
    const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float dx = rel[s * stride], dy = rel[s * stride + 1];
    const float r = __fadd_rn(norm2_rn(dx, dy), eps);                          // r += eps        (:56)
    const float nx = __fdiv_rn(dx, r), ny = __fdiv_rn(dy, r);                  // dir = dr / r    (:58)
    float a;
    if (version == 0) {
        a = __fmul_rn(A, expf(__fmul_rn(B, r)));                               // A*exp(B*r)      (:57)
        out[s] = make_float2(__fmul_rn(-a, nx), __fmul_rn(-a, ny));
        return;
    }
    // cos = sum(dr*dv)/r/v with dv == dr and v == r                           (:73,:90)
    const float c = __fdiv_rn(__fdiv_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), r), r);
    if (version == 1) {
        a = __fmul_rn(A, expf(__fadd_rn(__fmul_rn(B, r), __fmul_rn(C, c))));   // (:74)
        out[s] = make_float2(__fmul_rn(-a, nx), __fmul_rn(-a, ny));
    } else {
        a = __fmul_rn(A, expf(__fadd_rn(__fadd_rn(__fmul_rn(B, r), __fmul_rn(C, c)),
                                        __fmul_rn(__fmul_rn(D, r), c))));      // (:91)
        const float bx = __fadd_rn(__fmul_rn(ct, nx), __fmul_rn(-st, ny));     // rotate by theta (:93-99)
        const float by = __fadd_rn(__fmul_rn(st, nx), __fmul_rn(ct, ny));
        out[s] = make_float2(__fmul_rn(-a, bx), __fmul_rn(-a, by));
    }
}

}  // namespace piml

using namespace piml;

extern "C" int piml_calc_acceleration_f32(const float *rel, int64_t S, int stride, int version, float A, float B,
                                          float C, float D, float theta, float eps, float *out, void *stream) {
    PIML_REQUIRE(rel && out, "piml_calc_acceleration_f32: null pointer");
    PIML_REQUIRE(S >= 0 && stride >= 2, "piml_calc_acceleration_f32: bad size/stride");
    PIML_REQUIRE(version >= 0 && version <= 2, "piml_calc_acceleration_f32: equation version %d unknown", version);
    if (S == 0) return PIML_OK;
    const float ct = static_cast<float>(cos(static_cast<double>(theta)));
    const float st = static_cast<float>(sin(static_cast<double>(theta)));
    const int threads = 256;
    calc_acceleration_kernel<<<static_cast<unsigned>((S + threads - 1) / threads), threads, 0,
                               static_cast<cudaStream_t>(stream)>>>(rel, S, stride, version, A, B, C, D, ct, st, eps,
                                                                    reinterpret_cast<float2 *>(out));
    count_launch();
    return check_launch("calc_acceleration_kernel");
}
