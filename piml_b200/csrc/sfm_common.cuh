// sfm_common.cuh -- the arithmetic of the pure social-force model, shared by the per-step kernel (sfm.cu) and the
// persistent rollout kernel (rollout_sfm.cu) so that both produce bit-identical accelerations.
#pragma once
#include "common.cuh"

namespace piml {

// v0 repulsion of one slot: -A exp(B (r + eps)) dr / (r + eps)            (utils.py:53-59)
__device__ __forceinline__ float2 sfm_v0(float dx, float dy, float A, float B, float eps) {
    const float r = __fadd_rn(norm2_rn(dx, dy), eps);                          // r += eps        (utils.py:56)
    const float a = __fmul_rn(A, expf(__fmul_rn(B, r)));                       // A*exp(B*r)      (:57)
    return make_float2(__fmul_rn(-a, __fdiv_rn(dx, r)), __fmul_rn(-a, __fdiv_rn(dy, r)));   // -acc * dr/r   (:58-59)
}

// acc = (sum ped messages + sum obs messages) + (v0 * dest_dir - v) / tau   (model.py:1205-1212)
__device__ __forceinline__ float2 sfm_total(float ax, float ay, float ox, float oy, float dfx, float dfy, float vx,
                                            float vy, float v0, float tau) {
    float n = norm2_rn(dfx, dfy);
    if (n == 0.f) n = 0.1f;                                                    // temp_[temp_ == 0] += 0.1   (:1208)
    const float dxs = __fdiv_rn(__fsub_rn(__fmul_rn(v0, __fdiv_rn(dfx, n)), vx), tau);
    const float dys = __fdiv_rn(__fsub_rn(__fmul_rn(v0, __fdiv_rn(dfy, n)), vy), tau);
    return make_float2(__fadd_rn(__fadd_rn(ax, ox), dxs), __fadd_rn(__fadd_rn(ay, oy), dys));
}

}  // namespace piml
