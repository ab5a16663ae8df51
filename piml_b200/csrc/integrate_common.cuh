// integrate_common.cuh -- one agent's state update of get_multiple_rollouts (simulators.py:603-626), shared by
// integrate_kernel and the persistent rollout kernel (rollout_sfm.cu).
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace piml {

struct AgentState { float2 p, v, a, dest; int64_t di; float2 hv; };

// Lagged explicit Euler with the OLD a and v, arrival / waypoint switch, removal on arrival.  `wp` points at
// waypoints[scene][0][n]; consecutive waypoints of the agent are `wp_stride` float2 apart.
__device__ __forceinline__ void integrate_update(AgentState &s, float2 a_next, float dt, int remove_on_arrival,
                                                 int64_t dest_num, const float2 *__restrict__ wp, int64_t wp_stride) {
    // v_next = v + a*dt ; p_next = p + v*dt                                   (simulators.py:603-604)
    const float2 vn = make_float2(__fadd_rn(s.v.x, __fmul_rn(s.a.x, dt)), __fadd_rn(s.v.y, __fmul_rn(s.a.y, dt)));
    float2 pn = make_float2(__fadd_rn(s.p.x, __fmul_rn(s.v.x, dt)), __fadd_rn(s.p.y, __fmul_rn(s.v.y, dt)));
    int64_t di = s.di;
    const float dis = norm2_rn(__fsub_rn(s.p.x, s.dest.x), __fsub_rn(s.p.y, s.dest.y));   // :608
    if (dis < 0.5f) di += 1;                                                   // :609
    if (di > dest_num - 1) {
        if (remove_on_arrival) pn = make_float2(CUDART_NAN_F, CUDART_NAN_F);   // :611
        di -= 1;                                                               // :613
    }
    s.dest = wp[di * wp_stride];                                               // :614-616
    s.p = pn; s.v = vn; s.a = a_next; s.di = di;
    s.hv = vn;                                                                 // :624-626
}

}  // namespace piml
