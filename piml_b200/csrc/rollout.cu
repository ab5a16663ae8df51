// rollout.cu -- the whole inference rollout loop of BaseSimulator.get_multiple_rollouts (reference
// src/models/simulators.py:595-652) enqueued by ONE C call: per step the fused network forward (tensor cores when the
// network allows it), the fused state update (record / Euler / arrival / entry) and the fused feature rebuild.
// No host synchronisation and no Python between steps: a GC-sized scene is launch-latency bound, and the launches of
// step t+1 are queued while step t runs.
#include "common.cuh"

namespace piml {
bool sfm_rollout_fits(const piml_rollout_args *r);                 // rollout_sfm.cu
int sfm_rollout_launch(const piml_rollout_args *r, cudaStream_t st);
}  // namespace piml

using namespace piml;

extern "C" int piml_rollout_f32(const piml_rollout_args *r, void *stream) {
    PIML_REQUIRE(r && (r->sfm || (r->desc && (r->packed || r->packed_tc))),
                 "piml_rollout_f32: null descriptor / parameters");
    PIML_REQUIRE(r->S >= 1 && r->N >= 1 && r->T >= 1 && r->D >= 1 && r->M >= 0 && r->t_start >= 0 && r->t_start < r->T,
                 "piml_rollout_f32: bad dimensions S=%d N=%d T=%d D=%d M=%d t_start=%d", r->S, r->N, r->T, r->D, r->M,
                 r->t_start);
    PIML_REQUIRE(r->pos_tm && r->vel_tm && r->acc_tm && r->dest_tm && r->dest_idx_tm && r->entry_tm && r->dest_num &&
                     r->waypoints && r->desired_speed,
                 "piml_rollout_f32: null ground-truth / static input");
    PIML_REQUIRE(r->p && r->v && r->a && r->dest && r->dest_idx && r->hist_v && r->a_next && r->ped_f && r->self_f &&
                     r->dest_f && (r->M == 0 || (r->obs_f && r->obstacles)),
                 "piml_rollout_f32: null state / feature buffer");
    PIML_REQUIRE(r->rec_p && r->rec_v && r->rec_a && r->rec_mask, "piml_rollout_f32: null output");
    const int64_t SN = static_cast<int64_t>(r->S) * r->N;
    const int kp = r->kp < r->N ? r->kp : r->N;
    const int ko = r->M > 0 ? (r->ko < r->M ? r->ko : r->M) : 0;
    const int has_obs = (r->has_obs && ko > 0) ? 1 : 0;
    // pure social-force model on GC-shaped scenes: the whole loop in one persistent kernel (rollout_sfm.cu)
    if (r->sfm && sfm_rollout_fits(r)) return sfm_rollout_launch(r, static_cast<cudaStream_t>(stream));
    for (int t = r->t_start; t < r->T; ++t) {
        int rc;
        // a_next = model(*state_features)[0]                                               (simulators.py:602)
        if (r->sfm)                            // pure social-force mode (BASELINE config 2)
            rc = piml_sfm_forward_f32(r->sfm, r->ped_f, has_obs ? r->obs_f : nullptr, r->self_f, SN, kp,
                                      has_obs ? ko : 0, r->a_next, nullptr, nullptr, stream);
        else if (r->packed_tc)
            rc = piml_pinnsf_forward_tc_f32(r->desc, r->packed_tc, has_obs, r->tau, r->ped_f, r->obs_f, r->self_f, SN,
                                            kp, ko, 0, r->a_next, nullptr, nullptr, stream);
        else
            rc = piml_pinnsf_forward_f32(r->desc, r->packed, has_obs, r->tau, r->ped_f, r->obs_f, r->self_f, SN, kp,
                                         ko, 0, nullptr, nullptr, r->a_next, nullptr, nullptr, nullptr, stream);
        if (rc) return rc;
        // record, Euler, arrival / waypoint switch, entry from the data at t+1, hist_v           (:596-639)
        const bool last = t >= r->T - 1;
        const int64_t o2 = static_cast<int64_t>(t + 1) * SN * 2, o1 = static_cast<int64_t>(t + 1) * SN;
        const int64_t q2 = static_cast<int64_t>(t) * SN * 2, q1 = static_cast<int64_t>(t) * SN;
        rc = piml_integrate_step_f32(r->p, r->v, r->a, r->a_next, r->dest, r->dest_idx, r->dest_num, r->waypoints,
                                     r->S, r->D, r->N, r->dt, 1, last ? nullptr : r->entry_tm + o1,
                                     last ? nullptr : r->pos_tm + o2, last ? nullptr : r->vel_tm + o2,
                                     last ? nullptr : r->acc_tm + o2, last ? nullptr : r->dest_tm + o2,
                                     last ? nullptr : r->dest_idx_tm + o1, r->hist_v, r->rec_p + q2, r->rec_v + q2,
                                     r->rec_a + q2, r->rec_mask + q1, stream);
        if (rc) return rc;
        // features of the new state + self_features = cat(dest_f, hist_v, a, desired_speed)      (:642-652)
        rc = piml_state_features_f32(r->p, r->v, r->a, r->dest, r->obstacles, r->obs_per_scene, r->S, r->N, r->M,
                                     r->kp, r->cos_p, r->thr_p, r->ko, r->cos_o, r->thr_o, r->hist_v,
                                     r->desired_speed, r->ped_f, r->obs_f, r->self_f, r->dest_f, stream);
        if (rc) return rc;
    }
    return PIML_OK;
}
