// rollout.cu -- the whole inference rollout loop of BaseSimulator.get_multiple_rollouts (reference
// src/models/simulators.py:595-652) enqueued by ONE C call: per step the fused network forward (tensor cores when the
// network allows it), the fused state update (record / Euler / arrival / entry) and the fused feature rebuild.
// No host synchronisation and no Python between steps: a GC-sized scene is launch-latency bound, and the launches of
// step t+1 are queued while step t runs.
#include "common.cuh"

#include <stdlib.h>

namespace piml {
bool sfm_rollout_fits(const piml_rollout_args *r);                 // rollout_sfm.cu
int sfm_rollout_launch(const piml_rollout_args *r, cudaStream_t st);
int integrate_step_indirect(const piml_rollout_args *r, const int *t_dev, cudaStream_t st);   // integrate.cu
int advance_counter(int *t_dev, cudaStream_t st);
struct ShardInfo;
int nn_step_launch(const piml_nn_step_args *r, const int *t_dev, cudaStream_t st, const ShardInfo *sh);   // nn_step.cu

// frame counter of the captured loop + a capture-capable stream (the legacy default stream cannot be captured): one
// per calling thread and device
static int frame_counter(int **out, cudaStream_t *side, cudaEvent_t *ev) {
    struct Slot { int dev; int *p; cudaStream_t st; cudaEvent_t ev; };
    static thread_local Slot slots[16] = {};
    static thread_local int used = 0;
    int dev = 0;
    PIML_CUDA(cudaGetDevice(&dev));
    for (int i = 0; i < used; ++i)
        if (slots[i].dev == dev) { *out = slots[i].p; *side = slots[i].st; *ev = slots[i].ev; return PIML_OK; }
    PIML_REQUIRE(used < 16, "piml_rollout_f32: too many devices in one thread");
    Slot s{dev, nullptr, nullptr, nullptr};
    PIML_CUDA(cudaMalloc(&s.p, sizeof(int)));
    PIML_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    PIML_CUDA(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
    slots[used++] = s;
    *out = s.p; *side = s.st; *ev = s.ev;
    return PIML_OK;
}
}  // namespace piml

using namespace piml;

// The fused step (nn_step.cu) on the rollout's buffers, for frame t: features of the current state -> model -> record /
// Euler / arrival / entry from frame t + 1.  t_dev != nullptr: time-major BASE pointers, frame index in device memory.
static int fused_step(const piml_rollout_args *r, int t, const int *t_dev, void *stream) {
    piml_nn_step_args n = {};
    n.desc = r->desc; n.packed_tc = r->packed_tc; n.has_obs = r->has_obs; n.tau = r->tau;
    n.S = r->S; n.N = r->N; n.M = r->M; n.D = r->D; n.dt = r->dt; n.remove_on_arrival = 1;
    n.kp = r->kp; n.cos_p = r->cos_p; n.thr_p = r->thr_p; n.ko = r->ko; n.cos_o = r->cos_o; n.thr_o = r->thr_o;
    n.obstacles = r->obstacles; n.obs_per_scene = r->obs_per_scene;
    n.dest_num = r->dest_num; n.waypoints = r->waypoints; n.desired_speed = r->desired_speed;
    n.p = r->p; n.v = r->v; n.a = r->a; n.dest = r->dest; n.dest_idx = r->dest_idx; n.hist_v = r->hist_v;
    const int64_t SN = static_cast<int64_t>(r->S) * r->N;
    const bool last = !t_dev && t >= r->T - 1;
    const int64_t o2 = t_dev ? 0 : static_cast<int64_t>(t + 1) * SN * 2, o1 = t_dev ? 0 : static_cast<int64_t>(t + 1) * SN;
    const int64_t q2 = t_dev ? 0 : static_cast<int64_t>(t) * SN * 2, q1 = t_dev ? 0 : static_cast<int64_t>(t) * SN;
    if (!last) {
        n.entry = r->entry_tm + o1; n.p_gt = r->pos_tm + o2; n.v_gt = r->vel_tm + o2; n.a_gt = r->acc_tm + o2;
        n.dest_gt = r->dest_tm + o2; n.dest_idx_gt = r->dest_idx_tm + o1;
    }
    n.rec_p = r->rec_p + q2; n.rec_v = r->rec_v + q2; n.rec_a = r->rec_a + q2; n.rec_mask = r->rec_mask + q1;
    n.a_next = r->a_next;
    return nn_step_launch(&n, t_dev, static_cast<cudaStream_t>(stream), nullptr);
}

// One step of the loop: a_next = model(features); record / Euler / arrival / entry; features of the new state.
// t_dev != nullptr: the frame index comes from device memory (constant launch arguments: graph capture).
// features_after = false: the caller continues with fused steps, which rebuild the features themselves.
static int rollout_step(const piml_rollout_args *r, int t, const int *t_dev, int64_t SN, int kp, int ko, int has_obs,
                        void *stream, bool features_after = true) {
    int rc;
    // a_next = model(*state_features)[0]                                               (simulators.py:602)
    if (r->sfm)                                // pure social-force mode (BASELINE config 2)
        rc = piml_sfm_forward_f32(r->sfm, r->ped_f, has_obs ? r->obs_f : nullptr, r->self_f, SN, kp, has_obs ? ko : 0,
                                  r->a_next, nullptr, nullptr, stream);
    else if (r->packed_tc)
        rc = piml_pinnsf_forward_tc_f32(r->desc, r->packed_tc, has_obs, r->tau, r->ped_f, r->obs_f, r->self_f, SN, kp,
                                        ko, 0, r->a_next, nullptr, nullptr, stream);
    else
        rc = piml_pinnsf_forward_f32(r->desc, r->packed, has_obs, r->tau, r->ped_f, r->obs_f, r->self_f, SN, kp, ko, 0,
                                     nullptr, nullptr, r->a_next, nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    // record, Euler, arrival / waypoint switch, entry from the data at t+1, hist_v           (:596-639)
    if (t_dev) {
        rc = integrate_step_indirect(r, t_dev, static_cast<cudaStream_t>(stream));
    } else {
        const bool last = t >= r->T - 1;
        const int64_t o2 = static_cast<int64_t>(t + 1) * SN * 2, o1 = static_cast<int64_t>(t + 1) * SN;
        const int64_t q2 = static_cast<int64_t>(t) * SN * 2, q1 = static_cast<int64_t>(t) * SN;
        rc = piml_integrate_step_f32(r->p, r->v, r->a, r->a_next, r->dest, r->dest_idx, r->dest_num, r->waypoints,
                                     r->S, r->D, r->N, r->dt, 1, last ? nullptr : r->entry_tm + o1,
                                     last ? nullptr : r->pos_tm + o2, last ? nullptr : r->vel_tm + o2,
                                     last ? nullptr : r->acc_tm + o2, last ? nullptr : r->dest_tm + o2,
                                     last ? nullptr : r->dest_idx_tm + o1, r->hist_v, r->rec_p + q2, r->rec_v + q2,
                                     r->rec_a + q2, r->rec_mask + q1, stream);
    }
    if (rc || !features_after) return rc;
    // features of the new state + self_features = cat(dest_f, hist_v, a, desired_speed)      (:642-652)
    return piml_state_features_f32(r->p, r->v, r->a, r->dest, r->obstacles, r->obs_per_scene, r->S, r->N, r->M, r->kp,
                                   r->cos_p, r->thr_p, r->ko, r->cos_o, r->thr_o, r->hist_v, r->desired_speed,
                                   r->ped_f, r->obs_f, r->self_f, r->dest_f, stream);
}

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
    // Frames t_start + 1 .. T - 2 are identical launches once the frame index lives in device memory: capture ONE step
    // into a CUDA graph and replay it (a GC-sized scene is launch-latency bound: ~15 launches per step).  The first
    // step runs eagerly (it sizes the library's scratch buffers, which must not happen during capture), the last
    // one too (no successor frame to take entries from).  PIML_ROLLOUT_GRAPH=0 keeps the eager loop.
    // Large batches (many scenes, or one big crowd) with a network the fused step runs: from the second frame on a step
    // is piml_nn_step_f32's chain -- cell-list features in compact form -> tensor cores -> slot sums + state update
    // (nn_step.cu), the same operations in the same order with the loop boundary moved in front of the feature
    // rebuild: the first frame uses the caller's features, and the rebuild of the LAST iteration (simulators.py:642-652,
    // whose only effect is the in-place NaN -> 0 on v and a and the feature buffers) runs after the loop.
    // PIML_ROLLOUT_FUSED=0 / 1 forces the choice.
    const float thr_max = fmaxf(r->thr_p, r->M > 0 ? r->thr_o : 0.f);
    const char *fe = getenv("PIML_ROLLOUT_FUSED");
    const bool fused_ok = !r->sfm && r->packed_tc && piml_nn_step_supported(r->desc) && thr_max > 0.f && thr_max < 1e18f &&
                          kp >= 1 && r->T - r->t_start >= 3;
    const bool fused = fused_ok && (fe ? atoi(fe) != 0 : SN >= 16384);
    cudaStream_t user = static_cast<cudaStream_t>(stream), st = user;
    const char *ge = getenv("PIML_ROLLOUT_GRAPH");
    const int replays = r->T - 2 - r->t_start - (fused ? 1 : 0);   // frames t_start + 1 (fused: + 2) .. T - 2
    const bool use_graph = !(ge && atoi(ge) == 0) && replays >= 8;
    int *t_dev = nullptr;
    cudaEvent_t ev = nullptr;
    if (use_graph) {
        cudaStream_t side = nullptr;
        int rc0 = frame_counter(&t_dev, &side, &ev);
        if (rc0) return rc0;
        if (user == nullptr || user == cudaStreamLegacy) {         // the legacy default stream cannot be captured:
            PIML_CUDA(cudaEventRecord(ev, user));                  // run the loop on a side stream ordered after it
            PIML_CUDA(cudaStreamWaitEvent(side, ev, 0));
            st = side;
        }
    }
    stream = st;
    int t = r->t_start;
    int rc = rollout_step(r, t, nullptr, SN, kp, ko, has_obs, stream, !fused);
    if (rc) return rc;
    ++t;
    if (fused) {                                                   // eager: sizes the fused step's scratch buffers
        rc = fused_step(r, t, nullptr, stream);
        if (rc) return rc;
        ++t;
    }
    if (use_graph) {
        PIML_CUDA(cudaMemcpyAsync(t_dev, &t, sizeof(int), cudaMemcpyHostToDevice, st));
        PIML_CUDA(cudaStreamSynchronize(st));                      // `t` is a stack variable
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const int64_t c0 = piml_launch_count();
        PIML_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        rc = fused ? fused_step(r, -1, t_dev, stream) : rollout_step(r, -1, t_dev, SN, kp, ko, has_obs, stream);
        if (!rc) rc = advance_counter(t_dev, st);
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        const int64_t per_step = piml_launch_count() - c0;         // kernels of one captured step
        if (rc || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            if (rc) return rc;
            return piml::fail(PIML_ERR_CUDA, "piml_rollout_f32: stream capture failed: %s", cudaGetErrorString(ce));
        }
        PIML_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        for (int i = 0; i < replays; ++i) {
            const cudaError_t le = cudaGraphLaunch(exec, st);
            if (le != cudaSuccess) {
                cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
                return piml::fail(PIML_ERR_CUDA, "piml_rollout_f32: graph launch failed: %s", cudaGetErrorString(le));
            }
        }
        count_launch(static_cast<int>((replays - 1) * per_step)); // the capture counted one step's kernels
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        t += replays;
    }
    for (; t < r->T; ++t) {
        rc = fused ? fused_step(r, t, nullptr, stream) : rollout_step(r, t, nullptr, SN, kp, ko, has_obs, stream);
        if (rc) return rc;
    }
    if (fused) {                                                   // the last iteration's feature rebuild
        rc = piml_state_features_f32(r->p, r->v, r->a, r->dest, r->obstacles, r->obs_per_scene, r->S, r->N, r->M, r->kp,
                                     r->cos_p, r->thr_p, r->ko, r->cos_o, r->thr_o, r->hist_v, r->desired_speed,
                                     r->ped_f, r->obs_f, r->self_f, r->dest_f, stream);
        if (rc) return rc;
    }
    if (st != user) {                                              // the caller's stream continues after the loop
        PIML_CUDA(cudaEventRecord(ev, st));
        PIML_CUDA(cudaStreamWaitEvent(user, ev, 0));
    }
    return PIML_OK;
}
