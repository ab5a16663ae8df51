"""Free-running rollout drift against the reference's own trajectories (golden fixtures): max / mean ||dp|| of the
active agents vs step, for the FP32-pipe and the tensor-core forward.  Writes gpurun_out/rollout_drift.md."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_gpu_parity import _rollout_inputs, cu, npy   # noqa: E402
from piml_b200 import models as M                             # noqa: E402
from piml_b200.rollout import rollout_scenes                  # noqa: E402

lines = ["# Free-running rollout drift vs the reference (golden trajectories of `get_multiple_rollouts`)", "",
         "Positions after k free-running steps (no re-synchronisation) compared with the reference's own CPU rollout of",
         "the same clip and seed-666 weights.  The dynamics are chaotic: the reference drifts 3e-4 .. 6e-3 m from ITSELF",
         "after 725 steps when its start frame is perturbed by 1 ulp (SURVEY.md 8d), which is the envelope to read these",
         "numbers against.  `mask_p` and the NaN (arrival) pattern are compared exactly.", "",
         "| clip / model | forward kernel | steps | max drift [m] at +1 / +10 / +100 / +300 / end | mean drift at end [m] | mask_p equal | NaN pattern equal |",
         "|---|---|---|---|---|---|---|"]
for name in ("rollout_gc_bm", "rollout_toy5_m", "rollout_ucy_bm"):
    z, i, o, args, m = _rollout_inputs(name)
    T, t0 = int(i["num_frames"]), int(i["t_start"])
    scene = {k: cu(i[k])[None] for k in ("position", "velocity", "acceleration", "destination", "waypoints",
                                          "mask_p", "mask_p_pred", "desired_speed")}
    scene["dest_idx"] = cu(i["dest_idx"], torch.int64)[None]
    scene["dest_num"] = cu(i["dest_num"], torch.int64)[None]
    scene["obstacles"] = cu(i["obstacles"])
    for k in ("ped_features0", "obs_features0", "self_features0"):
        scene[k] = cu(i[k])[None]
    packed = M.pack_device(m.state_dict(), m.spec)
    ptc = M.pack_device_tc(m.state_dict(), m.spec)
    for label, tc in (("FP32 pipe", None), ("tcgen05 3xTF32", ptc)):
        p_res, v_res, a_res, mask = rollout_scenes(m.spec, packed, args, scene, t0, T, packed_tc=tc)
        p_res, mask = npy(p_res[0]), npy(mask[0])
        drift = np.linalg.norm(p_res - o["position"], axis=-1)
        def mx(k):
            t = min(T - 1, t0 + k)
            return np.nanmax(drift[:t + 1]) if np.isfinite(drift[:t + 1]).any() else float("nan")
        end_mean = np.nanmean(drift[T - 1]) if np.isfinite(drift[T - 1]).any() else np.nanmean(drift[np.isfinite(drift).any(1)][-1])
        lines.append(f"| {name} | {label} | {T - t0} | {mx(1):.2e} / {mx(10):.2e} / {mx(100):.2e} / {mx(300):.2e} / "
                     f"{np.nanmax(drift):.2e} | {end_mean:.2e} | {np.array_equal(mask, o['mask_p'])} | "
                     f"{np.array_equal(np.isnan(p_res), np.isnan(o['position']))} |")
# pure social-force rollout (BASELINE config 2): the composed reference module on the synthetic clip
import piml_b200 as P                                         # noqa: E402
from tests.golden_args import base_args                       # noqa: E402
from tests.test_gpu_parity import _sfm_scene                  # noqa: E402
from tests.util import golden, group                          # noqa: E402
z = golden("rollout_syn_sfm")
i, o = group(z, "in"), group(z, "out")
T, t0 = int(i["num_frames"]), int(i["t_start"])
args = base_args(model="sfm", dataset_name="gc1560", time_unit=float(i["time_unit"]))
p_res, v_res, a_res, mask = rollout_scenes(P.SocialForce("gc1560").spec, None, args, _sfm_scene(i), t0, T)
p_res, mask = npy(p_res[0]), npy(mask[0])
drift = np.linalg.norm(p_res - o["position"], axis=-1)
mx = lambda k: np.nanmax(drift[:min(T - 1, t0 + k) + 1])
lines.append(f"| rollout_syn_sfm (pure social force) | sfm_forward_kernel | {T - t0} | {mx(1):.2e} / {mx(10):.2e} / "
             f"{mx(100):.2e} / {mx(300):.2e} / {np.nanmax(drift):.2e} | {np.nanmean(drift[np.isfinite(drift).any(1)][-1]):.2e} | "
             f"{np.array_equal(mask, o['mask_p'])} | {np.array_equal(np.isnan(p_res), np.isnan(o['position']))} |")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "rollout_drift.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
