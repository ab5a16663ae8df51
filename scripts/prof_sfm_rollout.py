"""One pure social-force rollout (S scenes x 122 slots x T frames) for ncu / timing of sfm_rollout_kernel."""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import piml_b200 as P
from piml_b200.rollout import rollout_scenes, state_features

ap = argparse.ArgumentParser(); ap.add_argument("--scenes", type=int, default=1); ap.add_argument("--frames", type=int, default=300)
a = ap.parse_args()
dev = torch.device("cuda"); S2, T2, Ns = a.scenes, a.frames, 122
fargs = (6, 90, 4, 10, 90, 4)
rargs = argparse.Namespace(time_unit=0.08, topk_ped=6, sight_angle_ped=90, dist_threshold_ped=4, topk_obs=10,
                           sight_angle_obs=90, dist_threshold_obs=4)
g = torch.Generator().manual_seed(3)
ob = (torch.rand(100, 2, generator=g) * 20).to(dev)
P0 = (torch.rand(S2, T2, Ns, 2, generator=g) * 20).to(dev); P0[:, :, 30:] = float('nan')
scene = {"position": P0, "velocity": torch.randn(S2, T2, Ns, 2, generator=g).to(dev) * 0.5,
         "acceleration": torch.zeros(S2, T2, Ns, 2, device=dev),
         "destination": (torch.rand(S2, T2, Ns, 2, generator=g) * 20).to(dev),
         "dest_idx": torch.zeros(S2, T2, Ns, dtype=torch.int64, device=dev),
         "waypoints": (torch.rand(S2, 1, Ns, 2, generator=g) * 20).to(dev),
         "dest_num": torch.ones(S2, Ns, dtype=torch.int64, device=dev), "obstacles": ob,
         "mask_p": torch.ones(S2, T2, Ns, device=dev), "mask_p_pred": torch.ones(S2, T2, Ns, device=dev),
         "desired_speed": torch.full((S2, Ns), 1.3, device=dev)}
f0 = state_features(P0[:, 0].contiguous(), scene["velocity"][:, 0].contiguous(), scene["acceleration"][:, 0].contiguous(),
                    scene["destination"][:, 0].contiguous(), ob, scene["velocity"][:, 0].contiguous(),
                    scene["desired_speed"], *fargs)
scene["ped_features0"], scene["obs_features0"], scene["self_features0"] = f0
spec = P.SocialForce("gc1560").spec
for mode in ("1", "0"):
    os.environ["PIML_SFM_PERSISTENT"] = mode
    for _ in range(2):
        rollout_scenes(spec, None, rargs, scene, 0, T2)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        rollout_scenes(spec, None, rargs, scene, 0, T2)
    e1.record(); torch.cuda.synchronize()
    print(f"persistent={mode}: S={S2} T={T2}: {e0.elapsed_time(e1) / 3:.3f} ms per rollout (events), {(time.perf_counter() - t0) / 3 * 1e3:.3f} ms wall")
