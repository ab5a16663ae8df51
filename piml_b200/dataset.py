"""Feature build of a whole clip on the CUDA path: drop-in body for `TimeIndexedPedData.make_dataset`
(reference src/data/data.py:746-833).

The reference spends 7.2 s per GC clip here: `get_relative_features` over all T frames (:766-771), a Python double
loop over pedestrians x frames for the desired speed (:797-806) and `calculate_collision_label` (:814).  Here those
three are kernels of libpiml_b200.so (T is the batch dimension of the feature kernels); what remains on the host is
tensor plumbing (slices, concatenations) and the reference's OWN `turn_detection` / `move_index_matrix`, which are
called on `self` and stay untouched reference code.
"""
import torch

from . import _lib as L
from .features import Pedestrians

_PEDS = Pedestrians()


def desired_speed(velocity, skip_frames):
    """data.py:797-806: per pedestrian the mean speed over the `skip_frames` frames from its first moving frame.
    velocity (T,N,2) -> (N,)."""
    if velocity.dim() != 3:
        raise ValueError("desired_speed expects velocity (T,N,2)")
    _, origin, (velocity,) = L.stage(velocity)
    v = L.f32c(velocity)
    T, N = v.shape[0], v.shape[1]
    out = torch.empty(N, dtype=torch.float32, device=v.device)
    L.check(L.load().piml_desired_speed_f32(L.ptr(v), T, N, int(skip_frames), L.ptr(out), L.stream_ptr(v.device)),
            "piml_desired_speed_f32")
    return out.to(origin)


def history_velocity(velocity, h):
    """data.py:788-794: slot i of frame t holds the velocity of frame t-(h-1-i) (zeros before the clip starts)."""
    T, N = velocity.shape[0], velocity.shape[1]
    hist = velocity.new_zeros(T, N, h, 2)
    for i in range(h):
        lag = h - 1 - i
        hist[lag:, :, i, :] = velocity[:T - lag]
    return hist.reshape(T, N, 2 * h)


def make_dataset(self, args, raw_data):
    """`self` is the reference's TimeIndexedPedData (or anything with turn_detection / move_index_matrix /
    get_relative_features / calculate_collision_label); fills the same attributes as data.py:766-833."""
    get_features = getattr(self, "get_relative_features", _PEDS.get_relative_features)
    ped_f, obs_f, dest_f = get_features(
        raw_data.position, raw_data.velocity, raw_data.acceleration, raw_data.destination, raw_data.obstacles,
        args.topk_ped, args.sight_angle_ped, args.dist_threshold_ped, args.topk_obs, args.sight_angle_obs,
        args.dist_threshold_obs)
    raw_data.to(args.device)
    ped_f, obs_f, dest_f = ped_f.to(args.device), obs_f.to(args.device), dest_f.to(args.device)
    self.abnormal_mask = self.turn_detection(raw_data)                                   # reference code, :779
    self.ped_features = ped_f
    T, N = ped_f.shape[0], ped_f.shape[1]
    self.obs_features = obs_f if len(obs_f) > 0 else torch.tensor([[] for _ in range(T)], device=ped_f.device)
    hist = history_velocity(raw_data.velocity, args.num_history_velocity)
    speed = desired_speed(raw_data.velocity, args.skip_frames).to(ped_f.device)          # kernel, replaces :797-806
    self.self_features = torch.cat((dest_f, hist, raw_data.acceleration, speed.reshape(1, N, 1).repeat(T, 1, 1)), -1)
    labels = getattr(self, "calculate_collision_label", _PEDS.calculate_collision_label)(ped_f)
    self.labels = torch.cat((raw_data.position, raw_data.velocity, raw_data.acceleration, labels), dim=-1)
    skip = args.skip_frames
    self.mask_a_pred = self.move_index_matrix(raw_data.mask_a, 'backward', skip - 1, dim=0)
    self.mask_v_pred = self.move_index_matrix(raw_data.mask_v, 'backward', skip - 1, dim=0)
    self.mask_p_pred = self.move_index_matrix(raw_data.mask_p, 'backward', skip - 1, dim=0)
    self.mask_a_pred = self.move_index_matrix(self.mask_a_pred, 'forward', 1, dim=0)     # last frame: no label (:824)
    self.meta_data = raw_data.meta_data
    self.topk_obs = args.topk_obs
    self.num_frames = self.dataset_len = T
    self.num_pedestrians = N
    self.ped_feature_dim = ped_f.shape[-1]
    self.obs_feature_dim = self.obs_features.shape[-1]
    self.self_feature_dim = self.self_features.shape[-1]
