"""`args` namespaces for the tests (the subset of reference src/main.py:26-112 the hot path reads)."""
import argparse


def base_args(**over):
    a = argparse.Namespace(
        model='pinnsf_m', dataset_name='ucy', dropout=0.5, activation='relu',
        topk_ped=6, topk_obs=10, sight_angle_ped=90, sight_angle_obs=90, dist_threshold_ped=4, dist_threshold_obs=4,
        encoder_hidden_size=128, processor_hidden_size=128, decoder_hidden_size=64,
        encoder_hidden_layers=3, processor_hidden_layers=16, decoder_hidden_layers=2,
        num_history_velocity=1, skip_frames=25, ped_feature_dim=6, obs_feature_dim=6, self_feature_dim=7,
        time_unit=0.08, device='cuda', gpus='0')
    for k, v in over.items():
        setattr(a, k, v)
    return a


def model_args(kind, cfg, dataset_name):
    hs_e, hs_p, hs_d, nl_e, nl_p, nl_d, has_obs = [int(v) for v in cfg]
    return base_args(model=kind, dataset_name=dataset_name, encoder_hidden_size=hs_e, processor_hidden_size=hs_p,
                     decoder_hidden_size=hs_d, encoder_hidden_layers=nl_e, processor_hidden_layers=nl_p,
                     decoder_hidden_layers=nl_d, obs_feature_dim=6 if has_obs else 0)
