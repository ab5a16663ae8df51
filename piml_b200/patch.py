"""Drop the CUDA path in behind the reference's own call signatures (SURVEY.md 8b "selection switch").

    import data.data as DATA, models.mlapm as MLAPM_MOD, models.model as MODEL, models.simulators as SIM, utils.utils as UTILS
    import piml_b200.patch as patch
    patch.install(DATA=DATA, MLAPM_MOD=MLAPM_MOD, MODEL=MODEL, SIM=SIM, UTILS=UTILS)     # or patch.install_from_env(...)
    (optionally METRIC=functions.metrics for the evaluation metrics)

After `install` the unmodified reference scripts (src/main.py, src/main_mlapm.py) run their hot path in
libpiml_b200.so: the five signatures below keep their names, argument order and return types; everything else
(data loading, args, trainer, metrics) is untouched reference code.  `uninstall()` restores the originals, so with the
switch off the reference is byte-for-byte itself.  Opt-in only: `install_from_env` does nothing unless PIML_B200=1.

    Pedestrians.get_heading_direction / get_nearby_obj_in_sight / get_relative_features / calculate_collision_label
                                                                   (src/data/data.py:351, :416, :466, :515)
    PINNSF / PINNSF_bottleneck / PINNSF_bottleneck_multitask / PINNSF_multitask .forward
                                                                   (src/models/model.py:762, :1104, :1185, :1271)
    MLAPM.step                                                     (src/models/mlapm.py:10)
    UTILS.calc_acceleration                                        (src/utils/utils.py:31)
    BaseSimulator.get_multiple_rollouts                            (src/models/simulators.py:556)
    BaseSimulator.test_multiple_rollouts_for_training              (src/models/simulators.py:659)
    Pedestrians.collision_detection                                (src/data/data.py:538)
    TimeIndexedPedData.make_dataset                                (src/data/data.py:746; feature build of a whole clip)
    METRIC.collision_count / mae_with_time_mask / ot_with_time_mask / mmd_with_time_mask   (src/functions/metrics.py:16-91;
                                                                   pass METRIC=functions.metrics; (T,N,2) inputs)

Training: the patched model forwards record a CUDA backward (piml_b200.autograd), so the reference's own
`loss.backward(); optimizer.step()` (simulators.py:359-360) runs the library's backward kernels unchanged.
"""
import os

from . import dataset as _dataset
from . import features as _features
from . import metrics as _metrics
from . import mlapm as _mlapm
from . import models as _models
from . import rollout as _rollout
from . import sfm as _sfm
from . import train_rollout as _train_rollout

_saved = []          # (owner, attribute name, original)

PINNSF_CLASSES = ("PINNSF", "PINNSF_bottleneck", "PINNSF_bottleneck_multitask", "PINNSF_multitask")
PEDESTRIAN_METHODS = ("get_heading_direction", "get_nearby_obj_in_sight", "get_relative_features",
                      "_relative_features_raw", "calculate_collision_label", "collision_detection")


_MISSING = object()


def _swap(owner, name, new):
    if name in owner.__dict__:
        orig = owner.__dict__[name]
    else:
        orig = getattr(owner, name, _MISSING)       # helpers the reference class does not have are removed again
    _saved.append((owner, name, orig))
    setattr(owner, name, new)


def install(DATA=None, MLAPM_MOD=None, MODEL=None, SIM=None, UTILS=None, METRIC=None):
    """Patch whichever of the reference modules are given.  Returns the list of patched qualified names."""
    done = []
    if DATA is not None:
        for m in PEDESTRIAN_METHODS:
            _swap(DATA.Pedestrians, m, _features.Pedestrians.__dict__[m])
            done.append(f"data.data.Pedestrians.{m}")
        if hasattr(DATA, "TimeIndexedPedData"):
            _swap(DATA.TimeIndexedPedData, "make_dataset", _dataset.make_dataset)
            done.append("data.data.TimeIndexedPedData.make_dataset")
    if MODEL is not None:
        def forward(self, ped_features, obs_features, self_features):
            return _models.forward_from_module(self, ped_features, obs_features, self_features)
        for cls in PINNSF_CLASSES:
            if hasattr(MODEL, cls):
                _swap(getattr(MODEL, cls), "forward", forward)
                done.append(f"models.model.{cls}.forward")
    if MLAPM_MOD is not None:
        def step(self, position, velocity, desired_speed, destination, dt, radius=0.3):
            impl = self.__dict__.get("_piml_b200")
            if impl is None or impl.args is not self.args:
                impl = _mlapm.MLAPM(**self.args)
                impl.args = self.args
                self.__dict__["_piml_b200"] = impl
            return impl.step(position, velocity, desired_speed, destination, dt, radius)
        _swap(MLAPM_MOD.MLAPM, "step", step)
        done.append("models.mlapm.MLAPM.step")
    if UTILS is not None:
        _swap(UTILS, "calc_acceleration", _sfm.calc_acceleration)
        done.append("utils.utils.calc_acceleration")
    if SIM is not None:
        raw_cls = getattr(getattr(SIM, "DATA", None), "RawData", None)

        def get_multiple_rollouts(self, data, t_start=0, load_model=True):
            return _rollout.get_multiple_rollouts(self, data, t_start, load_model, result_cls=raw_cls)
        _swap(SIM.BaseSimulator, "get_multiple_rollouts", get_multiple_rollouts)
        done.append("models.simulators.BaseSimulator.get_multiple_rollouts")

        def test_multiple_rollouts_for_training(self, data, t_start=0):
            return _train_rollout.test_multiple_rollouts_for_training(self, data, t_start)
        _swap(SIM.BaseSimulator, "test_multiple_rollouts_for_training", test_multiple_rollouts_for_training)
        done.append("models.simulators.BaseSimulator.test_multiple_rollouts_for_training")
    if METRIC is not None:
        for m in ("collision_count", "mae_with_time_mask", "ot_with_time_mask", "mmd_with_time_mask"):
            _swap(METRIC, m, getattr(_metrics, m))
            done.append(f"functions.metrics.{m}")
    return done


def install_from_env(**modules):
    """install(**modules) iff the environment says PIML_B200=1 (the opt-in switch); else a no-op."""
    if os.environ.get("PIML_B200", "0") == "1":
        return install(**modules)
    return []


def uninstall():
    while _saved:
        owner, name, orig = _saved.pop()
        if orig is _MISSING:
            delattr(owner, name)
        else:
            setattr(owner, name, orig)
