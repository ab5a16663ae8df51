"""CPU: the plain-torch restatements in tests/torch_ref.py (the fp32 reference of the backward kernels' random-input
GPU tests) are pinned against golden vectors produced by the UNMODIFIED reference modules
(tests/golden/make_golden.py `training` group -> training_step.npz)."""
import numpy as np
import pytest
import torch

from . import torch_ref as TR
from .util import golden, group


def drop_masks(g, p=0.5):
    out = []
    for name in ("ped", "obs"):
        shape = tuple(int(x) for x in g["dropshape_" + name])
        n = int(np.prod(shape))
        bits = np.unpackbits(g["dropbits_" + name])[:n].reshape(shape)
        out.append(torch.from_numpy(bits.astype(np.float32)) / (1.0 - p))
    return out


def load_case(z, case):
    g = group(z, case)
    sd_grads = {k[len("grad/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("grad/")}
    return g, sd_grads


def reference_weights(kind, dsn):
    """The weights the golden case used: models.npz holds the state_dict of the same seeded construction."""
    zm = golden("models")
    g = group(zm, kind)
    assert str(g["dataset_name"]) == dsn or kind == "pinnsf_bm"
    return {k[len("sd/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd/")}, float(g["tau"])


CASES = [("pinnsf_bm_eval", "pinnsf_bm", "ped", False), ("pinnsf_bm_train", "pinnsf_bm", "ped", True),
         ("pinnsf_m_eval", "pinnsf_m", "ped", False), ("pinnsf_m_train", "pinnsf_m", "ped", True),
         ("pinnsf_bm_chan", "pinnsf_bm", "ped_c", False)]


@pytest.mark.parametrize("case,kind,inp,train", CASES)
def test_torch_ref_matches_reference_autograd(case, kind, inp, train):
    z = golden("training_step")
    g, ref_grads = load_case(z, case)
    sd, tau = reference_weights(kind, str(g["dataset_name"]))
    if case == "pinnsf_bm_chan":
        tau = 5 / 6                                   # dataset_name 'ucy' (model.py:1151-1154)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    sfx = "_c" if inp == "ped_c" else ""
    ped, obs, slf = [torch.from_numpy(z[n + sfx]).clone().requires_grad_(True) for n in ("ped", "obs", "self")]
    dp, do = drop_masks(g) if train else (None, None)
    outs = TR.pinnsf_forward_ref(sd, kind, tau, ped, obs, slf, True, dp, do)
    loss = sum((o * torch.from_numpy(g[f"w{i}"])).sum() for i, o in enumerate(outs))
    loss.backward()
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.detach().numpy(), g[f"out{i}"], rtol=2e-5, atol=2e-5)
    assert abs(float(loss) - float(g["loss"])) <= 2e-4 * max(1.0, abs(float(g["loss"])))
    for name, t in (("g_ped", ped), ("g_obs", obs), ("g_self", slf)):
        ref = g[name]
        err = np.abs(t.grad.numpy() - ref).max() / max(np.abs(ref).max(), 1e-12)
        assert err < 2e-5, (name, err)
    for k, ref in ref_grads.items():
        got = sd[k].grad.numpy()
        err = np.abs(got - ref.numpy()).max() / max(np.abs(ref.numpy()).max(), 1e-12)
        assert err < 5e-5, (k, err)
    # dead weights (ResDNN block 0) get no gradient in the reference
    assert any("processor" in d for d in g["dead"].tolist())
    for k in g["dead"].tolist():
        assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0


def test_collision_detection_ref_shapes():
    torch.manual_seed(0)
    p = torch.rand(6, 9, 2) * 2
    p[2, 3] = float("nan")
    c3 = TR.collision_detection_ref(p, 0.5)
    assert c3.shape == (6, 9, 9) and float(c3.diagonal(dim1=-2, dim2=-1).abs().max()) == 0
    c4 = TR.collision_detection_ref(p.reshape(2, 3, 9, 2), 0.5)
    assert c4.shape == (2, 3, 9, 9)
    assert set(np.unique(c3.numpy()).tolist()) <= {0.0, 1.0}


@pytest.mark.parametrize("case", ["ucy", "gc", "tiny"])
def test_loss_restatements_match_reference_methods(case):
    """f-3: the torch restatements of the rollout losses (tests/torch_ref.py multiple_rollout_*) against the
    reference's own BaseSimulator methods (simulators.py:172-249) called on the same seeded tensors: values and d/d pred."""
    from tests import torch_ref as TR
    from tests.util import golden, group
    g = group(golden("losses"), case)
    wide = torch.from_numpy(g["wide"])
    pred = torch.from_numpy(g["pred"]).requires_grad_(True)
    a_pred = torch.from_numpy(g["a_pred"]).requires_grad_(True)
    am = torch.from_numpy(g["abnormal_mask"]) if "abnormal_mask" in g else None
    decay = float(g["decay"])
    labels = wide[..., :2]
    mse = TR.multiple_rollout_mse_loss(pred, labels, decay, 'sum')
    cl = TR.multiple_rollout_collision_loss(pred, labels, decay, 10, torch.from_numpy(g["coll"]).clone(), 'sum', am)
    hl = TR.multiple_rollout_collision_loss(pred, labels, decay, 10, torch.from_numpy(g["hard"]).clone(), 'sum', am)
    (mse + 10.0 * cl + 100.0 * hl).backward()
    amse = TR.multiple_rollout_mse_loss(a_pred, wide[..., 4:6], decay, 'sum', reverse=True)
    amse.backward()
    for got, key in ((mse, "mse"), (cl, "collision"), (hl, "hard_collision"), (amse, "a_mse")):
        assert abs(float(got.detach()) - float(g[key])) <= 1e-6 * max(abs(float(g[key])), 1.0), key
    assert np.allclose(pred.grad.numpy(), g["g_pred"], rtol=1e-5, atol=1e-5)
    assert np.allclose(a_pred.grad.numpy(), g["g_a_pred"], rtol=1e-5, atol=1e-6)
