"""GPU: the Masksembles BATCHED formulation through the fused plan, against vectors frozen from
the live reference (tests/golden/make_golden_masksembles_batched.py):
  (A) training branch of the reference's Masksembles modules (utils.py:158-164, :220-226) inside its unmodified
      ResNet18MCEarlyExit - batch split into n groups, group g uses mask g;
  (B) the Keras converter's inference formulation (converter/keras/Masksembles.py:216-239) - input tiled n times,
      outputs reshaped [n, -1, C], averaged over the masks and then across the exits."""
import os

import numpy as np
import pytest
import torch

from bayesnn_fpga_b200 import masksembles_batched_predict
from tests.cases import GOLDEN, build_seeded
from tests.gpu_util import report

pytestmark = pytest.mark.gpu


def _model():
    model, sd, gold = build_seeded("resnet18_mask_block")
    return model.cuda().eval()


@pytest.mark.parametrize("dt,tol", [("fp32", 1e-5), ("fp16", 1e-3)])
def test_training_branch_groups_through_the_fused_plan(dt, tol):
    z = np.load(os.path.join(GOLDEN, "masksembles_batched.npz"))
    model = _model()
    model.bnn_dtype = dt
    mods = model._masksembles_modules()
    assert len(mods) == 7                                  # 3 Masksembles2D behind the stages + 4 Masksembles1D at the exits
    for m in mods:
        m.train()                                          # ONLY the Masksembles layers: BatchNorm keeps running statistics
        m.cnt = 3                                          # the eval-mode counter must neither matter nor move
    x = torch.from_numpy(z["groups_x"]).cuda()
    outs = model(x)
    want = z["groups_logits"]
    scale = max(1.0, float(np.abs(want).max()))
    err = max(float(np.abs(o.double().cpu().numpy() - w).max()) for o, w in zip(outs, want))
    report(test="masksembles_groups", dtype=dt, err=err, scale=scale)
    assert len(outs) == 4 and outs[0].shape == (8, 100) and err <= tol * scale
    assert all(int(m.cnt) == 3 for m in mods)
    with pytest.raises(ValueError) as e:
        model(x[:6])
    assert str(e.value) == str(z["groups_error"][0])
    # back in eval mode the rotating single-mask branch is used again
    for m in mods:
        m.eval()
        m.cnt = 0
    assert model(x)[0].shape == (8, 100) and int(mods[0].cnt) == 1


@pytest.mark.parametrize("dt,tol", [("fp32", 1e-5), ("fp16", 1e-3)])
def test_keras_batched_inference_formulation(dt, tol):
    z = np.load(os.path.join(GOLDEN, "masksembles_batched.npz"))
    model = _model()
    for m in model._masksembles_modules():
        m.cnt = 2
    r = masksembles_batched_predict(model, torch.from_numpy(z["batched_x"]), dtype=dt)
    e1 = float(np.abs(r["per_exit"].double().cpu().numpy() - z["batched_exit_probs"]).max())
    e2 = float(np.abs(r["prediction"].double().cpu().numpy() - z["batched_avg"]).max())
    report(test="masksembles_batched_keras", dtype=dt, per_exit=e1, prediction=e2)
    assert r["n"] == 4 and e1 <= tol and e2 <= tol
    assert all(int(m.cnt) == 2 for m in model._masksembles_modules())
    # the prefix (conv1 + layer1, in front of the first Masksembles layer) ran once per image, not n times
    eng = model.bnn_engine(dt)
    pre, suf = eng.graph.macs()
    assert pre == 152764416 and suf > 0
