"""GPU: plan lifetime.  A model that has run through the engine still pickles / deep-copies, and a plan never outlives
the weights it was folded from (fine-tuning step, load_state_dict, appended site, generic modules, the converter)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

from bayesnn_fpga_b200 import Dropouts, mc_predict, nn2bnn
from tests.cases import build_seeded, oracle_run

pytestmark = pytest.mark.gpu


def test_model_pickles_and_copies_after_a_run_and_replans_on_weight_change(tmp_path):
    model, sd, gold = build_seeded("resnet18_mcd_block")
    model.cuda()
    x = torch.from_numpy(gold["x"])
    S = 3
    r0 = mc_predict(model, x, S, seed=5, dtype="fp32").mean_probs.clone()
    model(x.cuda())                                                      # the single-pass forward too
    torch.save(model, str(tmp_path / "ran.pt"))                          # CDLL / CUDA graphs must not be in the module
    again = torch.load(str(tmp_path / "ran.pt"), weights_only=False)
    clone = copy.deepcopy(model)
    for other in (again, clone):
        r = mc_predict(other.cuda(), x, S, seed=5, dtype="fp32").mean_probs
        assert torch.equal(r, r0)
    # an in-place update of ONE sub-module's weights (what an optimiser step or a partial load_state_dict does)
    with torch.no_grad():
        model.ex1linear.weight.mul_(0.5)
        model.layer3[0][1].bn2.bias.add_(0.25)
    sd2 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    want = oracle_run("resnet18_mcd_block", sd2, x, S, 5, 0.5)
    got = mc_predict(model, x, S, seed=5, dtype="fp32").mean_probs.double().cpu().numpy()
    assert np.abs(got - want["mean_probs"]).max() <= 1e-5
    assert np.abs(got - r0.double().cpu().numpy()).max() > 1e-3          # and it really is a different answer


def test_generic_module_and_converter_replan_when_weights_change(tmp_path):
    torch.manual_seed(0)
    net = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Flatten(), nn.Linear(8 * 8 * 8, 10)).cuda().eval()
    x = torch.randn(4, 3, 8, 8).cuda()
    bnn = nn2bnn.MCDropout(copy.deepcopy(net), nSamples=4, p=0.25).reseed(3).eval()
    a = bnn(x).clone()
    with torch.no_grad():
        bnn.model[3].layer.weight.mul_(2.0)
        bnn.model[3].layer.bias.zero_()
    b = bnn.reseed(3)(x).clone()
    assert (a - b).abs().max().item() > 1e-3
    torch.save(bnn, str(tmp_path / "conv.pt"))                                  # wrapper stays picklable after running
    g = copy.deepcopy(net)
    g[3] = nn.Sequential(Dropouts.MCDropout(0.25), g[3])
    r1 = mc_predict(g, x, 4, seed=1, dtype="fp32").mean_logits.clone()
    with torch.no_grad():
        g[0].weight.mul_(-1.0)
    r2 = mc_predict(g, x, 4, seed=1, dtype="fp32").mean_logits
    assert (r1 - r2).abs().max().item() > 1e-3
