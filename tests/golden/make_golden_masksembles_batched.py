"""Freeze golden vectors of the Masksembles BATCHED formulation from the LIVE reference (build container only).

    python tests/golden/make_golden_masksembles_batched.py    # needs /root/reference; writes masksembles_batched.npz

(A) the PyTorch training branch of the reference's own Masksembles1D / 2D modules (utils.py:158-164, :220-226: the
    batch is split into n groups, group g is multiplied by mask g) inside the reference's unmodified
    ResNet18MCEarlyExit: the network is in eval mode (BatchNorm uses running statistics), ONLY its Masksembles modules
    are switched to training mode - which is what selects that branch (`if self.training`).
(B) the Keras converter's inference formulation (converter/keras/Masksembles.py:216-239 `MasksemblesModel.call(
    training=False)`), which cannot run here (no TensorFlow), executed literally on the PyTorch reference network: the
    input is tiled n times along the batch, the layers split it into n groups (== branch (A) on the tiled batch), every
    output is soft-maxed (the Keras models end in softmax), reshaped [n, -1, C] and averaged over the n groups, and the
    per-exit predictions are then averaged across exits (:233).
The Masksembles tables are the ones frozen in resnet18_mask_block.npz.  Asserts oracle == reference before writing."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_SW = "/root/reference/Software_Artifact/software"
sys.path.insert(0, REF_SW)

from oracle import nets, seeded                                            # noqa: E402
from models.resnet18 import resnet18 as ref_resnet                         # noqa: E402


def main():
    base = np.load(os.path.join(HERE, "resnet18_mask_block.npz"))
    np.random.seed(0)
    torch.manual_seed(0)
    model = ref_resnet.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", out_dim=100, mask_type="mask", num_masks=4,
                                           mask_scale=2.0)
    sd = seeded.seeded_state_dict(model.state_dict(), seed=1234)
    for k in list(sd):
        if k.endswith(".masks"):
            sd[k] = torch.from_numpy(base["masks/" + k[:-len(".masks")]])
    model.load_state_dict(sd)
    model.eval()
    n = 4
    for m in model.modules():
        if type(m).__name__ in ("Masksembles1D", "Masksembles2D"):
            m.train()                                    # selects the batch-group branch; BatchNorm stays in eval mode
    tables = {k[:-len(".masks")]: v for k, v in sd.items() if k.endswith(".masks")}
    spec = nets.SiteSpec("mask", 0.0, tables)
    out = {}
    # ---- (A) one forward over a batch of 8 = 4 groups of 2
    x = seeded.seeded_input((8, 3, 32, 32), seed=41)
    with torch.no_grad():
        ref = [o.numpy() for o in model(x)]
        got = [o.numpy() for o in nets.resnet18_forward(sd, x, nets.GroupSites(spec), "block", True)]
    for a, b in zip(ref, got):
        assert np.abs(a - b).max() < 2e-6
    out["groups_x"], out["groups_logits"] = x.numpy(), np.stack(ref)
    try:
        model(x[:6])
        raise AssertionError("batch 6 accepted")
    except ValueError as e:
        out["groups_error"] = np.array([str(e)])
    # ---- (B) MasksemblesModel.call(training=False)
    xb = seeded.seeded_input((3, 3, 32, 32), seed=42)
    with torch.no_grad():
        tiled = torch.cat([xb] * n)                                              # tf.tile(input, [num_masks, 1, 1, 1])
        preds = [torch.softmax(o, dim=1) for o in model(tiled)]
        per_exit = [p.reshape(n, -1, p.shape[-1]).mean(0) for p in preds]        # reduce_mean(reshape(pred, [n, -1, C]), 0)
        avg = sum(per_exit) / len(per_exit)                                      # prediction = sum(prediction) / len
        o_preds = [torch.softmax(o, dim=1) for o in nets.resnet18_forward(sd, tiled, nets.GroupSites(spec), "block", True)]
        o_exit = [p.reshape(n, -1, p.shape[-1]).mean(0) for p in o_preds]
    for a, b in zip(per_exit, o_exit):
        assert (a - b).abs().max().item() < 2e-6
    out["batched_x"] = xb.numpy()
    out["batched_exit_probs"] = np.stack([p.numpy() for p in per_exit])
    out["batched_avg"] = avg.numpy()
    np.savez_compressed(os.path.join(HERE, "masksembles_batched.npz"), **out)
    print("masksembles batched formulation: training-branch forward and Keras inference formulation frozen "
          "(oracle == live reference)")


if __name__ == "__main__":
    main()
