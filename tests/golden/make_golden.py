"""Freeze golden vectors from the LIVE reference (runs only in the build container).

    python tests/golden/make_golden.py          # needs /root/reference, writes tests/golden/*.npz

For every case the unmodified reference model classes are imported from
/root/reference/Software_Artifact/software, loaded with the seeded synthetic parameters of
oracle/seeded.py, and run through the reference's own ``_get_output`` source
(results_analyzer.py:236-270, exec'd inside a stub class because the module itself needs KDEpy
and matplotlib, which are not installed).  The only intervention is mask injection: each
``MCDropout`` module's forward is replaced by ``x * keep / (1-p)`` with ``keep`` from the Philox
contract of oracle/philox.py, because torch's own dropout stream is not reproducible across
devices.  Masksembles modules are left untouched (they rotate their own ``cnt``).

The script also asserts that the oracle restatement reproduces the reference on every case
before anything is written, so a committed fixture is at once a pin of the oracle and a
known-answer vector for the CUDA path.
"""
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_SW = "/root/reference/Software_Artifact/software"
sys.path.insert(0, REF_SW)

from oracle import nets, philox, seeded, stats, masksembles as o_masks   # noqa: E402

import utils as ref_utils                                                  # noqa: E402
from models.resnet18 import resnet18 as ref_resnet                         # noqa: E402
from models.vgg19 import vgg19 as ref_vgg                                  # noqa: E402


def _load_reference_methods():
    """exec the reference's _get_output and ece_hist_binary sources inside a stub class."""
    src = open(os.path.join(REF_SW, "train/results_analyzer.py")).read().split("\n")
    body = "\n".join(src[235:270]) + "\n\n" + "\n".join(src[445:495])
    code = "class Stub:\n" + textwrap.indent(textwrap.dedent(body), "    ")
    ns = {"np": np, "torch": torch, "nn": torch.nn}
    exec(code, ns)
    return ns["Stub"]


Stub = _load_reference_methods()


class Ctx:
    site = None


def inject(model, ctx):
    for name, mod in model.named_modules():
        if type(mod).__name__ == "MCDropout":
            mod.forward = (lambda x, _n=name: ctx.site(_n, x))


def run_case(tag, build, oracle_forward, spec_kind, p, B, S, seed, in_shape=(3, 32, 32)):
    np.random.seed(0)
    torch.manual_seed(0)
    model = build()
    sd = seeded.seeded_state_dict(model.state_dict(), seed=1234)
    model.load_state_dict(sd)
    model.eval()
    x = seeded.seeded_input((B,) + in_shape, seed=99)
    mask_tables = {k[:-len(".masks")]: v for k, v in sd.items() if k.endswith(".masks")}
    spec = nets.SiteSpec(kind=spec_kind, p=p, masks=mask_tables, cnt0=0)

    ctx = Ctx()
    inject(model, ctx)
    stub = Stub()
    stub.mc_dropout, stub.mc_passes, stub.outputs = True, S, list(range(model.n_exits))

    class PassCounter:                      # the reference calls self.model(b_x) S times
        i = -1
        out_dim = model.out_dim

        def __call__(self, b_x):
            PassCounter.i += 1
            ctx.site = nets.InjectedSites(spec, seed, PassCounter.i)
            return model(b_x)

    stub.model = PassCounter()
    with torch.no_grad():
        output, output_sm, output_sm_np, ens_out, ens_sm = stub._get_output(x)
    ref = {"mean_logits": np.stack([o.numpy() for o in output]),
           "mean_probs": np.asarray(output_sm_np),
           "ens_logits": np.stack([o.numpy() for o in ens_out]),
           "ens_probs": np.stack([o.numpy() for o in ens_sm])}

    with torch.no_grad():
        got = stats.mc_get_output(
            lambda i: oracle_forward(sd, x, nets.InjectedSites(spec, seed, i)), S)
    for k, v in ref.items():
        err = np.abs(got[k] - v).max()
        assert err < 2e-6, (tag, k, err)
    print("%-28s oracle == live reference (max err %.2e over %s)" % (
        tag, max(np.abs(got[k] - v).max() for k, v in ref.items()), list(ref)))
    np.savez_compressed(
        os.path.join(HERE, tag + ".npz"), x=x.numpy(), seed=np.int64(seed), S=np.int64(S),
        p=np.float64(p), all_logits=got["all_logits"].astype(np.float32),
        state_keys=np.array(sorted(model.state_dict().keys())),
        **{k: v for k, v in ref.items()},
        **{"masks/" + k: v.numpy() for k, v in mask_tables.items()})


def main():
    R, V = ref_resnet, ref_vgg
    run_case("resnet18_mcd_block",
             lambda: R.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", dropout_p=0.5, out_dim=10),
             lambda sd, x, s: nets.resnet18_forward(sd, x, s, "block", True), "mc", 0.5, 4, 4, 0x5EED)
    run_case("resnet18_mcd_exitonly",
             lambda: R.ResNet18MCEarlyExit(dropout_exit=True, dropout=None, dropout_p=0.125, out_dim=100),
             lambda sd, x, s: nets.resnet18_forward(sd, x, s, None, True), "mc", 0.125, 4, 4, 0x5EED)
    run_case("resnet18_mcd_layer",
             lambda: R.ResNet18MCEarlyExit(dropout_exit=True, dropout="layer", dropout_p=0.25, out_dim=10),
             lambda sd, x, s: nets.resnet18_forward(sd, x, s, "layer", True), "mc", 0.25, 2, 3, 0xABCDEF01)
    run_case("resnet18_mask_block",
             lambda: R.ResNet18MCEarlyExit(dropout_exit=True, dropout="block", out_dim=100,
                                           mask_type="mask", num_masks=4, mask_scale=2.0),
             lambda sd, x, s: nets.resnet18_forward(sd, x, s, "block", True), "mask", 0.0, 4, 6, 0)
    run_case("resnet18_mc_single",
             lambda: R.ResNet18MC(dropout_exit=True, dropout="block", dropout_p=0.375, out_dim=10),
             lambda sd, x, s: nets.resnet18_forward(sd, x, s, "block", True, early_exit=False),
             "mc", 0.375, 2, 3, 77)

    def vgg_last3():
        m = V.VGG19MCEarlyExit(dropout_exit=True, dropout=None, dropout_p=0.5, out_dim=100, image_size=32, n_exits=5)
        for i in (2, 3, 4):
            m.blocks[i].append(V.MCDropout(0.5))
        return m
    run_case("vgg19_mcd_last3", vgg_last3,
             lambda sd, x, s: nets.vgg19_forward(sd, x, s, True, (2, 3, 4)), "mc", 0.5, 4, 4, 0x5EED)
    run_case("vgg19_mcd_exitonly",
             lambda: V.VGG19MCEarlyExit(dropout_exit=True, dropout=None, dropout_p=0.25, out_dim=10, image_size=32, n_exits=5),
             lambda sd, x, s: nets.vgg19_forward(sd, x, s, True, ()), "mc", 0.25, 2, 3, 5)

    # ---- Masksembles mask generator: bit-exact restatement -----------------------------
    gen = {}
    for (c, n, scale, npseed) in [(64, 4, 2.0, 0), (128, 4, 2.0, 1), (256, 4, 2.0, 2), (512, 4, 2.0, 3),
                                  (512, 4, 4.0, 4), (512, 4, 6.0, 5), (64, 8, 3.0, 6), (100, 2, 1.5, 7)]:
        np.random.seed(npseed)
        want = ref_utils.generation_wrapper(c, n, scale)
        np.random.seed(npseed)
        got = o_masks.generation_wrapper(c, n, scale)
        assert want.shape == got.shape and (want == got).all(), (c, n, scale)
        gen["c%d_n%d_s%g_seed%d" % (c, n, scale, npseed)] = want.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "masksembles_masks.npz"), **gen)
    print("masksembles generator: %d configurations bit-exact" % len(gen))

    # ---- equal-mass histogram ECE --------------------------------------------------------
    cases = {}
    for k, (N, C) in enumerate([(1000, 10), (600, 20), (250, 7)]):
        logits = 3.0 * philox.normal(11, k, 0, N * C).reshape(N, C)
        p = np.exp(logits - logits.max(1, keepdims=True))
        p /= p.sum(1, keepdims=True)
        lab = seeded.seeded_labels(N, C, seed=k)
        # make labels correlate with the prediction so accuracy is not ~1/C
        agree = philox.uniform01(13, k, 0, N) < 0.6
        lab = np.where(agree, p.argmax(1), lab)
        onehot = np.eye(C)[lab]
        want = float(Stub().ece_hist_binary(p, onehot))
        got = stats.ece_hist(p, onehot)
        assert abs(want - got) < 1e-7, (want, got)
        cases["p%d" % k], cases["label%d" % k], cases["ece%d" % k] = p, lab, np.float64(want)
    np.savez_compressed(os.path.join(HERE, "ece_hist.npz"), **cases)
    print("ece_hist: 3 cases match the reference source")


if __name__ == "__main__":
    main()
