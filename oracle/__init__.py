"""CPU oracle for the multi-exit MC-dropout / Masksembles inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  ``bayesnn_fpga_b200`` never imports this package.

What it restates (reference = os-hxfan/BayesNN_FPGA, paths relative to its root):

* ``Software_Artifact/software/models/resnet18/resnet18.py:302-346`` (forward of the
  multi-exit ResNet-18), ``:32-48`` (BasicBlock), ``:207-210`` (MCDropout)
* ``Software_Artifact/software/models/vgg19/vgg19.py:294-324`` (multi-exit VGG-19)
* ``Software_Artifact/software/utils.py:18-110`` (Masksembles mask generator),
  ``:156-169``/``:218-231`` (Masksembles eval branch)
* ``Software_Artifact/software/train/results_analyzer.py:236-270`` (S-pass loop, softmax,
  fp64 means, cumulative exit ensembles), ``:446-495`` (equal-mass histogram ECE)
* ``Hardware_Artifact/bayes_hw/metric_utils.py:3-6`` (predictive entropy)
* ``Hardware_Artifact/bayes_hw/models/t_qmodels_bayes_me.py:41-147`` (multi-exit LeNet spec)
* ``Hardware_Artifact/converter/pytorch/{Dropouts,nn2bnn}.py`` (converter semantics)

The arithmetic of the reference lives in PyTorch ATen (third party, not vendored in the
reference; ``torch==1.12.1`` pinned in ``Hardware_Artifact/requirements.txt:6``, 2.11.0
installed here), so the oracle is a functional fp32 torch-CPU restatement over a plain
``state_dict``.

Parity pin: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 8c).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE
ITSELF: ``tests/golden/make_golden.py`` imports the live reference modules from
``/root/reference`` in the build container, injects the oracle's Philox masks into the
reference's own dropout modules, and freezes inputs/outputs as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` replays them without the reference present.
"""
