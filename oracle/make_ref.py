"""Recipe for ``oracle/_ref/``: the reference's OWN source files of the hot path, staged unmodified next to the oracle
(TEST INFRASTRUCTURE, see oracle/__init__.py).

    python oracle/make_ref.py            # needs /root/reference (build container); writes oracle/_ref/ + MANIFEST.json

The reference is pure Python, so "building" it is staging it: the files below are copied byte for byte from where they
lie under /root/reference into ``oracle/_ref/`` (git-ignored - no reference source enters this repository's history -
but not gpurun-ignored, so the directory travels to the GPU box, where /root/reference does not exist).  They are used
for two things only: ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's own CPU implementation of
the path (``oracle/ref_arm.py``), and the oracle restatement is validated against them.  ``MANIFEST.json`` records the
source path and sha256 of every file so a stale or edited copy is detected.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("BNN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")

SW = "Software_Artifact/software/"
HW = "Hardware_Artifact/"
FILES = [
    # the multi-exit networks, their stochastic layers and factories
    SW + "utils.py",
    SW + "models/__init__.py",
    SW + "models/model_loader.py",
    SW + "models/resnet18/__init__.py",
    SW + "models/resnet18/resnet18.py",
    SW + "models/resnet18/resnet18_loader.py",
    SW + "models/vgg19/__init__.py",
    SW + "models/vgg19/vgg19.py",
    # FullAnalysis: the S-pass loop and the statistics (rows a1, a2, a9-a14); its module-level imports (KDEpy,
    # matplotlib, sacred) are not installed, so ref_arm.py exec's the method source inside a stub class
    SW + "train/results_analyzer.py",
    # converter and predictive entropy (rows a12, a16)
    HW + "converter/pytorch/Dropouts.py",
    HW + "converter/pytorch/nn2bnn.py",
    HW + "bayes_hw/metric_utils.py",
]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def main():
    if not os.path.isdir(REF_ROOT):
        print("make_ref: %s not present - keeping whatever oracle/_ref already holds" % REF_ROOT)
        return 0 if os.path.isdir(DST) else 1
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = sha256(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"reference_root": REF_ROOT, "files": manifest}, f, indent=1, sort_keys=True)
    print("make_ref: staged %d reference files under %s" % (len(FILES), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
