"""Recipe that puts the UNMODIFIED reference on the GPU box: oracle/_ref/.

TEST / BASELINE INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, and the reference is pure
Python (nothing to compile), so its "build" is a byte-for-byte copy of the files on and next to the hot path
(SURVEY 8a + the scripts that drive them) into oracle/_ref/, which is git-ignored (the sources never enter this
repository's history) but NOT gpurun-ignored (it travels with the snapshot like a built .so).  A MANIFEST.json
with the sha256 of every file is written beside them; `verify()` re-hashes, so a test on the GPU box can show
that what it ran is the reference as shipped.

    python oracle/make_ref.py            # in the build container (also called by __graft_entry__.build())

Users: bench.py's `cpu_baseline` / `--impl reference` (oracle/ref_harness.py times `RANSACLayer.forward` of THIS
code on the host cores), tests/test_gpu_reference_scripts.py (runs the reference's own test.py / train.py
functions on the B200 path through dropin/), tests/test_ref_harness_cpu.py.  The product never imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DRB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

# the hot path, what it star-imports, the scripts that drive it, and the shipped 5PC checkpoint
FILES = [
    "ransac.py", "model_cl.py", "loss.py", "cv_utils.py", "utils.py", "math_utils.py", "feature_utils.py",
    "datasets.py", "test.py", "train.py", "train_point.py",
    "estimators/essential_matrix_estimator_nister.py", "estimators/essential_matrix_estimator_stewenius.py",
    "estimators/fundamental_matrix_estimator.py", "estimators/rigid_transformation_SVD_based_solver.py",
    "samplers/gumbel_sampler.py", "samplers/uniform_sampler.py", "scorings/msac_score.py",
    "pretrained_models/saved_model_5PC_l_epi/model.net", "LICENSE",
]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def make(src=SRC, dst=DST):
    """Copy FILES from `src` to `dst` unchanged and write the manifest.  Returns dst, or None when there is no
    reference checkout (the GPU box: the directory made in the build container is used as is)."""
    if not os.path.isdir(src):
        return None
    manifest = {}
    for rel in FILES:
        a, b = os.path.join(src, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(b), exist_ok=True)
        if not (os.path.exists(b) and _sha(a) == _sha(b)):
            if os.path.exists(b):
                os.chmod(b, 0o644)
            shutil.copyfile(a, b)
        manifest[rel] = _sha(b)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump(dict(source="weitong8591/differentiable_ransac (unmodified files)", sha256=manifest), f, indent=1)
    return dst


def available(dst=DST):
    return os.path.exists(os.path.join(dst, "MANIFEST.json"))


def verify(dst=DST):
    """True when every file of the manifest is present with its recorded hash."""
    with open(os.path.join(dst, "MANIFEST.json")) as f:
        manifest = json.load(f)["sha256"]
    return all(os.path.exists(os.path.join(dst, rel)) and _sha(os.path.join(dst, rel)) == h
               for rel, h in manifest.items())


if __name__ == "__main__":
    out = make()
    print("oracle/_ref:", out if out else f"no reference checkout at {SRC}; left as is")
    sys.exit(0 if (out or available()) else 1)
