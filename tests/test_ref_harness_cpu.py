"""oracle/_ref + oracle/ref_harness.py: the CPU baseline is the reference as shipped (every file hashes to the
manifest written by oracle/make_ref.py) and the harness drives its own entry points.  Skips without oracle/_ref."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "oracle", "_ref")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "MANIFEST.json")),
                                reason="oracle/_ref absent (python oracle/make_ref.py in the build container)")


def test_ref_is_the_reference_as_shipped():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_ref

    assert make_ref.verify(REF)
    if os.path.isdir("/root/reference"):                 # build container: byte-identical to the checkout
        for rel in make_ref.FILES:
            with open(os.path.join("/root/reference", rel), "rb") as a, open(os.path.join(REF, rel), "rb") as b:
                assert a.read() == b.read(), rel
    # and it stays out of the repository's history
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=ROOT, capture_output=True, text=True).stdout
    assert tracked.strip() == ""


def test_harness_times_the_reference_layer():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_harness.py"), "--pairs", "1", "--hyps", "64",
                        "--corrs", "500", "--threads", "4", "--budget", "1"], capture_output=True, text=True, cwd=ROOT,
                       timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["kind"] == "reference" and d["value"] > 0 and d["dir"] == "oracle/_ref"
    assert d["verified"] == "sha256 of every file equals MANIFEST.json"


def test_reference_layer_agrees_with_the_oracle_port_on_the_same_noise():
    """The port (oracle/driver.test_loop) that the parity tests lean on, against the reference layer itself run from
    oracle/_ref with the same injected Gumbel noise: same winning score and inlier mask.  In fp64 (`-pr 2`): the
    reference's fp32 run loses genuine five-point models to LAPACK rounding and is not reproducible across thread
    counts, its fp64 run is."""
    code = r'''
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import ref_harness as rh
from differentiable_ransac_b200 import synth
from oracle import driver
torch.set_num_threads(4)
m, E, inl = synth.relative_pose_pair(600, 0.5, seed=3, noise=2e-4)
m = m.double()
w = synth.logits_regime(1, 600, "L0", seed=4)[0].double()
G = synth.gumbel_noise((48, 600), seed=7).double()
layer = rh.make_layer(48, precision=2)
layer.estimator.sampler.gumbel_dist.sample = lambda shape: G
K1, K2, im1, im2 = (t.double() for t in rh.intrinsics())
drv = layer.estimator
with torch.no_grad():
    model, mask, score, its = drv(m, w, K1, K2, None)
ref = driver.full_test_driver(m, w, [G], K1, K2, 0.75)
print("SCORES", float(score), float(ref[2]), int(mask.sum()), int(ref[1].sum()))
assert abs(float(score) - float(ref[2])) <= 1e-6 * float(ref[2])
assert int((mask != ref[1]).sum()) <= 1
''' % (ROOT, os.path.join(ROOT, "oracle"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert p.returncode == 0, (p.stdout[-500:], p.stderr[-2000:])
