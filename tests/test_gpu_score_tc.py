"""GPU tests of the EXPERIMENTAL tensor-core soft-MSAC scorer (drb_score_msac_tc, csrc/score_tc.cu) against the CPU
oracle (oracle/scoring.py <- scorings/msac_score.py:12-55) and the FP32 kernel (drb_score_msac).

The kernel was written after round 1's GPU budget was spent and has never run on hardware, so these tests
are opt-in: set DRB_EXPERIMENTAL=1.  Each case runs in a CHILD process under a timeout -- a wrong mbarrier
phase in a warp-specialised kernel is a hang, not an exception, and must not take the test session (or the
GPU box) with it.  Its host model is tested on the CPU in test_host_math.py::test_msac_tc_*."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module", autouse=True)
def _need_gpu_and_opt_in():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if os.environ.get("DRB_EXPERIMENTAL", "0") != "1":
        pytest.skip("experimental kernel: set DRB_EXPERIMENTAL=1")


CHILD = r"""
import sys, torch
sys.path.insert(0, {root!r})
from differentiable_ransac_b200 import ops, synth
from oracle import scoring
B, M, N, counts = {case!r}
matches, _, _ = synth.relative_pose_batch(B, max(N, 8), seed=5 + N)
matches = matches[:, :N].contiguous()
gen = torch.Generator().manual_seed(M)
models = torch.randn(B, M, 3, 3, generator=gen)
models = models / models.flatten(-2).norm(dim=-1)[..., None, None]
thr = torch.rand(B, generator=gen) * 0.05 + 0.002
count = None if counts is None else torch.tensor(counts, dtype=torch.int32)
ids = torch.stack([torch.randperm(4 * M + 7, generator=gen)[:M] for _ in range(B)]).int()
dev = "cuda"
args = (matches.to(dev), models.to(dev), thr.to(dev))
kw = dict(count=None if count is None else count.to(dev), ids=ids.to(dev))
s_tc, b_tc = ops.score_msac(*args, kernel="tc", **kw)
s_again, b_again = ops.score_msac(*args, kernel="tc", **kw)
s_ref, b_ref = ops.score_msac(*args, kernel="block", **kw)
torch.cuda.synchronize()
assert torch.equal(s_tc.isnan(), s_again.isnan())
for b in range(B):
    c = M if count is None else int(count[b])
    if c == 0:
        assert int(b_tc[b]) == 0
        continue
    want, _ = scoring.msac_score(matches[b].double(), models[b, :c].double(), float(thr[b]))
    got = s_tc[b, :c].cpu().double()
    rel = (got - want).abs() / want.clamp_min(1.0)
    assert rel.max() < 1e-4, (b, float(rel.max()))
    assert torch.equal(s_tc[b, :c], s_again[b, :c]), "not deterministic"
    key = int(b_tc[b]) & 0xFFFFFFFFFFFFFFFF
    best_id = 0xFFFFFFFF - (key & 0xFFFFFFFF)
    pos = int((ids[b, :c] == best_id).nonzero()[0])
    assert got[pos] >= got.max() - 1e-6 * max(1.0, float(got.max()))
print("OK")
"""

CASES = [
    (1, 1, 1, None),
    (1, 5, 3, None),
    (2, 33, 64, None),
    (3, 70, 257, [70, 0, 41]),
    (3, 300, 2500, [300, 0, 129]),
    (4, 1000, 2000, [1000, 517, 1, 32]),
    (32, 4400, 2000, None),        # the headline shape: more units than SMs, every ring wraps many times
    (40, 200, 500, None),
]


@pytest.mark.parametrize("case", CASES)
def test_tc_kernel_matches_oracle(case):
    r = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, case=case)], capture_output=True, text=True,
                       timeout=180)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (r.stdout[-500:], r.stderr[-2000:])
