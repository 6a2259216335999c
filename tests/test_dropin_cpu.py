"""The drop-in shim against the real reference checkout (build container only: the GPU box has no
/root/reference, so this test skips there).  Structural: construction, class substitution, checkpoint
compatibility.  Running the forward needs a GPU and the reference together, which no box offers."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = "/root/reference"

SCRIPT = r"""
import sys, types
from model_cl import *          # exactly what train.py:6 / test.py:3 do
from loss import *
import differentiable_ransac_b200.model_cl as ours
opt = types.SimpleNamespace(device='cuda', fmat=0, sampler=2, precision=1, tr=0, threshold=0.75,
                            ransac_batch_size=64, weighted=0)
model = DeepRansac_CLNet(opt)
assert isinstance(model.ransac_layer, ours.RANSACLayer), type(model.ransac_layer)
assert RANSACLayer is ours.RANSACLayer and RANSACLayer3D is ours.RANSACLayer3D
import differentiable_ransac_b200.loss as ol
assert MatchLoss is ol.MatchLoss and 'PoseLoss' in globals() and 'CLNet' in globals()
import differentiable_ransac_b200.cv_utils as oc
assert eval_essential_matrix is oc.eval_essential_matrix and recoverPose is oc.recoverPose and AUC is oc.AUC
assert 'evaluate_R_t' in globals() and 'denormalize_pts' in globals()       # the rest of cv_utils stays the reference's
sd = torch.load('/root/reference/pretrained_models/saved_model_5PC_l_epi/model.net', map_location='cpu')
missing, unexpected = model.load_state_dict(sd, strict=True), None
print('OK', len(sd), sum(p.numel() for p in model.parameters()))
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reference_scripts_namespace_with_shim():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), REF]))
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().startswith("OK 176 ")      # the checkpoint holds only ds_0.* tensors
