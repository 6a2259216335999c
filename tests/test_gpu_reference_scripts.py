"""The reference's own scripts, UNMODIFIED, on the B200 path through the drop-in shim:

  * `test.py` run as `__main__` exactly as its README says (`python test.py -m <checkpoint> -pth <data> -ds <scene>
    -fmat 0 -sam 2 ...`): its argparse, its `Dataset` loader reading synthetic pairs written in the loader's own
    .npy layout, `DeepRansac_CLNet` with the shipped 5PC checkpoint, `model(correspondences, K1, K2, im1, im2)`
    (test.py:38), `eval_essential_matrix` per pair, AUC;
  * `train.py`'s `train_step` (train.py:11-97) on one batch: forward, MatchLoss (`-w2 1`), backward into CLNet.

The scripts come from oracle/_ref (the byte-for-byte copy made by oracle/make_ref.py; hashed against its manifest
here), `dropin/` is first on PYTHONPATH, so `from model_cl import *` / `from loss import *` resolve to the shim and
RANSACLayer / MatchLoss / eval_essential_matrix are this package's.  Each runs in a child process (module names
like `utils`, `loss`, `test` must not leak into the test session)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "oracle", "_ref")
CKPT = os.path.join(REF, "pretrained_models", "saved_model_5PC_l_epi", "model.net")


@pytest.fixture(scope="module", autouse=True)
def _need():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if not os.path.exists(os.path.join(REF, "MANIFEST.json")):
        pytest.skip("oracle/_ref absent: run `python oracle/make_ref.py` in the build container")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_ref

    assert make_ref.verify(REF), "oracle/_ref differs from the reference as shipped"


def write_scene(folder, pairs, N, seed):
    """Synthetic image pairs in the on-disk layout datasets.py:36-50 reads: an object array
    (pts1 [1,N,2] pixels, pts2, ratios [1,N,1], im_size1, im_size2, K1, K2, gt_R, gt_t [3,1], size1, ang1, size2, ang2)."""
    from differentiable_ransac_b200 import synth

    os.makedirs(folder, exist_ok=True)
    K = np.array([[800.0, 0, 320.0], [0, 800.0, 240.0], [0, 0, 1.0]])
    gen = np.random.default_rng(seed)
    for i in range(pairs):
        m, E, inl, R, t = synth.relative_pose_pair(N, (0.5, 0.65, 0.8)[i % 3], seed=seed + i, noise=3e-4,
                                                   return_pose=True)
        m = m.double().numpy()
        p1 = (m[:, 0:2] * 800.0 + np.array([320.0, 240.0]))[None].astype(np.float32)
        p2 = (m[:, 2:4] * 800.0 + np.array([320.0, 240.0]))[None].astype(np.float32)
        ratios = np.where(inl.numpy(), 0.4, 0.7)[None, :, None].astype(np.float32)       # all pass the 0.8 filter
        one = np.ones((1, N, 1), dtype=np.float32)
        ang = gen.uniform(0, 180, (1, N, 1)).astype(np.float32)
        rec = np.empty(13, dtype=object)
        for j, v in enumerate((p1, p2, ratios, np.array([480, 640]), np.array([480, 640]), K.astype(np.float32),
                               K.astype(np.float32), R.numpy().astype(np.float32),
                               t.numpy().astype(np.float32).reshape(3, 1), one, ang, one * 1.1, ang)):
            rec[j] = v
        np.save(os.path.join(folder, f"pair_{i:03d}.npy"), rec, allow_pickle=True)


RUN_TEST_PY = r"""
import runpy, sys
sys.argv = ['test.py', '-m', {ckpt!r}, '-pth', {data!r}, '-ds', 'synth', '-bs', '4', '-fmat', '0', '-sam', '2',
            '-d', 'cuda', '-t', '0.75', '-nf', '{N}']
runpy.run_path({script!r}, run_name='__main__')
import model_cl, differentiable_ransac_b200.model_cl as ours
from differentiable_ransac_b200 import _lib
assert model_cl.RANSACLayer is ours.RANSACLayer and model_cl.DeepRansac_CLNet.__module__ == 'model_cl'
assert _lib._lib is not None, 'libdrb.so was never loaded'
print('SHIM_OK')
"""


def _env():
    return dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), REF, ROOT]),
                DRB_REFERENCE_DIR=REF)


def test_unmodified_test_py_runs_on_the_b200_path(tmp_path):
    N = 1000
    write_scene(str(tmp_path / "data" / "synth" / "test_data_rs"), pairs=8, N=N, seed=4000)
    code = RUN_TEST_PY.format(ckpt=CKPT, data=str(tmp_path / "data"), script=os.path.join(REF, "test.py"), N=N)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), env=_env(),
                       timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    out = r.stdout
    assert "SHIM_OK" in out and "AUC scores" in out and "Run time" in out
    auc = [float(x) for x in out.split("AUC scores = [")[1].split("]")[0].replace("np.float32(", "").replace(")", "").split(",")]
    # synthetic pairs with 50-80 % inliers: RANSAC finds the pose whatever the (untrained-on-this-data) weights say
    assert auc[1] >= 0.5, out[-800:]
    assert os.path.exists(os.path.join(str(tmp_path), "results"))           # test.py:104-109 wrote its summary


RUN_TRAIN_STEP = r"""
import sys, types
sys.modules.setdefault('tensorboardX', types.SimpleNamespace(SummaryWriter=object))     # train.py:8, logging only
import importlib.util, torch
spec = importlib.util.spec_from_file_location('ref_train', {script!r})
train = importlib.util.module_from_spec(spec); spec.loader.exec_module(train)
from datasets import Dataset
opt = train.create_parser('t').parse_args(['-fmat', '0', '-sam', '2', '-tr', '1', '-w2', '1', '-bs', '4', '-t', '0.75',
                                           '-nf', '{N}', '-rbs', '64'])
opt.device = torch.device('cuda:0')
opt.w = [opt.w0, opt.w1, opt.w2]
model = train.DeepRansac_CLNet(opt).to(opt.device)
model.load_state_dict(torch.load({ckpt!r}, map_location=opt.device))
model.train()
loader = torch.utils.data.DataLoader(Dataset([{data!r} + '/synth/test_data_rs/'], opt.snn, nfeatures=opt.nfeatures,
                                             fmat=opt.fmat), batch_size=4, shuffle=False)
loss_fn = [train.PoseLoss(opt.fmat), train.ClassificationLoss(opt.fmat), train.MatchLoss(opt.fmat)]
batch = next(iter(loader))
loss, Es = train.train_step(batch, model, opt, loss_fn)
loss.backward()
g = [p.grad for p in model.parameters() if p.grad is not None]
assert len(g) > 0 and all(torch.isfinite(x).all() for x in g) and any(float(x.abs().max()) > 0 for x in g)
import differentiable_ransac_b200.loss as ol
assert train.MatchLoss is ol.MatchLoss
print('TRAIN_STEP_OK', float(loss), len(Es), tuple(Es[0].shape))
"""


def test_unmodified_train_step_backpropagates_into_clnet(tmp_path):
    N = 1000
    write_scene(str(tmp_path / "data" / "synth" / "test_data_rs"), pairs=4, N=N, seed=5000)
    code = RUN_TRAIN_STEP.format(ckpt=CKPT, data=str(tmp_path / "data"), script=os.path.join(REF, "train.py"), N=N)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), env=_env(),
                       timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "TRAIN_STEP_OK" in r.stdout
    loss = float(r.stdout.split("TRAIN_STEP_OK")[1].split()[0])
    assert 0.0 <= loss <= 1.0                                                # mean of min(episym, 1)
