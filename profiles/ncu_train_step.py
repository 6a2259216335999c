"""A few eager training steps of a config (default cfg5: 5PC-E, 32 x 1000 x 2000) for ncu captures of the training
kernels:  ncu --set full -k regex:sample_train_kernel -s 2 -c 1 ... python profiles/ncu_train_step.py [cfg5|cfg3|cfg4]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from differentiable_ransac_b200 import engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
cfg = bench.CONFIGS[name]
dev = "cuda"
data = bench.make_train_inputs(cfg, cfg["B"], seed=300)
P = None if data["pts"] is None else data["pts"].shape[1]
step = engine.TrainStep(cfg["kind"], cfg["B"], cfg["N"], cfg["K"], dev, P=P, seed=5, graph=False)
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
for _ in range(4):
    step.run(d["matches"], d["logits"], d.get("gt"), d["pts"], d["npts"])
torch.cuda.synchronize()
print("ok", float(step.loss))
