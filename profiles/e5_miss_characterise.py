"""Which of the fp64 oracle's genuine five-point models does the fp32 solver miss, and does it matter for the winner?

TEST / PROFILING INFRASTRUCTURE (CPU only): the device math headers compiled for the host (tests/hostcheck, the
kernels' exact fp32 arithmetic) against oracle/nister.py in fp64 (nister.py:69-408) on cfg2's own inputs
(bench.make_inputs, 2000 correspondences, minimal samples drawn from softmax(logits) without replacement = the
law of drb_sample_sets).  Per pair: every genuine oracle model (trace-constraint residual < 1e-8) is matched to
the closest model of the same sample up to sign; a miss (> 1e-3) is classified by
  * the conditioning of its root: |P'(z)| relative to the coefficient scale (near-double roots),
  * |z| (the reversed-polynomial domain |z| > 1),
  * its fp64 MSAC score (does a missed model ever win?).
Prints one JSON line per pair and a summary line.

    python profiles/e5_miss_characterise.py [pairs] [K]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import hostcheck  # noqa: E402
from helpers import trace_constraint_residual  # noqa: E402
from oracle import nister, scoring  # noqa: E402


def host_solve(pts32, polish=2):
    lib = hostcheck.load()
    K = pts32.shape[0]
    models = np.zeros((K, 10, 9), dtype=np.float32)
    nsol = np.zeros(K, dtype=np.int32)
    p = np.ascontiguousarray(pts32.numpy(), dtype=np.float32)
    lib.hc_e5_solve_f32(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(K), models.ctypes.data_as(ctypes.c_void_p),
                        nsol.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(polish))
    return torch.from_numpy(models).reshape(K, 10, 3, 3), torch.from_numpy(nsol)


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    N = 2000
    matches, logits, thr, _ = bench.make_inputs(pairs, N, seed=1234)
    tot = dict(genuine=0, missed=0, missed_1e2=0, same_best=0, tie_1e4=0, worse=0)
    for b in range(pairs):
        gen = torch.Generator().manual_seed(77 + b)
        p = torch.softmax(logits[b], 0)
        idx = torch.stack([torch.multinomial(p, 5, replacement=False, generator=gen).sort().values for _ in range(K)])
        pts = matches[b][idx]                                        # [K,5,4]
        m64 = nister.five_point(pts.double()).reshape(K, 10, 3, 3)
        ours, nsol = host_solve(pts)
        live = torch.arange(10)[None] < nsol[:, None]
        genuine = (trace_constraint_residual(m64.reshape(-1, 3, 3)) < 1e-8).reshape(K, 10)
        c = torch.where(live[..., None, None], ours.double(), torch.full_like(ours.double(), 1e3)).flatten(2)
        r = m64.flatten(2)
        r = r / r.norm(dim=-1, keepdim=True)
        d = torch.minimum((c[:, :, None] - r[:, None]).norm(dim=-1), (c[:, :, None] + r[:, None]).norm(dim=-1)).min(1).values
        miss = genuine & (d > 1e-3)
        s64, _ = scoring.msac_score(matches[b].double(), m64.reshape(-1, 3, 3), float(thr[b]))
        s64 = torch.where(genuine.flatten(), s64, torch.full_like(s64, -1.0))
        so, _ = scoring.msac_score(matches[b].double(), ours.reshape(-1, 3, 3).double(), float(thr[b]))
        so = torch.where(live.flatten(), so, torch.full_like(so, -1.0))
        best64, bestours = int(s64.argmax()), int(so.argmax())
        same = best64 // 10 == bestours // 10
        rel = abs(float(so.max()) - float(s64.max())) / float(s64.max())
        top_missed = float(s64[miss.flatten()].max()) if miss.any() else 0.0
        line = dict(pair=b, genuine=int(genuine.sum()), ours=int(nsol.sum()), missed_1e3=int(miss.sum()),
                    missed_1e2=int((genuine & (d > 1e-2)).sum()), median_dist=float(d[genuine].median()),
                    best_oracle=float(s64.max()), best_ours=float(so.max()), same_best_hyp=bool(same), rel_best=rel,
                    best_score_among_missed=top_missed,
                    missed_share_of_high_scores=float((miss.flatten() & (s64 > 0.5 * s64.max())).sum()) /
                    max(1.0, float((s64 > 0.5 * s64.max()).sum())))
        print(json.dumps(line), flush=True)
        tot["genuine"] += line["genuine"]
        tot["missed"] += line["missed_1e3"]
        tot["missed_1e2"] += line["missed_1e2"]
        tot["same_best"] += int(same)
        tot["tie_1e4"] += int((not same) and rel <= 1e-4)
        tot["worse"] += int((not same) and rel > 1e-4)
    tot["pairs"] = pairs
    tot["miss_rate_1e3"] = tot["missed"] / max(1, tot["genuine"])
    print(json.dumps(tot))


if __name__ == "__main__":
    main()
