"""Host-model precision sweep of the tensor-core scorer (tests/hostcheck: images -> descriptors -> MMA steps -> epilogue)
against the fp64 oracle over inlier ratios, noise levels, coordinate scales and thresholds; CPU only.
Output of the run committed as profiles/r1_tc_precision_sweep.txt."""
import sys, ctypes, numpy as np, torch
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,os.path.join(ROOT,'tests')); sys.path.insert(0,ROOT)
import hostcheck
from differentiable_ransac_b200 import synth
from oracle import nister, scoring
from helpers import trace_constraint_residual
lib = hostcheck.load()
vp=lambda a:a.ctypes.data_as(ctypes.c_void_p)
def run(matches, E, thr, words):
    m=np.ascontiguousarray(matches.numpy().astype(np.float32)); md=np.ascontiguousarray(E.reshape(-1,9).numpy().astype(np.float32))
    out=np.zeros(md.shape[0],dtype=np.float32); l=ctypes.c_int()
    lib.hc_msac_tc_scores(vp(m),m.shape[0],vp(md),md.shape[0],ctypes.c_float(thr),words,vp(out),ctypes.byref(l))
    return torch.from_numpy(out)
N,K=1500,60
for seed in range(4):
    for rho,noise,scale,thr_px in ((0.2,1e-3,1.0,0.75),(0.6,2e-4,1.0,0.75),(0.4,5e-4,2.5,0.75),(0.4,5e-4,1.0,3.0)):
        matches,_,_=synth.relative_pose_batch(3,N,seed=10+seed,noise=noise)
        b = {0.2:0,0.4:1,0.6:2}[rho]
        mt = matches[b]*scale
        g=torch.Generator().manual_seed(seed)
        idx=torch.stack([torch.randperm(N,generator=g)[:5] for _ in range(K)])
        inl=torch.arange(N-int(rho*N),N); idx[:K//3]=inl[torch.stack([torch.randperm(len(inl),generator=g)[:5] for _ in range(K//3)])]
        E=nister.five_point(mt[idx].double()); E=E[trace_constraint_residual(E)<1e-8].float()
        thr=thr_px/800*scale
        want,_=scoring.msac_score(mt.double(),E.double(),thr)
        fp32,_=scoring.msac_score(mt,E,thr)
        res=[]
        for w in (2,3,18):
            got=run(mt,E,thr,w)
            rel=(got.double()-want).abs()/want.clamp_min(1)
            res.append(float(rel.max()))
        r32=float(((fp32.double()-want).abs()/want.clamp_min(1)).max())
        print(f"seed {seed} rho {rho} noise {noise} scale {scale} thr_px {thr_px}: models {len(want)} best {float(want.max()):.1f}  tf32x3 {res[0]:.2e}  bf16x6 {res[1]:.2e}  tf32x3+pair {res[2]:.2e}  fp32 {r32:.2e}  argmax ok {int(run(mt,E,thr,2).argmax())==int(want.argmax())}")
