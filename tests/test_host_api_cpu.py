"""The host mirror keeps the reference's plugin surface: class names, constructor
arguments, attributes.  CPU only -- nothing here launches a kernel."""
import inspect
import types

import pytest
import torch


def _opt(**kw):
    d = dict(device="cuda", fmat=0, sampler=2, precision=1, tr=0, threshold=0.75, ransac_batch_size=64, weighted=0)
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_plugin_classes_and_signatures():
    from differentiable_ransac_b200.estimators.essential_matrix_estimator_nister import EssentialMatrixEstimatorNister
    from differentiable_ransac_b200.estimators.essential_matrix_estimator_stewenius import EssentialMatrixEstimator
    from differentiable_ransac_b200.estimators.fundamental_matrix_estimator import FundamentalMatrixEstimatorNew
    from differentiable_ransac_b200.estimators.rigid_transformation_SVD_based_solver import RigidTransformationSVDBasedSolver
    from differentiable_ransac_b200.ransac import RANSAC, RANSAC3D
    from differentiable_ransac_b200.samplers.gumbel_sampler import GumbelSoftmaxSampler
    from differentiable_ransac_b200.scorings.msac_score import MSACScore

    assert EssentialMatrixEstimatorNister("cuda").sample_size == 5
    assert EssentialMatrixEstimator("cuda").sample_size == 5
    assert FundamentalMatrixEstimatorNew("cuda", 0).sample_size == 7
    assert RigidTransformationSVDBasedSolver().sample_size == 3
    assert MSACScore("cuda").provides_inliers is True
    smp = GumbelSoftmaxSampler(64, 5, tau=1.0, device="cuda", data_type=torch.float32)
    assert (smp.batch_size, smp.num_samples, smp.tau) == (64, 5, 1.0)
    ref_args = ["estimator", "sampler", "scoring", "fmat", "train", "ransac_batch_size", "sampler_id", "weighted",
                "threshold", "confidence", "max_iterations", "lo", "lo_iters", "eps"]          # ransac.py:8-24
    got = list(inspect.signature(RANSAC.__init__).parameters)[1:]
    assert got[: len(ref_args)] == ref_args
    assert list(inspect.signature(RANSAC.__call__).parameters)[1:] == ["matches", "logits", "K1", "K2", "gt_model"]
    assert list(inspect.signature(RANSAC3D.__call__).parameters)[1:] == ["matches", "logits", "gt_model", "valid"]
    nister_args = ["matches", "weights", "K1", "K2", "inlier_indices", "best_model", "unnormalzied_threshold", "best_score"]
    assert list(inspect.signature(EssentialMatrixEstimatorNister.estimate_model).parameters)[1:] == nister_args


def test_layers_read_the_reference_opt_namespace():
    from differentiable_ransac_b200.model_cl import RANSACLayer, RANSACLayer3D

    lay = RANSACLayer(_opt())
    assert lay.estimator.max_iterations == 5000 and lay.estimator.sampler.num_samples == 5        # model_cl.py:213-219
    assert RANSACLayer(_opt(tr=1)).estimator.max_iterations == 100
    f = RANSACLayer(_opt(fmat=1, sampler=3, tr=1))
    assert f.estimator.max_iterations == 1000 and f.estimator.sampler.num_samples == 8
    assert sum(p.numel() for p in lay.parameters()) == 0
    assert RANSACLayer3D(_opt(tr=1)).estimator.sampler.num_samples == 3
    with pytest.raises(NotImplementedError):
        RANSACLayer(_opt(sampler=0))


def test_threshold_normalisation_quirk():
    from differentiable_ransac_b200.ransac import normalized_threshold

    K1 = torch.tensor([[800.0, 0, 320], [0, 820.0, 240], [0, 0, 1]])
    K2 = torch.tensor([[700.0, 0, 320], [0, 900.0, 240], [0, 0, 1]])
    assert abs(normalized_threshold(0.75, K1, K2, False) - 0.75 / ((800 + 820 + 800 + 900) / 4)) < 1e-12   # ransac.py:52
    assert normalized_threshold(0.75, K1, K2, True) == 0.75


def test_no_cpu_fallback():
    from differentiable_ransac_b200 import _lib, ops

    with pytest.raises(_lib.DrbError):
        ops.sample(torch.zeros(1, 16), 4, 5)


def test_service_scorer_selection_logic(monkeypatch):
    """engine.service_scorer / tc_scorer_agrees with a stand-in for the CUDA scorers: "auto" adopts the first
    tensor-core variant whose scores agree with the FP32 kernel's, falls back to "block" (with a warning) when none
    does, never second-guesses an explicit request, and keeps batches of more than 1024 pairs on "block"."""
    import types
    import warnings

    import torch

    from differentiable_ransac_b200 import engine

    def reference_scores(matches, models, thr):
        B, N, _ = matches.shape
        m = models.reshape(B, -1, 3, 3)
        one = torch.ones(B, N, 1)
        x1, x2 = torch.cat([matches[..., :2], one], -1), torch.cat([matches[..., 2:], one], -1)
        mx = torch.einsum("bmij,bnj->bmni", m, x1)
        mtx = torch.einsum("bmji,bnj->bmni", m, x2)
        r = (mx * x2[:, None]).sum(-1)
        j = (mx[..., :2] ** 2).sum(-1) + (mtx[..., :2] ** 2).sum(-1)
        t = (1.5 * thr)[:, None, None] ** 2
        return (1 - (r * r / j) / t).clamp(0, 1).sum(-1)

    def stand_in(bad):
        def score_msac(matches, models, thr, kernel=None, **kw):
            s = reference_scores(matches, models, thr)
            return s * (1.01 if kernel in bad else 1.0 + 2e-5 if kernel != "block" else 1.0), None
        return types.SimpleNamespace(score_msac=score_msac)

    monkeypatch.setattr(engine, "SERVICE_SCORER", "auto")
    for bad, want in ((set(), "tc_bf16p_s"), ({"tc_bf16p_s"}, "tc_bf16p"), ({"tc_bf16p_s", "tc_bf16p"}, "tc_bf16"),
                      ({"tc_bf16p_s", "tc_bf16p", "tc_bf16"}, "tc_tf32"),
                      ({"tc_bf16p_s", "tc_bf16p", "tc_bf16", "tc_tf32"}, "block")):
        monkeypatch.setattr(engine, "ops", stand_in(bad))
        monkeypatch.setattr(engine, "_TC_CHECKED", {})
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            assert engine.service_scorer("cpu", 32) == want
        assert bool(caught) == (want == "block")
    monkeypatch.setattr(engine, "ops", stand_in(set()))
    monkeypatch.setattr(engine, "_TC_CHECKED", {})
    assert engine.service_scorer("cpu", 32, "stream") == "stream"          # explicit request
    assert engine.service_scorer("cpu", 2000) == "block"                   # more pairs than the unit table holds
    monkeypatch.setattr(engine, "SERVICE_SCORER", "block")
    assert engine.service_scorer("cpu", 32) == "block"
