"""`RANSAC.__call__` honours plugin objects that are not this package's classes (ransac.py:63-76, 111 call
`sampler.sample`, `estimator.estimate_model`, `scoring.score` on whatever it was given): a user's estimator /
scoring / sampler -- or a subclass overriding a method -- is driven through its own methods by the generic chunked
loop, never silently replaced by the fused kernels.  Runs on the CPU: with foreign plugins the loop is torch only."""
import pytest
import torch

from differentiable_ransac_b200 import synth
from differentiable_ransac_b200.estimators.essential_matrix_estimator_nister import EssentialMatrixEstimatorNister
from differentiable_ransac_b200.ransac import RANSAC
from differentiable_ransac_b200.samplers.gumbel_sampler import GumbelSoftmaxSampler
from differentiable_ransac_b200.scorings.msac_score import MSACScore


class OneHotSampler:
    """Deterministic stand-in with the reference sampler's interface (gumbel_sampler.py:9-42)."""

    def __init__(self, batch_size, num_samples, idx):
        self.batch_size, self.num_samples, self.tau, self.calls, self.idx = batch_size, num_samples, 1.0, 0, idx

    def sample(self, logits=None, num_points=2000, selected=None):
        self.calls += 1
        ret = torch.zeros(self.batch_size, logits.shape[0])
        ret.scatter_(1, self.idx, 1.0)
        return ret, ret / self.num_samples


class GtEstimator:
    """Returns the ground truth for samples made of inliers only, garbage otherwise."""
    sample_size = 5

    def __init__(self, E, inl):
        self.E, self.inl, self.calls, self.shapes = E, inl, 0, []

    def estimate_model(self, matches, weights=None, **kw):
        self.calls += 1
        self.shapes.append(tuple(matches.shape))
        if matches.shape[1] != 5:                        # the final refit call (ransac.py:160-168): nothing better
            return None
        out = torch.eye(3).repeat(matches.shape[0], 1, 1)
        return out


class CountingScore:
    provides_inliers = True

    def __init__(self):
        self.calls = 0

    def score(self, matches, models, threshold=0.75):
        self.calls += 1
        x1 = torch.cat((matches[:, 0:2], torch.ones(matches.shape[0], 1)), 1)
        x2 = torch.cat((matches[:, 2:4], torch.ones(matches.shape[0], 1)), 1)
        r = torch.einsum("ni,mij,nj->mn", x2, models, x1).abs()
        masks = r < threshold
        return masks.float().sum(1) + torch.arange(models.shape[0]) * 1e-3, masks


def _problem():
    m, E, inl = synth.relative_pose_pair(200, 0.6, seed=2)
    Kc = torch.tensor([[800.0, 0, 320], [0, 800.0, 240], [0, 0, 1]])
    g = torch.Generator().manual_seed(0)
    idx = torch.stack([torch.randperm(200, generator=g)[:5] for _ in range(8)])
    return m, E, inl, Kc, idx


def test_foreign_plugins_are_called_not_replaced():
    m, E, inl, Kc, idx = _problem()
    smp, est, sc = OneHotSampler(8, 5, idx), GtEstimator(E, inl), CountingScore()
    drv = RANSAC(est, smp, sc, max_iterations=24, ransac_batch_size=8, sampler_id=2, threshold=0.75, adaptive=False)
    assert not drv.plugins_are_native()
    model, mask, score, its = drv(m, torch.ones(200) / 200, Kc, Kc, None)
    assert its == 24 and smp.calls == 3 and sc.calls == 3
    assert est.calls == 4 and est.shapes[0] == (8, 5, 4) and est.shapes[-1] == (1, 200, 4)     # 3 chunks + the final refit
    assert model.shape == (3, 3) and mask.dtype == torch.bool and mask.shape == (200,)
    assert float(score) == pytest.approx(float(mask.sum()) + 7e-3)          # arg-max picked the plugin's best-scored model


def test_train_mode_through_plugins_selects_the_slot_closest_to_gt():
    m, E, inl, Kc, idx = _problem()

    class TenSlots(GtEstimator):
        def estimate_model(self, matches, weights=None, **kw):
            self.calls += 1
            out = torch.randn(matches.shape[0], 10, 3, 3, generator=torch.Generator().manual_seed(self.calls))
            out[:, 3] = self.E                        # slot 3 is exactly the ground truth
            out[0, 5] = float("nan")                   # argmin lands on a NaN distance (as in the reference): sample dropped
            return out.reshape(-1, 3, 3)

    smp, est = OneHotSampler(8, 5, idx), TenSlots(E, inl)
    drv = RANSAC(est, smp, CountingScore(), train=True, max_iterations=16, ransac_batch_size=8, sampler_id=2)
    models, _, _, its = drv(m, torch.ones(200) / 200, Kc, Kc, E)
    assert its == 16 and sorted(models) == [0, 8]
    for chunk in models.values():               # ransac.py:87-108: NaN-chosen samples are filtered out of the chunk
        assert chunk.shape == (7, 3, 3) and torch.allclose(chunk, E.expand(7, 3, 3))


def test_subclass_overriding_a_method_counts_as_foreign_and_own_classes_as_native():
    class MyScore(MSACScore):
        def score(self, matches, models, threshold=0.75):
            raise AssertionError("never reached in this test")

    smp = GumbelSoftmaxSampler(8, 5, device="cpu")
    est = EssentialMatrixEstimatorNister("cpu")
    assert RANSAC(est, smp, MSACScore("cpu"), sampler_id=2).plugins_are_native()
    assert not RANSAC(est, smp, MyScore("cpu"), sampler_id=2).plugins_are_native()
    with pytest.raises(NotImplementedError):
        RANSAC(GtEstimator(None, None), smp, MSACScore("cpu"), sampler_id=2, lo=2)(torch.zeros(10, 4), torch.zeros(10),
                                                                                  torch.eye(3), torch.eye(3), None)
