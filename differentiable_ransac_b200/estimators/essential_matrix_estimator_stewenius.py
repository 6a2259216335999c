"""EssentialMatrixEstimator (Stewenius) -- host mirror of
`estimators/essential_matrix_estimator_stewenius.py:5-80`.

Stewenius' action-matrix solver and Nister's solve the SAME ten-solution polynomial system; the
reference's output differs only by an arbitrary per-model scale (the eigenvector normalisation of
`linalg.eig`, :74-78) and by bogus real parts of complex eigenvectors.  This class therefore shares
the CUDA 5-point kernel and returns unit-norm models (parity: tests/test_gpu_parity.py::
test_stewenius_solutions_contained).  Unlike the reference class it takes `device`
and the refit keyword arguments."""
from .essential_matrix_estimator_nister import EssentialMatrixEstimatorNister


class EssentialMatrixEstimator(EssentialMatrixEstimatorNister):
    pass
