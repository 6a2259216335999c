"""EssentialMatrixEstimator (Stewenius) -- host mirror of
`estimators/essential_matrix_estimator_stewenius.py:5-80`.

The reference's two five-point classes solve the SAME polynomial system -- the trace and determinant constraints on
E = x X + y Y + z Z + W over the 4-dimensional null space of the 5 x 9 epipolar system -- and differ only in how the
ten solutions are extracted: Nister eliminates to a degree-10 polynomial in z (nister.py:155-176, 355-370), Stewenius
reads them off the eigenvectors of a 10 x 10 action matrix (stewenius.py:58-78).  Run in fp64, the two reference
classes return the same set of essential matrices, model for model (tests/golden/stewenius_64.npz, made by the
reference itself: 224 of 224 genuine models coincide to 1e-8 up to scale and sign,
tests/test_oracle_golden.py::test_stewenius_fp64_and_the_two_classes_define_the_same_models).  What differs in the
reference's output is not a model:
  * scale: the eigenvector normalisation of `torch.linalg.eig` (:74-78) -- this class returns unit-norm models;
  * the slots of complex eigenvalues hold `vv.real` (:77), which satisfies no constraint -- this class
    returns the identity there and `last_nsol` says how many slots are genuine.
A general non-symmetric 10 x 10 eigen-decomposition per hypothesis buys nothing on the device: the eigenvalues ARE
the roots the Sturm isolation finds, and the eigenvector's last four entries ARE the back-substituted (x, y, z, 1).
So the class shares the CUDA five-point kernel (`drb_solve_e5`); the device path is pinned to the fp64 Stewenius
models by tests/test_gpu_parity.py::test_stewenius_fp64_solution_set_is_found (98.7 % within 1e-3, median 8e-7; the
reference's own fp32 run of the class: 99.6 % within 1e-3 -- eig copes better with the few near-double roots --
median 4e-6).  Unlike the reference class this one
takes `device` and the refit keyword arguments, so the whole test-mode `RANSAC.__call__` runs with it."""
from .essential_matrix_estimator_nister import EssentialMatrixEstimatorNister


class EssentialMatrixEstimator(EssentialMatrixEstimatorNister):
    """Same call shapes as the reference class; see the module docstring for why it shares the kernel."""
