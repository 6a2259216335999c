"""CPU oracle for the hypothesize-and-score path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU with torch / LAPACK arithmetic, the
algorithms of the reference (weitong8591/differentiable_ransac) that sit on the
hot path of SURVEY.md section 8(a).  It exists to check the CUDA kernels in
`differentiable_ransac_b200/csrc` and to provide the CPU baseline leg of
`bench.py`.  Nothing in the product package may import it: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs do.

Pinning: the oracle is checked against the reference ITSELF, imported from
`/root/reference` in the build container by `tests/golden/make_golden.py`,
which commits the reference's outputs on seeded inputs as fixtures under
`tests/golden/*.npz`.  `tests/test_oracle_golden.py` compares every oracle
function with those fixtures, so the oracle is pinned wherever the reference
itself is well defined (see DESIGN.md section "Parity contract" for the places
where the reference is not self-consistent: null-space gauge, root order,
complex-root slots, the broken 7-point path).

Each function cites the reference file:line it follows.
"""

from . import sampler, nister, stewenius, fundamental, rigid, scoring, driver  # noqa: F401
