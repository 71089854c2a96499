"""
CPU oracle for the zopt LQR / iLQR / DDP / LQR-MPC hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a plain NumPy / torch-CPU (fp64) restatement of the reference's
algorithms, one function per reference function, each citing the reference
file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it -- and
there only as the checker or as the timed CPU baseline, never as the thing
shipped.  Nothing under `zopt_b200/` imports it; the product path raises when
the CUDA library is missing.

Parity pinning (see DESIGN.md "Oracle"):
  * The reference cannot be imported as-is here: `jax`, `jaxlib`, `cvxpy` are
    absent from the image and there is no network.
  * Every known-answer value the reference's own tests hold for this path
    (tests/test_lqrUtils.py:61-98, tests/test_ilqrUtils.py:7-22,56-81,110-135,
    167-196, tests/test_pytrees.py, tests/test_quadcopter.py:12-86) is asserted
    against this oracle in `tests/test_oracle_known_answers.py`.
  * In addition, `scripts/gen_golden_from_reference.py` executes the
    UNMODIFIED reference sources from /root/reference with the absent `jax`
    primitives substituted by NumPy/torch equivalents (`oracle/jax_shim`); its
    outputs are frozen under `tests/golden/` and the oracle is checked against
    them.  The formulas exercised are the reference's own; the primitives
    (`solve`, `eigh`, `scan`, autodiff) are stand-ins, which is stated wherever
    the fixtures are used.
  * `lqrMpc` (cvxpy -> OSQP) has no numeric assertion in the reference:
    parity unpinned for the bound-active case.

`oracle/c/zopt_oracle.c` (bound by `oracle/c_oracle.py`, built into
`oracle/_build/`) is the same kind of thing in plain C + OpenMP for the
headline path only -- what `bench.py` times on the host cores; it is pinned
against the same goldens and against this package (tests/test_c_oracle.py).
"""
