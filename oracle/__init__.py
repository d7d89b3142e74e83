"""CPU oracle for the GaPro pseudo-label generator path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gapro_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or the timed CPU baseline — never as the product path.

PINNING.  The scene-level restatement (`gen_ps_oracle.py`, `heuristic_oracle.py`)
is pinned to the reference's own code: `tests/golden/make_ref_golden.py` imports
`/root/reference/gapro/gen_ps_utils.py` and `eval_ps_labels.py` as they lie,
executes them on CPU (stand-ins only for the two absent dependencies: a
sequential `torch_scatter` shim and `fit_gp_spp` -> this oracle's GP fit) and
commits their outputs as `tests/golden/ref_outputs.npz`;
`tests/test_reference_run.py` requires the oracle to reproduce them bit for bit.

PARITY UNPINNED for the inside of the GP fit (`gp_oracle.py`): its arithmetic
lives in `gpytorch` (version not pinned by the reference, restated at
1.8.1-style constants), which is not installable here, and the reference ships
no golden vectors for it.  That part is a restatement of
`/root/reference/gapro/gaussian_process_utils.py` and of gpytorch's published
algorithm, validated against itself (autograd vs. hand-derived gradients,
torch ops vs. explicit loops) and against independently written textbook
formulas of the same model (unwhitened SVGP predictive, dense-Gaussian KL,
adaptive quadrature of the expected log-likelihood, torch.optim.Adam) in
tests/test_oracle_gp.py - not against gpytorch's own float32 execution.
"""
