"""CPU oracle for the GaPro pseudo-label generator path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gapro_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or the timed CPU baseline — never as the product path.

PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or
fixtures for this path and its arithmetic lives in two
un-vendored third-party dependencies that are not installable here
(``gpytorch`` — version not pinned by the reference, restated at 1.8.1-style
constants — and ``torch_scatter==2.0.9``).  The oracle is therefore a
restatement of `/root/reference/gapro/gen_ps_utils.py`,
`/root/reference/gapro/gaussian_process_utils.py` and of those libraries'
published algorithms, validated only against itself (autograd vs. hand-derived
gradients, torch ops vs. explicit loops) and against hand-built known-answer
cases in ``tests/``.
"""
