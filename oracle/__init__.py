"""CPU oracle for the W4A8KV4 hot path of mit-han-lab/omniserve.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under
``omniserve_b200/`` imports it; the product path fails loudly when the CUDA
library is missing instead of falling back to anything here.

PARITY PINNING.  The reference holds no golden vectors / KATs for this path
(``unit_tests/`` has one stale script with no asserts), and its
kernels are CUDA-only so they cannot be run in the authoring container.  The
restatement in this package is therefore pinned two ways:

1. structural: the packers below are checked against a *replay of the
   reference's own reshape/permute chains* (``tests/test_oracle_w4a8.py``
   re-derives ``w4a8_linear.py:295-335`` index-for-index) and against
   fixtures under ``tests/golden/`` produced by ``tests/golden/make_golden.py``
   importing the reference's Python packer/quantiser;
2. numerical: on the GPU box the reference's own kernels rebuilt for sm_100 by
   ``oracle/build_ref.py`` (``oracle/_ref/omniserve_backend/*.so``) are run on
   the same seeded inputs (``tests/test_gpu_vs_ref.py``) -- that is "outputs of
   the reference itself run here".

Until (2) has been observed green on a GPU box the status is "parity unpinned";
DESIGN.md records the current status.
"""
