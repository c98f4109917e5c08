"""TEST INFRASTRUCTURE ONLY - CPU oracle for the cae_tools hot path.

Nothing under ``oracle/`` is part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker.

Parity status: the reference's own tests contain no assertions and no golden vectors,
so the oracle is pinned by DIFFERENTIAL runs of the live reference modules in the build container:
``oracle/gen_golden.py`` imports the unmodified reference (``/root/reference``) and writes the fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks the restatement against them.  The arithmetic itself
lives in a third-party dependency of the reference that is not vendored in it - PyTorch (unpinned by the
reference; effective pin = this image's torch 2.11.0) - see ``oracle/numpy_ops.py`` for the published
definitions restated independently of torch.
"""
