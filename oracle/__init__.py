"""CPU oracle for the VI-model-1 hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Everything under ``oracle/`` is a CPU restatement (plain PyTorch fp32/fp64 on the host) of the
reference algorithm in iacercalixto/variational_mmt, plus the tooling that pins that restatement to
the *executed* reference (``make_golden.py`` imports ``/root/reference`` unmodified, with external
shims, and writes the fixtures in ``tests/golden/``).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package, and only as the checker / reported CPU baseline.  The product
package ``variational_mmt_b200`` never imports it and has no CPU fallback.

Parity pinning: the reference repository ships no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself run in the
build container: ``tests/golden/*.npz`` (generator script: ``oracle/make_golden.py``).
"""
