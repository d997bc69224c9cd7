"""CPU oracle for the MoP attention hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a plain-PyTorch (CPU, eager, fp32/fp64)
restatement of the reference algorithm (Eran-BA/MoP, ``mop/models/*.py``),
written at the kernel boundary so that the CUDA path in ``mop_b200/`` can be
compared against it tensor by tensor.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product
package ``mop_b200`` never imports it: the product path has no CPU fallback
and raises if the CUDA library is missing.

Parity pinning: the reference's own tests hold no golden vectors for this path
(they assert shapes only).  The oracle is therefore pinned against
outputs of the *reference itself*, imported from ``/root/reference`` in the
authoring container by ``tests/golden/make_golden.py``; the resulting tensors
are committed under ``tests/golden/`` and ``tests/test_oracle_golden.py``
checks every oracle function against them on every run.
"""
