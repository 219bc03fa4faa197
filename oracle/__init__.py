"""TEST INFRASTRUCTURE ONLY -- CPU oracles for the DINO-MC head/loss/center/EMA hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker (or as the CPU arm that is timed
*beside* the GPU path).  The product package refuses to run without its CUDA library.

Contents
--------
np_oracle.py        float64 numpy restatement (closed-form forward AND hand-derived backward)
torch_port.py       loop-faithful functional restatement on torch CPU tensors (any dtype)
reference_loader.py imports the real reference from /root/reference (build container only)
gen_golden.py       writes tests/golden/*.npz from the real reference modules

Parity pinning: the reference repository has no tests, golden vectors or fixtures of its
own.  The oracles are therefore pinned against outputs of the
reference modules themselves, executed in the build container by ``gen_golden.py`` and
committed under ``tests/golden/`` (the generating script is committed next to them).
"""
