"""CPU oracle for the HAPPO/HATRPO on-policy hot path.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker (or the timed CPU baseline),
never as something the product path routes through.  ``harl_b200`` never imports it.

What it is: a functional restatement, in NumPy (integer / mask / scan work, strict fp32)
and CPU PyTorch fp32 (network maths and autograd), of the reference algorithm
PKU-MARL/HARL @ d539bad2 implements for the path BASELINE.json names.  Every function
cites the reference ``file:line`` it restates.

Pinning: the reference has no tests and no golden vectors of its own, so
the oracle is pinned against *outputs of the reference itself*: ``tests/golden/make_golden.py``
imports the unmodified reference from ``/root/reference`` in the build container, runs
its own functions on seeded inputs and commits the input/output vectors under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every oracle function
against those vectors.  Versions used to generate the vectors are recorded inside each
file (torch 2.11.0+cu128 CPU, numpy 2.3.5).
"""
