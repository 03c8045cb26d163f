"""CPU oracle for the embedding -> interaction -> sparse-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the product package
(``deeplearningrecommendationsystem_b200``); only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or as the
timed CPU baseline -- never as the thing shipped.

It is a from-scratch fp32 CPU restatement (torch CPU ops for floating point, numpy for integer/index
work) of the arithmetic in /root/reference/model/*.py, generalised to N fields.  All of the reference's
arithmetic lives in the third-party dependency ``torch`` (un-pinned by the reference; 2.11.0 here), so
the oracle uses the same CPU primitives.  Parity pinning: the reference ships no tests or golden
vectors, so the oracle is pinned against outputs of the reference itself, run in
the build container by ``tests/golden/make_golden.py`` and committed under ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` checks every model: predictions, loss, every gradient, and the
parameters after two Adam steps).
"""
from . import interactions, ml100k, optim, index  # noqa: F401
