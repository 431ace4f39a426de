"""TEST INFRASTRUCTURE - CPU oracle of the DeepCTR hot path (see oracle/ops.py header).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package; the product package ``deepctr_b200`` never does.
"""
