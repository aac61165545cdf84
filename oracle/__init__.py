"""CPU oracle for the MCHap hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``mchap_b200/`` may import this package.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
"""
