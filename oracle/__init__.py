"""CPU/GPU-agnostic restatement of the reference hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is imported by the product package ``nefii_b200``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / reported baseline.
"""
