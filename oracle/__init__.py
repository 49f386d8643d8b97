"""CPU oracle for the LANTERN verification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``torch_gpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the timed CPU baseline.  ``lantern_b200`` never imports it.
"""
