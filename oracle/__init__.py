"""CPU oracle for the uam_path_planning hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / CPU baseline.  The
product package ``uam_path_planning_b200`` never imports this package and
fails loudly when its CUDA library is missing.
"""
