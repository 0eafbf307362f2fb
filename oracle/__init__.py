"""CPU oracle for the hm-vae hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a checker: a CPU (torch fp32 / pure-Python int)
restatement of the reference algorithms, pinned against golden vectors that were
produced by running the real reference modules (``oracle/make_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing in ``hm_vae_b200/`` does.
"""
