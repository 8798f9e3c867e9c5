"""CPU oracle for the VidSGG-BIG per-video relation hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (torch-CPU / numpy / plain
Python loops) of the reference algorithms on the hot path.  It
exists to *check* the CUDA path and to be timed as the ``cpu_baseline`` leg of ``bench.py``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; nothing under ``vidsgg_big_b200/`` does (a test
greps for that).

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the *unmodified reference itself*, imported from
``/root/reference`` in the build container by ``tests/golden/make_golden.py`` and committed
as ``tests/golden/*.npz`` (inputs/weights are regenerated from numpy seeds), plus the
hand-computed known answers of SURVEY.md §8c.  ``tests/test_oracle_golden.py`` replays them.

Every function cites the reference ``file:line`` it follows (paths relative to the
reference root).
"""
