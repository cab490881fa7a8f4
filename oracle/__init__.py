"""CPU oracle for the detect+track hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy, the arithmetic of the reference's per-frame
detect-and-track path (YOLOv8-P2 forward, DFL decode, NMS, box rescale, the
constant-velocity Kalman multi-target tracker and the Ultralytics XYAH/XYWH
Kalman filters).  Every function cites the reference file:line it follows.

Rules (DESIGN.md "Oracle"):
  * Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
    ``--impl reference`` legs of ``bench.py`` may import this package.  The product
    package never imports it and has no CPU fallback.
  * Parity pinning: the reference ships no golden vectors for this path
    (SURVEY.md section 4), so the oracle is pinned against outputs of the reference
    itself, executed in the build container by ``tests/golden/make_golden.py``
    (``PYTHONPATH=/root/reference``) and committed as ``tests/golden/*.npz``.
    ``tests/test_oracle_vs_golden.py`` checks the oracle against those files.
"""
