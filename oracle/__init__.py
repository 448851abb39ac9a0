"""CPU oracle for the SMPL + reprojection-loss hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package; the product
(human-pose-estimation_b200/) never does and has no CPU fallback.
"""
