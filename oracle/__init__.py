"""Oracle: CPU restatements of the reference's hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package; the
product package (dinov2_distillation_b200) never does.
"""
