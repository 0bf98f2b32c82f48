"""CPU checkers for the sparse-times-dense path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (spblas_reference_b200) never does.
"""
