"""CPU oracle for the TraceFly GA hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product package never does (tests/test_no_oracle_in_product.py
enforces it).  See oracle/README.md.
"""
