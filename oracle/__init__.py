"""oracle/ -- CPU restatement of the reference's algorithm for the hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package, and only as the checker / the CPU baseline --
never as part of the product path (nglod_b200/ does not and must not import it).
"""
