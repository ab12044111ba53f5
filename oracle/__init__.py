"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of specimux's read-matching path, used as the *checker* by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under specimux_b200/ may import this package.
"""
