"""oracle/ — CPU restatement of the reference (gaetanserre/LiSA) render path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; lisa_b200 never does.
"""
