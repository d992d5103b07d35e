"""TEST INFRASTRUCTURE ONLY.  CPU oracles for the hot path: ``lw_oracle.c`` (a
plain-C restatement of the reference's algorithm) and ``_ref/`` (the reference's
own C++ compiled from /root/reference by oracle/Makefile).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; product code under lightweaver_b200/ never does."""
