"""TEST INFRASTRUCTURE — the checkers for reak_b200, never part of the product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  It holds
  * `_ref/libreak_ref.so`: the unmodified ReaK sources for the hot path (compiled from
    /root/reference by oracle/Makefile, wrapped by ref_lib.cpp), and
  * `libkte_oracle.so`: kte_oracle.c, a plain-C restatement of the same algorithm that cites
    the reference file:line it follows (pinned against `_ref` and tests/golden/).
"""
