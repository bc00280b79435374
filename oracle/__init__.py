"""Test infrastructure: CPU restatement of the reference's OTF degradation path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  See oracle/otf_oracle.py.
"""
