"""oracle/ — CPU restatements of the reference algorithm.  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by the product package xivo_b200/."""
