"""oracle/: CPU restatements of the reference algorithms and the recipe that compiles the
unmodified reference CPU path (oracle/_ref).  TEST INFRASTRUCTURE ONLY -- importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from the product."""
