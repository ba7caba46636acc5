"""CPU oracle for the MolNexTR hot path -- test infrastructure, never shipped.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package (oracle/README.md)."""
