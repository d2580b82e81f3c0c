"""Test infrastructure only: CPU restatements (oracles) of the two DUSt3R hot paths.
Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under dust3r_b200/ imports this package."""
