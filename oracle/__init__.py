"""Test infrastructure only: CPU oracle (C restatement) and, when built, the compiled reference (oracle/_ref).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package."""
