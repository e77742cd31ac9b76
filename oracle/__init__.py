"""CPU oracle of the hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this package.  Nothing under numbotics_b200/ does (tests/test_layout.py
enforces it).  See DESIGN.md section 6 for what is pinned against the reference and what
is "parity unpinned".
"""
