"""TEST INFRASTRUCTURE: CPU oracle of the CaDM CEM/MPC planner hot path (see cadm_oracle.py header).

Nothing under cadm_b200/ imports this package.  Allowed importers: tests/, __graft_entry__.smoke(),
and the cpu_baseline / --impl reference legs of bench.py.
"""
