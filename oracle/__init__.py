"""CPU oracle for the ERL-Fill hot path — TEST INFRASTRUCTURE ONLY (see erlfill_oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs.
"""
