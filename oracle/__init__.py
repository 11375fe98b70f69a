"""TEST INFRASTRUCTURE — CPU oracle for the InfoNCE / IIC loss hot path (see the module docstrings).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
