"""CPU oracle for the statdepth hot path -- TEST INFRASTRUCTURE, never imported by the product.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""
