"""CPU oracle of the reference's LDDMM hot path. TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product (lagomorph_b200) never imports it.
"""
