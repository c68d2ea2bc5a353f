"""CPU oracle of the Back2Future hot path -- TEST INFRASTRUCTURE, NOT PRODUCT (parity unpinned,
see b2f_oracle.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package."""
