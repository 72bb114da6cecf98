"""CPU oracle of the PANTERA hot path: TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), the cpu_baseline and
--impl reference legs of bench.py).  The product package never imports this."""
