"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/dgr_oracle.c).  Nothing under dogs_b200/ may
import this package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs do."""
