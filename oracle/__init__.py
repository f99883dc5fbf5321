"""TEST INFRASTRUCTURE ONLY — CPU oracle for i2v_b200 (see oracle/i2v_oracle.c, oracle/oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product path (i2v_b200, image_attacks, TPAMI_attack, base_attacks) never does.
"""
