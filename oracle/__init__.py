"""oracle/ -- TEST INFRASTRUCTURE: CPU truth for the FIR / CIC hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (ac_dsp_b200) never does.
"""
