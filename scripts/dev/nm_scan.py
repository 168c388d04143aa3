import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pmc_b200
from quick_time import run
for nm in (1, 2, 3, 4, 6, 8):
    run(2**24, 0.70, 20, n_M=nm)
