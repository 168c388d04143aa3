"""Development aid: time the fused sweep at the BASELINE metric config (N=2^24, phi=0.70)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_time import run
run(2 ** 24, 0.70, int(os.environ.get("PMC_SWEEPS", 40)))
