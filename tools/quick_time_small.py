"""Small-N probe: staged vs global-walk force kernel at 10k-100k particles.  python tools/quick_time_small.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import quick_time
import plife
from plife import synth

synth.CONFIGS["S50k"] = dict(n=50_000, m=6, rmax=0.0179, wrap=True, seed=0x5EED0011)    # nx=55, 16.5 per cell
synth.CONFIGS["S200k"] = dict(n=200_000, m=6, rmax=0.00895, wrap=True, seed=0x5EED0012)  # nx=111
for name in ("C1", "S50k", "S200k"):
    quick_time.run(name, steps=200)
    quick_time.run(name, steps=200, flags=plife.FLAG_FORCE_V1)
