"""fp64-mode timing probe: python tools/quick_time64.py [config] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import quick_time
import plife
quick_time.run(sys.argv[1] if len(sys.argv) > 1 else "C3", steps=int(sys.argv[2]) if len(sys.argv) > 2 else 5, precision=plife.F64)
