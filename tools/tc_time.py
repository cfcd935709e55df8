"""timing of the tensor-core kernel at one shape (development): python tools/tc_time.py B T [precisions...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tc_dev
B, T = int(sys.argv[1]), int(sys.argv[2])
tc_dev.timing(B, T, tuple(sys.argv[3:]) or ("tc",))
