"""Write a synthetic network's structure to a flat binary file for csrc/tools/analysis_bench.cpp."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmc_rates_b200 import synth

side = int(sys.argv[1]) if len(sys.argv) > 1 else 316
out = sys.argv[2] if len(sys.argv) > 2 else "/tmp/graph_%d.bin" % side
n, src, dst, k, A, B = synth.landscape_2d(side, T=0.3, seed=0)
with open(out, "wb") as f:
    np.array([n, src.size, len(A), len(B)], np.int64).tofile(f)
    np.asarray(src, np.int64).tofile(f); np.asarray(dst, np.int64).tofile(f)
    np.asarray(A, np.int64).tofile(f); np.asarray(B, np.int64).tofile(f)
print(out, n, src.size)
