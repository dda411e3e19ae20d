import sys, ctypes as C
import numpy as np
import torch
sys.path.insert(0, ".")
from cadm_b200 import _lib
lib = _lib.load()
torch.zeros(1).cuda()
for N in (208, 128, 64, 48, 16):
    for n in (13, 130, 1300):
        out = np.zeros(2, dtype=np.int64)
        rc = lib.cadm_selftest_tc_rate(N, n, 2048, out.ctypes.data_as(C.c_void_p))
        print(f"N={N:3d} n_mma={n:5d} rc={rc} issue={out[0]:8d} done={out[1]:8d} cycles/mma={out[1]/n:7.1f}  floor={128*N/256:.0f}")
