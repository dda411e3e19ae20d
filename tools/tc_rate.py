"""tcgen05.mma issue/execute rate on one SM for the shapes the rollout kernels use (cadm_selftest_tc_rate)."""
import sys, ctypes as C
import numpy as np
import torch
sys.path.insert(0, ".")
from cadm_b200 import _lib
lib = _lib.load()
torch.zeros(1).cuda()
for bg in (0, 1, 2, 3):
    for swapped in (0, 1):
        for N in (208, 64, 32):
            n = 1170
            out = np.zeros(3, dtype=np.int64)
            rc = lib.cadm_selftest_tc_rate(N, n, 2048 | (1 << 16), 1, swapped, bg, out.ctypes.data_as(C.c_void_p))
            print(f"bg={bg} swapped={swapped} N={N:3d} n_mma={n:5d} rc={rc} issue={out[0]:8d} done={out[1]:8d} "
                  f"cycles/mma={out[1]/n:7.1f}  floor={128*N/256:.0f}  copies={out[2]} ({out[2]*8192/max(out[1],1):.1f} B/clk)", flush=True)
