"""MMA stream of the swapped kernel in isolation (cadm_selftest_tcs_rate): cycles per (K16 block, M tile) pair of MMAs.
mode bits: 1 bulk copies stream into the ring, 2 alternate accumulators, 4 epilogue-like traffic, 8 on all 148 SMs, 16 on 74 CTAs."""
import sys, ctypes as C
import numpy as np
import torch
sys.path.insert(0, ".")
from cadm_b200 import _lib
lib = _lib.load()
torch.zeros(1).cuda()
for mode in (0, 8, 16, 9, 17, 15):
    for rows in (32, 48):
        for kps in (4,):
            iters = 200
            out = np.zeros(3, dtype=np.int64)
            rc = lib.cadm_selftest_tcs_rate(rows, iters, 128, kps, mode, out.ctypes.data_as(C.c_void_p))
            pairs = iters * 4 * kps
            if rc:
                print("error:", lib.cadm_last_error(None).decode())
            print(f"mode={mode:2d} rows={rows} kps={kps} rc={rc} issue={out[0]:8d} done={out[1]:8d} cycles/pair={out[1]/pairs:7.1f} "
                  f"stream={out[2]/max(out[1],1):6.1f} B/clk", flush=True)
