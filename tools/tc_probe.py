"""Run one tensor-core GEMM self-test in its own process (an illegal instruction poisons the CUDA context)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cadm_b200.engine import selftest_tc_gemm

K, N, terms = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(0)
X = rng.standard_normal((128, K)).astype(np.float32)
W = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
ref = X.astype(np.float64) @ W.astype(np.float64)
try:
    out = selftest_tc_gemm(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), terms=terms).cpu().numpy()
    err = np.max(np.abs(out - ref)) / np.sqrt(np.mean(ref ** 2))
    print(f"K={K} N={N} terms={terms}: max err / rms = {err:.3e}; out[0,:4]={out[0,:4]} ref[0,:4]={ref[0,:4]}")
except Exception as e:
    print(f"K={K} N={N} terms={terms}: FAILED {e}")
