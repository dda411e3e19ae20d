"""Tensor-core path (tcgen05 / TMEM): operand layouts, descriptors and split arithmetic, in isolation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(16, 16), (32, 208), (208, 208), (200, 48), (24, 208)])
def test_tc_gemm_selftest_matches_fp64(K, N):
    """One 128 x N x K product through the rollout kernel's UMMA layouts vs an fp64 matmul.
    terms=3 (fp16 hi + bf16 lo, three MMAs) must be fp32-class; terms=1 is fp16-class."""
    from cadm_b200.engine import selftest_tc_gemm
    rng = np.random.default_rng(K * 1000 + N)
    X = rng.standard_normal((128, K)).astype(np.float32)
    W = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    ref = X.astype(np.float64) @ W.astype(np.float64)
    scale = np.sqrt(np.mean(ref ** 2))
    out3 = selftest_tc_gemm(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), terms=3).cpu().numpy()
    err3 = np.max(np.abs(out3 - ref)) / scale
    assert err3 < 5e-6, err3
    out1 = selftest_tc_gemm(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), terms=1).cpu().numpy()
    err1 = np.max(np.abs(out1 - ref)) / scale
    assert err1 < 5e-3, err1
    assert err3 < err1


def test_tc_gemm_selftest_identity_layout():
    """X = one-hot rows, W = distinct integers: any row / column / k permutation error shows up exactly."""
    from cadm_b200.engine import selftest_tc_gemm
    K, N = 208, 208
    X = np.zeros((128, K), np.float32)
    for r in range(128):
        X[r, (r * 7) % K] = 1.0
    W = (np.arange(K)[:, None] * 0.5 + np.arange(N)[None, :] * 0.001953125).astype(np.float32)
    out = selftest_tc_gemm(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), terms=3).cpu().numpy()
    want = W[[(r * 7) % K for r in range(128)]]
    np.testing.assert_allclose(out, want, rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("rows,K,N,kps", [(16, 16, 16, 1), (32, 32, 208, 2), (48, 208, 208, 2), (64, 208, 208, 4),
                                          (32, 200, 48, 3), (64, 24, 200, 2), (48, 208, 128, 2)])
def test_tcs_gemm_selftest_matches_fp64(rows, K, N, kps):
    """The swapped-operand product (weights on the MMA's M axis incl. the second, partial M tile; rows on its N axis as an
    MN-major B operand; kps K16 blocks per weight stage) vs an fp64 matmul."""
    from cadm_b200.engine import selftest_tcs_gemm
    rng = np.random.default_rng(rows * 100000 + K * 1000 + N)
    X = rng.standard_normal((rows, K)).astype(np.float32)
    W = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    ref = X.astype(np.float64) @ W.astype(np.float64)
    scale = np.sqrt(np.mean(ref ** 2))
    out3 = selftest_tcs_gemm(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), kps=kps, terms=3).cpu().numpy()
    err3 = np.max(np.abs(out3 - ref)) / scale
    assert err3 < 5e-6, err3
    out1 = selftest_tcs_gemm(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), kps=kps, terms=1).cpu().numpy()
    err1 = np.max(np.abs(out1 - ref)) / scale
    assert err1 < 5e-3, err1


def test_tcs_gemm_selftest_identity_layout():
    """One-hot rows against distinct integers: any row / hidden-unit / k permutation error of the swapped layouts shows."""
    from cadm_b200.engine import selftest_tcs_gemm
    rows, K, N = 64, 208, 200
    X = np.zeros((rows, K), np.float32)
    for r in range(rows):
        X[r, (r * 7 + 3) % K] = 1.0
    W = (np.arange(K)[:, None] * 0.5 + np.arange(N)[None, :] * 0.001953125).astype(np.float32)
    out = selftest_tcs_gemm(torch.from_numpy(X).cuda(), torch.from_numpy(W).cuda(), kps=2, terms=3).cpu().numpy()
    want = W[[(r * 7 + 3) % K for r in range(rows)]]
    np.testing.assert_allclose(out, want, rtol=2e-6, atol=1e-6)
