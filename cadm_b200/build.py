"""Build the C-ABI shared library in-tree with nvcc for sm_100a (B200).

    python -m cadm_b200.build [--force]

Output: cadm_b200/libcadm_b200.so (git-ignored; travels to the GPU box with the gpurun snapshot).
The library has no torch / python dependency; the python package loads it with ctypes.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcadm_b200.so")
STAMP = os.path.join(HERE, ".libcadm_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr",
    "-shared", "-cudart", "shared",
]


# Experiments only (e.g. CADM_EXTRA_NVCC_FLAGS="-DCADM_SPLIT_FHADD=1"): extra flags enter the digest, so switching them rebuilds.
EXTRA_FLAGS = os.environ.get("CADM_EXTRA_NVCC_FLAGS", "").split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "cadm_b200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS + EXTRA_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force=False, verbose=False):
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + EXTRA_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libcadm_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
