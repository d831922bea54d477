"""TEST INFRASTRUCTURE ONLY: g++ build of tests/hostcheck/hostcheck.cpp (the kernels'
__host__ __device__ draw code, executed on the CPU so it can be compared with oracle/
without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck.cpp")
OUT = os.path.join(HERE, "libhostcheck.so")
DEPS = [SRC, os.path.join(HERE, "..", "..", "bayesnmf_b200", "csrc", "bnmf_rng.cuh")]


def build():
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    # -ffp-contract=off: same "no FMA contraction" contract as nvcc -fmad=false
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", SRC, "-o", OUT, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build())
