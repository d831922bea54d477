"""Build the CUDA library in-tree:  python -m bayesnmf_b200.build

One nvcc invocation, sm_100a only.  -fmad=false keeps a*b+c un-contracted so that the
arithmetic contract of the latent-count kernel (and the 1e-6 parity of every
conditional parameter) holds against the fp64 oracle.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "bnmf_api.cu")
OUT = os.path.join(HERE, "libbnmf_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC,-pthread", "-shared",
]


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(HERE, "..", "include", "bnmf.h")]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False, defines=(), out=None):
    """`defines` / `out`: experiment builds (tools/): the same sources with -D switches into another file,
    selected at run time with BNMF_LIB=<path>; the product is always OUT without defines."""
    if out is None and not force and not needs_build():
        return OUT
    out = out or OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
