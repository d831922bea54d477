"""TEST / BASELINE INFRASTRUCTURE ONLY (never imported by bayesnmf_b200/).

ctypes wrapper and g++ build of oracle/cpu_port.cpp: the C++ / OpenMP port of the Poisson
latent-count Gibbs iteration that bench.py times as `cpu_baseline` and `--impl reference`
("kind": "port-c++": R is not installed in this image) and that
tests/test_cpu_port.py compares with the numpy oracle."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_port.cpp")
OUT = os.path.join(HERE, "libcpu_port.so")
DEPS = [SRC, os.path.join(HERE, "..", "bayesnmf_b200", "csrc", "bnmf_rng.cuh")]
PRIORS = {"exponential": 1, "gamma": 2}
NAMES = {"P": 0, "E": 1, "SP": 2, "SE": 3, "Alpha_p": 4, "Beta_p": 5, "Alpha_e": 6, "Beta_e": 7, "Lambda_p": 8, "Lambda_e": 9}
ROW = ["iter", "RMSE", "KL", "loglikelihood", "logposterior", "n_params", "BIC", "rank", "temp"]


def build():
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    # -ffp-contract=off: no FMA contraction, the contract nvcc -fmad=false gives the kernels
    cmd = ["g++", "-O3", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", SRC, "-o", OUT, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        L.cp_create.restype = ctypes.c_void_p
        L.cp_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                                ctypes.c_uint64, dp, dp]
        L.cp_destroy.argtypes = [ctypes.c_void_p]
        L.cp_init.argtypes = [ctypes.c_void_p, dp]
        L.cp_step.argtypes = [ctypes.c_void_p, ctypes.c_int, dp]
        L.cp_get.argtypes = [ctypes.c_void_p, ctypes.c_int, dp]
        L.cp_get.restype = ctypes.c_longlong
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class CpuPort:
    """Poisson likelihood, gamma or exponential prior, fixed rank, fp64 state; scalar hyperprior
    parameters (the defaults of R/setup.R:123-181 unless `hyper` = {"A_p": ..} is given)."""

    def __init__(self, M, N, prior="gamma", seed=0, hyper=None, g0=0, G_total=None, mean_data=None):
        from oracle.gibbs import default_hyperprior_params
        M = np.asarray(M, dtype=np.float64)
        self.K, self.G = M.shape
        self.N = int(N)
        hp = default_hyperprior_params(prior, float(M.mean()) if mean_data is None else float(mean_data), self.N)
        hp.update({k.lower(): float(np.asarray(v).reshape(-1)[0]) for k, v in (hyper or {}).items()})
        hv = np.array([hp.get(f"{c}_{e}", 0.0) for e in "pe" for c in "abcd"], dtype=np.float64)
        flat = np.ascontiguousarray(M.T).reshape(-1)
        self._h = lib().cp_create(self.K, self.N, self.G, self.G if G_total is None else int(G_total), int(g0), PRIORS[prior],
                                  int(seed) & 0xFFFFFFFFFFFFFFFF, _dp(flat), _dp(hv))
        if not self._h:
            raise ValueError("cpu_port: gamma or exponential prior only")

    def init_from_prior(self):
        row = np.empty(9)
        lib().cp_init(self._h, _dp(row))
        return dict(zip(ROW, row))

    def step(self, n=1):
        rows = np.empty((int(n), 9))
        lib().cp_step(self._h, int(n), _dp(rows))
        return rows

    def get(self, name):
        shp = {"P": (self.K, self.N), "SP": (self.K, self.N)}.get(name, (self.K, self.N) if name.endswith("_p") else (self.N, self.G))
        out = np.empty(int(np.prod(shp)))
        assert lib().cp_get(self._h, NAMES[name], _dp(out)) == out.size
        return out.reshape(shp, order="F")

    def close(self):
        if self._h:
            lib().cp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


if __name__ == "__main__":
    print(build())
