"""ctypes binding of include/bnmf.h.  There is no CPU path: if the CUDA library is
missing, or no GPU is visible when a sampler is created, this fails loudly."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# BNMF_LIB: an experiment build of the same sources (bayesnmf_b200/build.py, `defines`); default: the product library
LIB_PATH = os.environ.get("BNMF_LIB") or os.path.join(HERE, "libbnmf_b200.so")

POISSON, NORMAL = 0, 1
TRUNCNORMAL, EXPONENTIAL, GAMMA = 0, 1, 2
SBFI, BFI, BIC = 0, 1, 2
F64, F32 = 0, 1
LIKELIHOODS = {"poisson": POISSON, "normal": NORMAL}
PRIORS = {"truncnormal": TRUNCNORMAL, "exponential": EXPONENTIAL, "gamma": GAMMA}
RANK_METHODS = {"SBFI": SBFI, "BFI": BFI, "BIC": BIC}
MC_COLS = 11
METRIC_NAMES = ["iter", "RMSE", "KL", "loglikelihood", "logposterior", "n_params", "BIC", "rank", "temp",
                "P_mean_acceptance_rate", "E_mean_acceptance_rate"]
HAVE = {"P": 1, "E": 2, "A": 4, "Z": 8, "sigmasq": 16}
# have_prior bits: 0..4 = Mu,Sigmasq,Lambda,Alpha,Beta on the P side, 8..12 on the E side
HAVE_PRIOR = {"Mu_p": 1 << 0, "Sigmasq_p": 1 << 1, "Lambda_p": 1 << 2, "Alpha_p": 1 << 3, "Beta_p": 1 << 4,
              "Mu_e": 1 << 8, "Sigmasq_e": 1 << 9, "Lambda_e": 1 << 10, "Alpha_e": 1 << 11, "Beta_e": 1 << 12}


class Config(ctypes.Structure):
    _fields_ = [("K", ctypes.c_int32), ("N", ctypes.c_int32), ("G", ctypes.c_int64),
                ("G_total", ctypes.c_int64), ("g0", ctypes.c_int64),
                ("likelihood", ctypes.c_int32), ("prior", ctypes.c_int32), ("MH", ctypes.c_int32),
                ("learning_rank", ctypes.c_int32), ("rank_method", ctypes.c_int32),
                ("precision", ctypes.c_int32), ("device", ctypes.c_int32), ("ring_cap", ctypes.c_int32),
                ("seed", ctypes.c_uint64)]


_lib = None


class BnmfError(RuntimeError):
    pass


def lib():
    """Load libbnmf_b200.so (built in-tree by bayesnmf_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BnmfError(
            f"{LIB_PATH} not found: build it with `python -m bayesnmf_b200.build` "
            "(nvcc, sm_100a).  bayesnmf_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, cp, i32, i64, u32 = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32
    dp = ctypes.POINTER(ctypes.c_double)
    L.bnmf_last_error.restype = cp
    L.bnmf_check_model.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, cp, ctypes.c_size_t]
    L.bnmf_create.argtypes = [ctypes.POINTER(Config), dp, ctypes.POINTER(vp)]
    L.bnmf_destroy.argtypes = [vp]
    L.bnmf_destroy.restype = None
    L.bnmf_set_hyper.argtypes = [vp, cp, dp, i64, i64]
    L.bnmf_set_state.argtypes = [vp, cp, dp, i64]
    L.bnmf_get_state.argtypes = [vp, cp, dp, i64]
    L.bnmf_set_temperature_schedule.argtypes = [vp, dp, i64]
    L.bnmf_init_from_prior.argtypes = [vp, u32, u32, dp]
    L.bnmf_step.argtypes = [vp, i32, i32, dp, dp, dp]
    L.bnmf_run.argtypes = [vp, ctypes.c_void_p, i32, dp, i64, dp, i64, ctypes.c_void_p]
    ip = ctypes.POINTER(i32)
    L.bnmf_assign_signatures.argtypes = [vp, i32, dp, i32, ctypes.c_double, ip, ip, dp, ip, dp, dp, dp, ip]
    L.bnmf_ring_count.argtypes = [vp, ctypes.POINTER(i32)]
    L.bnmf_get_sample.argtypes = [vp, cp, i32, dp, i64]
    L.bnmf_get_map.argtypes = [vp, i32, dp, dp, dp, ctypes.POINTER(i32)]
    L.bnmf_get_credible_intervals.argtypes = [vp, i32, ctypes.c_double, ctypes.c_double, dp, dp, dp, dp, ctypes.POINTER(i32)]
    L.bnmf_comm_unique_id.argtypes = [ctypes.c_char_p]
    L.bnmf_comm_init.argtypes = [vp, ctypes.c_char_p, i32, i32]
    L.bnmf_comm_share.argtypes = [vp, vp]
    L.bnmf_timing.argtypes = [vp, dp, dp, dp, ctypes.POINTER(i64)]
    L.bnmf_set_l2_flush.argtypes = [vp, ctypes.c_size_t]
    L.bnmf_sample_z.argtypes = [vp, i32, dp]
    L.bnmf_profile_iteration.argtypes = [vp, i32, ctypes.c_char_p, dp, ctypes.POINTER(i32), i32, ctypes.POINTER(i32)]
    L.bnmf_release_cached_memory.argtypes = []
    _lib = L
    return L


class ConvergenceControl(ctypes.Structure):
    """bnmf_convergence_control of include/bnmf.h."""
    _fields_ = [("MAP_over", ctypes.c_int32), ("MAP_every", ctypes.c_int32), ("tol", ctypes.c_double),
                ("Ninarow_nochange", ctypes.c_int32), ("Ninarow_nobest", ctypes.c_int32), ("miniters", ctypes.c_int32),
                ("maxiters", ctypes.c_int32), ("metric", ctypes.c_int32)]


class RunResult(ctypes.Structure):
    """bnmf_run_result of include/bnmf.h."""
    _fields_ = [("iter", ctypes.c_int32), ("converged", ctypes.c_int32), ("converged_iter", ctypes.c_int32), ("why", ctypes.c_int32),
                ("best_iter", ctypes.c_int32), ("n_checks", ctypes.c_int32), ("n_rows", ctypes.c_int32),
                ("inarow_no_change", ctypes.c_int32), ("inarow_no_best", ctypes.c_int32), ("inarow_na", ctypes.c_int32),
                ("best_MAP_metric", ctypes.c_double), ("prev_MAP_metric", ctypes.c_double)]


RUN_METRICS = {"logposterior": 0, "loglikelihood": 1, "BIC": 2}
RUN_WHY = {0: None, 1: "no change", 2: "no best", 3: "max iters"}
MAP_METRIC_NAMES = ["iter", "loglikelihood", "logposterior", "n_params", "BIC", "rank", "MAP_A_counts", "mean_temp"]

EXPORTS = ["bnmf_check_model", "bnmf_create", "bnmf_destroy", "bnmf_last_error", "bnmf_set_hyper",
           "bnmf_set_state", "bnmf_get_state", "bnmf_set_temperature_schedule", "bnmf_init_from_prior",
           "bnmf_step", "bnmf_run", "bnmf_ring_count", "bnmf_get_sample", "bnmf_get_map", "bnmf_get_credible_intervals", "bnmf_assign_signatures", "bnmf_comm_unique_id",
           "bnmf_comm_init", "bnmf_comm_share", "bnmf_timing", "bnmf_set_l2_flush", "bnmf_sample_z", "bnmf_profile_iteration",
           "bnmf_release_cached_memory"]


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _f64(a):
    """Column-major (R layout) contiguous float64 copy, flattened."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).reshape(-1) if np.ndim(a) == 2 \
        else np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(-1)


class Handle:
    """Thin owner of a bnmf_handle*; shapes follow the reference (K x N, N x G)."""

    def __init__(self, data, N, likelihood="poisson", prior="gamma", MH=False, learning_rank=False,
                 rank_method="SBFI", seed=0, precision="f64", device=0, ring_cap=0, g0=0, G_total=None):
        L = lib()
        data = np.asarray(data, dtype=np.float64)
        self.K, self.G = data.shape
        self.N = int(N)
        cfg = Config(K=self.K, N=self.N, G=self.G, G_total=self.G if G_total is None else int(G_total), g0=int(g0),
                     likelihood=LIKELIHOODS[likelihood], prior=PRIORS[prior], MH=int(bool(MH)),
                     learning_rank=int(bool(learning_rank)), rank_method=RANK_METHODS[rank_method],
                     precision=F64 if precision == "f64" else F32, device=int(device), ring_cap=int(ring_cap),
                     seed=int(seed) & 0xFFFFFFFFFFFFFFFF)
        self.cfg = cfg
        self._h = ctypes.c_void_p()
        flat = _f64(data)
        self._ck(L.bnmf_create(ctypes.byref(cfg), _dp(flat), ctypes.byref(self._h)))
        self.shapes = {
            "P": (self.K, self.N), "E": (self.N, self.G), "A": (self.N,), "R": (1,), "sigmasq": (self.G,),
            "SP": (self.K, self.N), "SE": (self.N, self.G), "Alpha": (self.G,), "Beta": (self.G,),
            "P_acceptance_rate": (self.K, self.N), "E_acceptance_rate": (self.N, self.G),
            "Mhat": (self.K, self.G), "rowsumE": (self.N,), "data_sum": (1,),
        }
        for nm in ("Mu", "Sigmasq", "Lambda", "Alpha", "Beta"):
            self.shapes[nm + "_p"] = (self.K, self.N)
            self.shapes[nm + "_e"] = (self.N, self.G)

    def _ck(self, rc):
        if rc != 0:
            raise BnmfError(lib().bnmf_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().bnmf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_hyper(self, name, value):
        v = np.asarray(value, dtype=np.float64)
        if v.ndim == 0 or v.size == 1:
            flat = v.reshape(1).copy()
            self._ck(lib().bnmf_set_hyper(self._h, name.encode(), _dp(flat), 1, 1))
        else:
            flat = _f64(v)
            self._ck(lib().bnmf_set_hyper(self._h, name.encode(), _dp(flat), v.shape[0], v.shape[1]))

    def set_state(self, name, value):
        flat = _f64(np.asarray(value, dtype=np.float64).reshape(self.shapes[name]))
        self._ck(lib().bnmf_set_state(self._h, name.encode(), _dp(flat), flat.size))

    def get_state(self, name):
        shp = self.shapes[name]
        n = int(np.prod(shp))
        out = np.empty(n, dtype=np.float64)
        self._ck(lib().bnmf_get_state(self._h, name.encode(), _dp(out), n))
        # (the library hands back R's layout: a column-major view, no transposing copy)
        return out.reshape(shp, order="F") if len(shp) == 2 else out

    def set_temperature_schedule(self, temps):
        t = np.ascontiguousarray(np.asarray(temps, dtype=np.float64))
        self._ck(lib().bnmf_set_temperature_schedule(self._h, _dp(t), t.size))

    def init_from_prior(self, have=(), have_prior=()):
        hv = sum(HAVE[h] for h in have if h in HAVE)
        hp = sum(HAVE_PRIOR[h] for h in have_prior if h in HAVE_PRIOR)
        row = np.empty(MC_COLS)
        self._ck(lib().bnmf_init_from_prior(self._h, hv, hp, _dp(row)))
        return dict(zip(METRIC_NAMES, row))

    def step(self, n_iters, converged=False, want_P=False, want_A=False):
        n_iters = int(n_iters)
        met = np.empty((n_iters, MC_COLS))
        P = np.empty(n_iters * self.K * self.N) if want_P else None
        A = np.empty(n_iters * self.N) if want_A else None
        self._ck(lib().bnmf_step(self._h, n_iters, int(bool(converged)), _dp(met),
                                 _dp(P) if want_P else None, _dp(A) if want_A else None))
        out = {"metrics": met}
        if want_P:
            out["P"] = P.reshape(n_iters, self.N, self.K).transpose(0, 2, 1)
        if want_A:
            out["A"] = A.reshape(n_iters, self.N)
        return out

    def ring_count(self):
        c = ctypes.c_int32()
        self._ck(lib().bnmf_ring_count(self._h, ctypes.byref(c)))
        return c.value

    def get_sample(self, name, ago=0):
        shp = self.shapes[name]
        n = int(np.prod(shp))
        out = np.empty(n)
        self._ck(lib().bnmf_get_sample(self._h, name.encode(), int(ago), _dp(out), n))
        # (the library hands back R's layout: a column-major view, no transposing copy)
        return out.reshape(shp, order="F") if len(shp) == 2 else out

    def run(self, convergence_control, post_warmup=0):
        """run_gibbs_sampler behind the ABI (bnmf_run): advances until convergence (+ post_warmup
        iterations with the real MH accept step).  Returns dict(result fields, metrics, MAP_metrics)."""
        cc = convergence_control
        c = ConvergenceControl(MAP_over=cc["MAP_over"], MAP_every=cc["MAP_every"], tol=cc["tol"],
                               Ninarow_nochange=cc["Ninarow_nochange"], Ninarow_nobest=cc["Ninarow_nobest"],
                               miniters=cc["miniters"], maxiters=cc["maxiters"], metric=RUN_METRICS[cc.get("metric", "logposterior")])
        rows_cap = cc["maxiters"] + int(post_warmup) + 8
        checks_cap = rows_cap // max(cc["MAP_every"], 1) + 8
        met = np.empty((rows_cap, MC_COLS)); mm = np.empty((checks_cap, len(MAP_METRIC_NAMES)))
        res = RunResult()
        self._ck(lib().bnmf_run(self._h, ctypes.byref(c), int(post_warmup), _dp(met), rows_cap, _dp(mm), checks_cap, ctypes.byref(res)))
        out = {f: getattr(res, f) for f, _ in RunResult._fields_}
        out["why"] = RUN_WHY[res.why]
        out["metrics"] = met[:res.n_rows]
        out["MAP_metrics"] = [dict(zip(MAP_METRIC_NAMES, r)) for r in mm[:res.n_checks]]
        return out

    def assign_signatures(self, n_samples, reference_P, credible_interval=0.95):
        """assign_signatures_ensemble_ (R/postprocessing.R:175-341) over the retained samples: dict(keep_sigs,
        assignment (reference column per included signature), votes (n_keep x n_ref shares), MAP_cosine,
        lower_cosine, upper_cosine, n_match)."""
        ref = np.asarray(reference_P, dtype=np.float64)
        if ref.shape[0] != self.K:
            raise ValueError(f"Reference matrix has {ref.shape[0]} rows, but data has {self.K} rows.")
        R = ref.shape[1]
        i32a = lambda n: np.zeros(n, dtype=np.int32)
        nk, nm, keep, asg = ctypes.c_int32(), ctypes.c_int32(), i32a(self.N), i32a(self.N)
        votes = np.zeros(self.N * R); mc, lo, hi = np.zeros(self.N), np.zeros(self.N), np.zeros(self.N)
        ipt = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        self._ck(lib().bnmf_assign_signatures(self._h, int(n_samples), _dp(_f64(ref)), R, float(credible_interval), ctypes.byref(nk),
                                              ipt(keep), _dp(votes), ipt(asg), _dp(mc), _dp(lo), _dp(hi), ctypes.byref(nm)))
        k = nk.value
        return dict(keep_sigs=keep[:k].copy(), assignment=asg[:k].copy(), votes=votes.reshape((self.N, R), order="F")[:k].copy(),
                    MAP_cosine=mc[:k].copy(), lower_cosine=lo[:k].copy(), upper_cosine=hi[:k].copy(), n_match=nm.value)

    def get_credible_intervals(self, n_samples, lower_p=0.025, upper_p=0.975):
        """(P_lower, P_upper, E_lower, E_upper, n_match): element-wise quantiles over the samples that
        match the modal A, computed on the device ring (R/utils.R:264-287)."""
        KN, NG = self.K * self.N, self.N * self.G
        Pl, Ph, El, Eh = np.empty(KN), np.empty(KN), np.empty(NG), np.empty(NG)
        nm = ctypes.c_int32()
        self._ck(lib().bnmf_get_credible_intervals(self._h, int(n_samples), float(lower_p), float(upper_p),
                                                   _dp(Pl), _dp(Ph), _dp(El), _dp(Eh), ctypes.byref(nm)))
        f = lambda a, r, c: a.reshape((r, c), order="F")
        return f(Pl, self.K, self.N), f(Ph, self.K, self.N), f(El, self.N, self.G), f(Eh, self.N, self.G), nm.value

    def get_map(self, n_samples):
        P = np.empty(self.K * self.N); E = np.empty(self.N * self.G); A = np.empty(self.N)
        nm = ctypes.c_int32()
        self._ck(lib().bnmf_get_map(self._h, int(n_samples), _dp(P), _dp(E), _dp(A), ctypes.byref(nm)))
        return P.reshape(self.N, self.K).T.copy(), E.reshape(self.G, self.N).T.copy(), A, nm.value

    def comm_init(self, uid, rank, world):
        self._ck(lib().bnmf_comm_init(self._h, uid, int(rank), int(world)))

    def comm_share(self, other):
        self._ck(lib().bnmf_comm_share(self._h, other._h))

    def timing(self):
        t = ctypes.c_double(); i = ctypes.c_double(); z = ctypes.c_double(); l = ctypes.c_int64()
        self._ck(lib().bnmf_timing(self._h, ctypes.byref(t), ctypes.byref(i), ctypes.byref(z), ctypes.byref(l)))
        return {"total_ms": t.value, "iter_ms": i.value, "zstat_ms": z.value, "launches": l.value}

    def set_l2_flush(self, nbytes):
        self._ck(lib().bnmf_set_l2_flush(self._h, int(nbytes)))

    def sample_z(self, it):
        ms = ctypes.c_double()
        self._ck(lib().bnmf_sample_z(self._h, int(it), ctypes.byref(ms)))
        return ms.value

    def profile_iteration(self, converged=False, cap=64):
        """One iteration with a CUDA event after every kernel: {kernel name: (total ms, launches)}."""
        names = ctypes.create_string_buffer(32 * cap)
        ms = np.zeros(cap)
        cnt = np.zeros(cap, dtype=np.int32)
        n = ctypes.c_int32()
        self._ck(lib().bnmf_profile_iteration(self._h, int(bool(converged)), names, _dp(ms),
                                              cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), cap, ctypes.byref(n)))
        out = {}
        for j in range(min(n.value, cap)):
            out[names.raw[32 * j:32 * j + 32].split(b"\0")[0].decode()] = (float(ms[j]), int(cnt[j]))
        return out


def release_cached_memory():
    """Hand the device blocks kept from closed handles back to the driver (bnmf_release_cached_memory)."""
    if lib().bnmf_release_cached_memory() != 0:
        raise BnmfError(lib().bnmf_last_error().decode())


def comm_unique_id():
    buf = ctypes.create_string_buffer(128)
    rc = lib().bnmf_comm_unique_id(buf)
    if rc != 0:
        raise BnmfError(lib().bnmf_last_error().decode())
    return buf.raw
