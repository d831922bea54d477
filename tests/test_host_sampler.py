"""Host logic of bayesnmf_b200.sampler that needs no GPU: convergence control, the
temperature schedule and the convergence state machine (R/convergence.R, R/utils.R:307-332)."""
import warnings

import numpy as np
import pytest

from bayesnmf_b200 import sampler as S
from oracle import gibbs as og


def test_new_convergence_control_defaults_and_guard():
    cc = S.new_convergence_control()
    assert cc == dict(MAP_over=1000, MAP_every=100, tol=0.001, Ninarow_nochange=5, Ninarow_nobest=10,
                      miniters=1000, maxiters=5000, minA=0, metric="logposterior")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert S.new_convergence_control(miniters=10, maxiters=10)["miniters"] == 0      # R/convergence.R:28-31
        assert len(w) == 1


def test_temperature_schedule_matches_oracle_restatement():
    for length, n_temp in ((6000, 1000), (700, 100), (50, 400)):
        a = S.get_temp_sched(length, n_temp, np.random.default_rng(0))
        b = og.get_temp_sched(length, n_temp, np.random.default_rng(0))
        np.testing.assert_array_equal(a, b)
        assert a[0] == 0.0 and np.all(np.diff(a) >= 0) and a[-1] == 1.0 or length < n_temp


class _Fake(S.bayesNMF_sampler):
    """The convergence state machine without a device: MAP metrics are injected."""

    def __init__(self, cc, temps):
        self.specs = dict(convergence_control=cc, MH=False)
        self.state = dict(iter=0, converged=False, MAP_metrics=[])
        self.temperature_schedule = temps
        self.feed = None

    def _update_MAP_metrics(self, final=False):
        self.state["MAP_metrics"].append(dict(iter=self.state["iter"], logposterior=self.feed))


def _run(values, cc, temps, every=100, start=1000):
    f = _Fake(cc, temps)
    for i, v in enumerate(values):
        f.state["iter"] = start + every * i
        f.feed = v
        f._check_convergence()
        if f.state["converged"]:
            break
    return f.state


def test_converges_on_no_change():
    cc = S.new_convergence_control()
    st = _run([-1000.0 - 1e-4 * i for i in range(20)], cc, np.ones(10000))
    # the first check compares with prev := metric + 1, i.e. a relative change of 1/1001 < tol for
    # a metric this large: it already counts as "no change" (R/convergence.R:82-99), so 5 checks do
    assert st["converged"] and st["why"] == "no change" and st["iter"] == 1000 + 100 * 4


def test_converges_on_no_best_and_waits_for_temperature():
    cc = S.new_convergence_control()
    vals = [-1000.0] + [-1100.0 - 30.0 * i for i in range(40)]     # keeps changing (> tol), never better again
    st = _run(vals, cc, np.ones(10000))
    assert st["converged"] and st["why"] == "no best" and st["iter"] == 1000 + 100 * 10
    temps = np.concatenate([np.linspace(0, 1, 2500), np.ones(7500)])
    st2 = _run([-1000.0] * 40, cc, temps)
    # not eligible before every temperature of the window [iter - MAP_over, iter] is 1
    assert st2["converged"] and st2["iter"] >= 2500 + 1000


def test_na_metric_resets_counters():
    cc = S.new_convergence_control()
    st = _run([-1000.0, -1000.0, np.nan, -1000.0, -1000.0], cc, np.ones(10000))
    assert not st["converged"] and st["inarow_na"] == 0 and st["inarow_no_change"] == 1


def test_max_iters():
    cc = S.new_convergence_control(maxiters=1300)
    st = _run([-1000.0 * (1 + 0.1 * i) for i in range(4)], cc, np.ones(5000))
    assert st["converged"] and st["why"] == "max iters" and st["iter"] == 1300
