"""The C-ABI library: loads, exports every symbol include/bnmf.h declares, validates
models like check_model (R/bayesNMF_sampler.R:623-645) and refuses to run without a
GPU (there is no CPU path).  No compute calls -- CPU only."""
import ctypes
import itertools
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bnmf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bnmf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    L = ctypes.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/bnmf.h but not exported"
    from bayesnmf_b200._lib import EXPORTS
    assert sorted(EXPORTS) == names


def test_check_model_matrix(built_lib):
    from bayesnmf_b200 import _lib
    from oracle.gibbs import check_model
    L = _lib.lib()
    for lik, pri, mh in itertools.product(_lib.LIKELIHOODS, _lib.PRIORS, (False, True)):
        buf = ctypes.create_string_buffer(256)
        rc = L.bnmf_check_model(_lib.LIKELIHOODS[lik], _lib.PRIORS[pri], int(mh), buf, 256)
        try:
            check_model(lik, pri, mh)
            ok, msg = True, ""
        except ValueError as e:
            ok, msg = False, str(e)
        assert (rc == 0) == ok, (lik, pri, mh)
        if not ok:
            assert buf.value.decode() == msg          # same wording as the reference's error
    buf = ctypes.create_string_buffer(256)
    assert L.bnmf_check_model(7, 0, 0, buf, 256) != 0


def test_config_struct_matches_header():
    from bayesnmf_b200._lib import Config
    # K,N:int32 x2 | G,G_total,g0:int64 x3 | 8 x int32 | seed:uint64
    assert ctypes.sizeof(Config) == 8 + 24 + 32 + 8
    assert Config.seed.offset == 64


def test_no_cpu_fallback(built_lib):
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from bayesnmf_b200 import BnmfError, Handle
    with pytest.raises(BnmfError, match="no CUDA device|CUDA"):
        Handle(np.ones((4, 3)), 2)


def test_null_arguments_fail_cleanly(built_lib):
    from bayesnmf_b200 import _lib
    L = _lib.lib()
    assert L.bnmf_create(None, None, None) != 0
    assert b"null" in L.bnmf_last_error()
    assert L.bnmf_step(None, 1, 0, None, None, None) != 0


def test_r_veneer_compiles_against_stub_and_uses_exported_entries():
    """r/rcall.c (the .Call veneer a maintainer adds to the reference) parses against a stub of R's C
    API and calls only entry points that include/bnmf.h declares."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "r", "rcall.c")
    r = subprocess.run(["gcc", "-fsyntax-only", "-I" + os.path.join(root, "tests", "stubs"), "-I" + os.path.join(root, "include"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    used = set(re.findall(r"\b(bnmf_[a-z_0-9]+)\s*\(", open(src).read()))
    declared = set(re.findall(r"\b(bnmf_[a-z_0-9]+)\s*\(", open(os.path.join(root, "include", "bnmf.h")).read()))
    assert used and used <= declared, used - declared


def test_have_prior_mask_of_the_r_patch_matches_the_python_binding():
    """r/bayesNMF_sampler_b200.R builds the same bit mask as bayesnmf_b200/_lib.py (CPU-only check of
    the file that cannot be run here: no R in the image)."""
    from bayesnmf_b200._lib import HAVE_PRIOR
    src = open(os.path.join(ROOT, "r", "bayesNMF_sampler_b200.R")).read()
    m = re.search(r"\.have_prior_bits <- c\((.*?)\)", src, flags=re.S)
    bits = {k: int(v) for k, v in re.findall(r"(\w+) = (\d+)L", m.group(1))}
    assert bits == HAVE_PRIOR
    assert "have_prior <- sum(.have_prior_bits[prior_mats])" in src
    assert "as.integer(length(init_prior_params) > 0)" not in src
