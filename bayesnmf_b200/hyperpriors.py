"""Host-side defaults of the hyperprior parameters (R/setup.R:123-181) and their
expansion to the names the C ABI takes (fill_hyperprior_params_, R/setup.R:15-88).
Scalars stay scalars on the device; a matrix is only uploaded when the user gave one
(fill_matrix_ only fills a missing matrix, R/setup.R:102-113)."""
import math

PRIOR_LETTERS = {"truncnormal": "msab", "exponential": "ab", "gamma": "abcd"}


def default_hyperprior_params(prior, mean_data, N):
    """get_default_{truncnorm,exp,gamma}_hyperprior_params_ (R/setup.R:123-181)."""
    if prior == "truncnormal":
        v = {"m": 0.0, "s": math.sqrt(mean_data / N), "a": N + 1.0, "b": math.sqrt(N)}
    elif prior == "exponential":
        v = {"a": 10.0 * math.sqrt(N), "b": 10.0 * math.sqrt(mean_data)}
    elif prior == "gamma":
        v = {"a": 10.0 * math.sqrt(N), "b": 10.0, "c": 10.0 * math.sqrt(mean_data), "d": 10.0}
    else:
        raise ValueError("prior must be one of truncnormal, exponential, gamma")
    return {f"{k}_{e}": x for e in ("p", "e") for k, x in v.items()}


def fill_hyperprior_params(user, prior, mean_data, N):
    """Defaults overridden by the user's list; returns {ABI name: scalar or matrix}.
    A matrix `A_p` wins over a scalar `a_p` (R/setup.R:102-113)."""
    hp = default_hyperprior_params(prior, mean_data, N)
    hp.update(user or {})
    out = {}
    for letter in PRIOR_LETTERS[prior]:
        for e in ("p", "e"):
            mat, sca = f"{letter.upper()}_{e}", f"{letter}_{e}"
            out[mat] = hp[mat] if mat in hp else hp[sca]
    return out
