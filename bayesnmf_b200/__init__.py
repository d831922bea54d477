"""bayesnmf_b200 -- B200-native Gibbs sampler behind the bayesNMF sampler API."""
from ._lib import BnmfError, Handle, comm_unique_id, release_cached_memory  # noqa: F401
from .sampler import bayesNMF, bayesNMF_sampler, new_convergence_control  # noqa: F401
