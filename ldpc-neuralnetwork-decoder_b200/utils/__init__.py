"""Utility functions of the engine (mirror of the reference's ldpc_neural_decoder.utils)."""
from .ldpc_utils import (get_LLR_indexes, create_LLR_mapping, expand_base_matrix, load_base_matrix,
                         QCCode, as_code)
from .channel import AWGNChannel, compute_ber_fer, count_errors

__all__ = ["get_LLR_indexes", "create_LLR_mapping", "expand_base_matrix", "load_base_matrix",
           "QCCode", "as_code", "AWGNChannel", "compute_ber_fer", "count_errors"]
