"""Utility functions of the engine (mirror of the reference's ldpc_neural_decoder.utils)."""
from .ldpc_utils import (get_LLR_indexes, create_LLR_mapping, expand_base_matrix, load_base_matrix,
                         QCCode, as_code)
from .encoder import SystematicEncoder
from .rate_match import RateMatcher, rate_match_tables
from .channel import (AWGNChannel, QPSKChannel, compute_ber_fer, count_errors, qpsk_modulate, awgn_channel,
                      qpsk_demodulate)

__all__ = ["get_LLR_indexes", "create_LLR_mapping", "expand_base_matrix", "load_base_matrix",
           "QCCode", "as_code", "AWGNChannel", "QPSKChannel", "compute_ber_fer", "count_errors",
           "qpsk_modulate", "awgn_channel", "qpsk_demodulate", "SystematicEncoder", "RateMatcher", "rate_match_tables"]
