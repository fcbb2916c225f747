"""b200-ldpc: B200-native LDPC message-passing engine (import as `ldpc_b200`).

Drop-in for the hot path of BananaFalls/LDPC-NeuralNetwork-Decoder: the reference's decoder
classes and utilities keep their names and signatures (`models`, `utils`) and call the CUDA
engine in csrc/ through the C ABI of include/ldpc_b200.h.  No CPU fallback.
"""
from . import _native  # noqa: F401
from . import utils, models  # noqa: F401

__version__ = "0.1.0"
