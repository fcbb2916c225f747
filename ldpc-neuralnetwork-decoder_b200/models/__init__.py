"""Decoder classes of the engine (mirror of the reference's ldpc_neural_decoder.models)."""
from .layers import CheckLayer, VariableLayer, ResidualLayer, OutputLayer
from .traditional_decoders import BeliefPropagationDecoder, MinSumScaledDecoder

__all__ = ["CheckLayer", "VariableLayer", "ResidualLayer", "OutputLayer",
           "BeliefPropagationDecoder", "MinSumScaledDecoder"]
