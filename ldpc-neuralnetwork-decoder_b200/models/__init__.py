"""Decoder classes of the engine (mirror of the reference's ldpc_neural_decoder.models)."""
from .layers import CheckLayer, VariableLayer, ResidualLayer, OutputLayer
from .decoder import LDPCNeuralDecoder, TiedNeuralLDPCDecoder
from .traditional_decoders import BeliefPropagationDecoder, MinSumScaledDecoder
from .message_gnn_decoder import (MessageGNNLayer, MessageGNNDecoder, TannerToMessageGraph,
                                  create_message_gnn_decoder)

__all__ = ["CheckLayer", "VariableLayer", "ResidualLayer", "OutputLayer", "LDPCNeuralDecoder", "TiedNeuralLDPCDecoder",
           "BeliefPropagationDecoder", "MinSumScaledDecoder",
           "MessageGNNLayer", "MessageGNNDecoder", "TannerToMessageGraph", "create_message_gnn_decoder"]
