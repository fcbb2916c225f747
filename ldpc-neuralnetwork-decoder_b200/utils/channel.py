"""BPSK-AWGN channel and BER/FER metrics on the device.

Mirror of the reference's ldpc_neural_decoder/utils/channel.py for the part on the hot path:
  AWGNChannel.transmit   channel.py:205-231   bits -> LLRs, sigma = 1/sqrt(10^(snr_db/10))
  compute_ber_fer        channel.py:156-190   (BER over all bits, FER = any bit wrong)
The reference draws noise with torch.randn on the host; here it comes from the engine's
counter-based Philox generator (csrc/channel.cuh), keyed by (seed, global frame index), so a
sweep gives the same frames whatever the batch split or GPU count.  The QPSK helpers of the
reference (channel.py:4-154) are outside the hot path and not provided.
"""
import torch

from .. import _native


def _cuda_device(t):
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("the LDPC engine needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class AWGNChannel:
    """`transmit(bits, snr_db)` as in the reference; `seed`/`first_frame` select the noise."""

    def __init__(self, seed=0, first_frame=0):
        self.seed = int(seed)
        self.next_frame = int(first_frame)

    def transmit(self, bits, snr_db):
        squeeze = bits.dim() == 1
        b2 = bits.unsqueeze(0) if squeeze else bits
        dev = _cuda_device(b2)
        B, N = b2.shape
        bits_u8 = b2.to(device=dev, dtype=torch.uint8).contiguous()
        out = torch.empty((B, N), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().ldpc_awgn_llr(
                _native.ptr(bits_u8), B, N, float(snr_db), self.seed, self.next_frame,
                _native.ptr(out), _native.stream_ptr(dev)))
        self.next_frame += B
        out = out.to(bits.device)
        return out.squeeze(0) if squeeze else out


def count_errors(transmitted_bits, decoded_bits, counters=None):
    """Integer counters [bit errors, frame errors, frames] (+= into `counters` if given)."""
    if transmitted_bits is not None and transmitted_bits.shape != decoded_bits.shape:
        raise AssertionError("Transmitted and decoded bits must have the same shape")
    d2 = decoded_bits.unsqueeze(0) if decoded_bits.dim() == 1 else decoded_bits
    dev = _cuda_device(d2)
    B, N = d2.shape
    if d2.dtype == torch.uint8:
        hard, dtype = d2.to(dev).contiguous(), _native.HARD_U8
    else:
        hard, dtype = d2.to(device=dev, dtype=torch.float32).contiguous(), _native.HARD_F32
    tx = None
    if transmitted_bits is not None:
        t2 = transmitted_bits.unsqueeze(0) if transmitted_bits.dim() == 1 else transmitted_bits
        tx = t2.to(device=dev, dtype=torch.uint8).contiguous()
    if counters is None:
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _native.check(_native.lib().ldpc_count_errors(
            _native.ptr(hard), dtype, _native.ptr(tx), B, N, _native.ptr(counters), _native.stream_ptr(dev)))
    return counters


def compute_ber_fer(transmitted_bits, decoded_bits):
    """(BER, FER) as Python floats (channel.py:156-190)."""
    c = count_errors(transmitted_bits, decoded_bits).tolist()
    frames = max(c[2], 1)
    nbits = decoded_bits.shape[-1]
    return c[0] / (frames * nbits), c[1] / frames
