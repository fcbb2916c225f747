"""BPSK-AWGN channel and BER/FER metrics on the device.

Mirror of the reference's ldpc_neural_decoder/utils/channel.py for the part on the hot path:
  AWGNChannel.transmit   channel.py:205-231   bits -> LLRs, sigma = 1/sqrt(10^(snr_db/10))
  compute_ber_fer        channel.py:156-190   (BER over all bits, FER = any bit wrong)
The reference draws noise with torch.randn on the host; here it comes from the engine's
counter-based Philox generator (csrc/channel.cuh), keyed by (seed, global frame index), so a
sweep gives the same frames whatever the batch split or GPU count.

QPSK (channel.py:4-154, the chain every shipped training / evaluation loop uses):
  qpsk_modulate / awgn_channel / qpsk_demodulate   same signatures and values, vectorised (no per-codeword loop)
  QPSKChannel.transmit                             the fused device path: bits -> LLRs in one kernel
"""
import math
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


class QPSKChannel:
    """bits -> LLRs through QPSK + AWGN in one kernel (ldpc_qpsk_llr): what qpsk_demodulate(awgn_channel(
    qpsk_modulate(bits), snr_db), snr_db) computes, with the engine's Philox noise.  `true_llr=False` keeps the
    reference's LLR scaling (2*r/noise_var, i.e. 1/sqrt(2) of the true LLR); True gives the exact LLR."""

    def __init__(self, seed=0, first_frame=0, true_llr=False):
        self.seed = int(seed)
        self.next_frame = int(first_frame)
        self.true_llr = bool(true_llr)

    def transmit(self, bits, snr_db):
        squeeze = bits.dim() == 1
        b2 = bits.unsqueeze(0) if squeeze else bits
        dev = _cuda_device(b2)
        B, N = b2.shape
        bits_u8 = b2.to(device=dev, dtype=torch.uint8).contiguous()
        out = torch.empty((B, N), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().ldpc_qpsk_llr(
                _native.ptr(bits_u8), B, N, float(snr_db), int(self.true_llr), self.seed, self.next_frame,
                _native.ptr(out), _native.stream_ptr(dev)))
        self.next_frame += B
        out = out.to(bits.device)
        return out.squeeze(0) if squeeze else out


def qpsk_modulate(bits):
    """channel.py:4-60: bit 0 -> +1/sqrt(2), bit 1 -> -1/sqrt(2); even positions ride on I, odd on Q; an odd
    number of bits is padded with one 0 bit.  (B, n) or (n,) -> complex (B, ceil(n/2)) or (ceil(n/2),)."""
    squeeze = bits.dim() == 1
    b2 = bits.reshape(1, -1) if squeeze else bits.reshape(bits.shape[0], -1)
    comp = 1 / math.sqrt(2) - b2.float() * math.sqrt(2)
    if comp.shape[1] % 2 == 1:
        # the reference pads with torch.tensor([1/np.sqrt(2)]) -- a float64 tensor -- so an odd-length block is
        # promoted to float64 / complex128 there; kept, so that the values are identical
        pad = torch.full((comp.shape[0], 1), 1 / math.sqrt(2), dtype=torch.float64, device=comp.device)
        comp = torch.cat([comp, pad], dim=1)
    sym = torch.complex(comp[:, 0::2].contiguous(), comp[:, 1::2].contiguous())
    return sym.squeeze(0) if squeeze else sym


def awgn_channel(symbols, snr_db):
    """channel.py:62-89: complex noise, each component N(0, 1/(2*snr_linear)) (torch.randn on the symbols' device)."""
    noise_power = 1 / (10 ** (snr_db / 10))
    nr = torch.randn(symbols.size(), device=symbols.device) * math.sqrt(noise_power / 2)
    ni = torch.randn(symbols.size(), device=symbols.device) * math.sqrt(noise_power / 2)
    return symbols + torch.complex(nr, ni)


def qpsk_demodulate(received_symbols, snr_db):
    """channel.py:91-154: llr = 2*r/noise_var per component, interleaved I, Q, I, Q ...  (the reference's scaling)."""
    squeeze = received_symbols.dim() == 1
    r2 = received_symbols.reshape(1, -1) if squeeze else received_symbols.reshape(received_symbols.shape[0], -1)
    noise_var = 1 / (10 ** (snr_db / 10))
    # computed in the symbols' precision, stored in a float32 buffer (channel.py:141-143)
    llr = torch.stack([2 * r2.real / noise_var, 2 * r2.imag / noise_var], dim=2).reshape(r2.shape[0], -1).to(torch.float32)
    return llr.squeeze(0) if squeeze else llr


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
