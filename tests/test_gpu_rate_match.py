"""Rate matching / punctured LLR layout on the device against the specification-loop oracle,
and end to end: payload -> filler padding -> encode -> rate match -> channel -> rate recover -> decode -> payload."""
import numpy as np
import pytest
import torch

from oracle import oracle
import ldpc_b200
from ldpc_b200.models import MinSumScaledDecoder, BeliefPropagationDecoder
from ldpc_b200.utils import QCCode, RateMatcher, SystematicEncoder

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.mark.parametrize("Z,Kp,E,rv,Qm,Ncb", [(4, 33, 128, 0, 2, None), (4, 40, 500, 2, 4, None), (32, 300, 2046, 1, 6, None),
                                             (32, 320, 1600, 3, 1, 1200), (32, 270, 4000, 0, 8, None)])
def test_device_rate_match_and_recover_equal_the_specification_loops(Z, Kp, E, rv, Qm, Ncb):
    code = QCCode.nr_2_0(Z)
    rm = RateMatcher(code, E, payload_bits=Kp, rv=rv, Qm=Qm, Ncb=Ncb, filler_llr=77.0)
    rng = np.random.default_rng(E)
    B = 5
    cw = rng.integers(0, 2, (B, code.N)).astype(np.float32)
    tx = rm.rate_match(torch.from_numpy(cw).to(DEV)).cpu().numpy()
    rx = rng.normal(size=(B, E)).astype(np.float32)
    llr = rm.rate_recover(torch.from_numpy(rx).to(DEV)).cpu().numpy()
    for b in range(B):
        assert np.array_equal(tx[b].astype(np.uint8), oracle.rate_match_38212(Z, code.cols, code.K, Kp, cw[b], E, rv, Qm, Ncb))
        want = oracle.rate_recover_38212(Z, code.cols, code.K, Kp, rx[b], E, rv, Qm, Ncb, filler_llr=77.0)
        assert np.array_equal(llr[b], want)                                   # same fp32 additions in the same order
    assert (llr[:, :2 * Z] == 0).all() and (llr[:, Kp:code.K] == 77.0).all()
    assert rm.rate_match(torch.zeros((0, code.N), device=DEV)).shape == (0, E)
    with pytest.raises(ValueError):
        rm.rate_recover(torch.zeros((1, E + 1), device=DEV))


@pytest.mark.parametrize("algo", ["minsum", "bp"])
def test_punctured_rate_matched_round_trip_recovers_the_payload(algo):
    """Real payloads through the whole 5G-shaped chain at a rate the code can carry: 300 payload bits (20 fillers),
    E = 1200 transmitted bits (rate 1/4 incl. the punctured 64 systematic bits), BPSK-AWGN at 2 dB, rv 0, Qm = 2."""
    code = QCCode.nr_2_0(32)
    Kp, E, B = 300, 1200, 256
    rm = RateMatcher(code, E, payload_bits=Kp, rv=0, Qm=2)
    g = torch.Generator().manual_seed(3)
    payload = torch.randint(0, 2, (B, Kp), generator=g).float()
    cw = SystematicEncoder(code).encode(rm.pad_info(payload).to(DEV))
    assert (cw[:, Kp:code.K] == 0).all()
    tx = rm.rate_match(cw)
    sigma = 10 ** (-2.0 / 20)
    noise = torch.randn((B, E), generator=g).to(DEV) * sigma
    rx_llr = 2.0 * ((1.0 - 2.0 * tx) + noise) / (sigma * sigma)
    llr = rm.rate_recover(rx_llr)
    assert (llr[:, :64] == 0).all() and (llr[:, Kp:code.K] == 1e4).all() and (llr[:, 64 + E + 20:] == 0).all()
    dec = (MinSumScaledDecoder(code, 20, 0.75, early_stopping=False) if algo == "minsum"
           else BeliefPropagationDecoder(code, 20, early_stopping=False))
    soft, hard = dec.forward(llr)
    got = hard[:, :Kp].cpu()
    frame_ok = (got == payload).all(dim=1)
    assert frame_ok.float().mean() >= 0.97, float(frame_ok.float().mean())
    # the punctured systematic bits were never sent and are recovered by the code
    assert (hard[frame_ok.to(DEV)][:, :64].cpu() == payload[frame_ok][:, :64]).all()
    # and the decoder's output on the layout equals the reference-order oracle fed the same vector (min-sum, hard bits)
    if algo == "minsum":
        o = oracle.decode(code.shifts, 32, llr[:16].cpu().numpy(), 20, "minsum", 0.75)
        assert np.array_equal(o["hard"], hard[:16].cpu().numpy().astype(np.uint8))
