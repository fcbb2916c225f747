"""Device channel / counters / fused simulation and the edge-space layers against the oracle
and the reference's golden vectors.  Needs a B200 (-m gpu)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import oracle
import ldpc_b200
from ldpc_b200 import _native
from ldpc_b200.models import CheckLayer, VariableLayer, ResidualLayer, OutputLayer
from ldpc_b200.utils import QCCode, AWGNChannel, QPSKChannel, compute_ber_fer, count_errors

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_awgn_generator_matches_oracle():
    """Same Philox stream; fp32 log/sin/cos differ by a few ulp between CUDA and libm, so the
    LLRs agree to 2e-5 relative of sigma-scale, not bit for bit (tolerance stated here)."""
    N = 1664
    for snr_db in (-4.15, 0.0, 2.0):
        ch = AWGNChannel(seed=1234)
        llr = ch.transmit(torch.zeros(96, N, device=DEV), snr_db).cpu().numpy()
        ref = oracle.awgn_llr(None, 96, N, snr_db, 1234)
        scale = 2.0 / (10 ** (-snr_db / 10))
        assert np.max(np.abs(llr - ref)) <= 2e-5 * scale * 8
        nxt = ch.transmit(torch.zeros(4, N, device=DEV), snr_db).cpu().numpy()      # frame counter advanced
        assert np.max(np.abs(nxt - oracle.awgn_llr(None, 4, N, snr_db, 1234, first_frame=96))) <= 2e-5 * scale * 8
    bits = (torch.rand(8, 208, device=DEV) > 0.5).float()
    llr = AWGNChannel(seed=5).transmit(bits, 3.0).cpu().numpy()
    ref = oracle.awgn_llr(bits.cpu().numpy().astype(np.uint8), 8, 208, 3.0, 5)
    assert np.max(np.abs(llr - ref)) <= 1e-3
    one = AWGNChannel(seed=5).transmit(bits[0], 3.0)
    assert one.shape == (208,)


def test_qpsk_generator_matches_oracle():
    """ldpc_qpsk_llr == oracle_qpsk_llr (same Philox stream, same fp32 operation order; libm vs CUDA log/sincos
    differ in the last ulp -> 2e-5 of the LLR scale), for random bits, both LLR scalings, ragged N."""
    for N, snr_db, true_llr in ((1664, -2.0, False), (208, 1.5, False), (51, 6.0, True)):
        bits = (torch.rand(33, N, device=DEV) > 0.5).float()
        ch = QPSKChannel(seed=77, true_llr=true_llr)
        llr = ch.transmit(bits, snr_db).cpu().numpy()
        ref = oracle.qpsk_llr(bits.cpu().numpy().astype(np.uint8), 33, N, snr_db, 77, true_llr=true_llr)
        scale = 2.0 / (10 ** (-snr_db / 10)) * (np.sqrt(2) if true_llr else 1.0)
        assert np.max(np.abs(llr - ref)) <= 2e-5 * scale * 8
        nxt = ch.transmit(bits[:5], snr_db).cpu().numpy()                              # frame counter advanced
        ref2 = oracle.qpsk_llr(bits[:5].cpu().numpy().astype(np.uint8), 5, N, snr_db, 77, first_frame=33, true_llr=true_llr)
        assert np.max(np.abs(nxt - ref2)) <= 2e-5 * scale * 8
    # min-sum is scale-invariant: decoding the reference-scaled and the true LLRs of the same frames gives the same bits
    code = QCCode.nr_2_0(32)
    from ldpc_b200.models import MinSumScaledDecoder
    dec = MinSumScaledDecoder(code, 10, 0.75, early_stopping=False)
    zeros = torch.zeros(256, code.N, device=DEV)
    a, _ = dec.decode(QPSKChannel(seed=5).transmit(zeros, -2.0))
    b, _ = dec.decode(QPSKChannel(seed=5, true_llr=True).transmit(zeros, -2.0))
    assert torch.equal(a, b)


def test_compute_ber_fer_known_answers():
    g = load_golden("mapping_layers")
    tx, rx = torch.from_numpy(g["berfer_tx"]).to(DEV), torch.from_numpy(g["berfer_rx"]).to(DEV)
    ber, fer = compute_ber_fer(tx, rx)
    assert abs(ber - g["berfer"][0]) < 1e-7 and abs(fer - g["berfer"][1]) < 1e-7
    c = count_errors(tx, rx)
    assert c.tolist()[:3] == [5, 2, 4]
    assert compute_ber_fer(tx.cpu(), rx.cpu()) == (ber, fer)
    with pytest.raises(AssertionError):
        compute_ber_fer(tx, rx[:, :5])


@pytest.mark.parametrize("Z,algo,frames", [(32, "minsum", 5000), (4, "minsum", 4001), (32, "bp", 600), (16, "minsum", 700)])
def test_fused_simulation_counts(Z, algo, frames):
    """ldpc_sim_fer == generate (oracle Philox) -> decode (oracle) -> count, up to the frames
    whose decision flips because the device's log/sincos differ in the last ulp (allow 0.2 %)."""
    code = QCCode.nr_2_0(Z)
    snr_db, iters, seed = (-2.5 if Z == 32 else 0.5), 8, 42
    counters = torch.zeros(4, dtype=torch.int64, device=DEV)
    a = _native.ALGO_MINSUM if algo == "minsum" else _native.ALGO_BP
    # two calls covering [0, frames): the global frame index keys the noise
    split = frames // 3
    for first, n in ((0, split), (split, frames - split)):
        _native.check(_native.lib().ldpc_sim_fer(code.handle(DEV), a, iters, 0.75, snr_db, seed, first, n,
                                                 _native.ptr(counters), _native.stream_ptr(DEV)))
    got = counters.tolist()
    llr = oracle.awgn_llr(None, frames, code.N, snr_db, seed)
    order = "fast" if (algo == "minsum" and Z in (4, 32)) else "reference"
    o = oracle.decode(code.shifts, Z, llr, iters, algo, 0.75, order=order, want_mask=True)
    nerr = o["hard"].sum(axis=1)
    valid = ((o["valid_mask"][:, 0] >> np.uint64(iters - 1)) & np.uint64(1)).astype(bool)
    want = [int(nerr.sum()), int((nerr > 0).sum()), frames, int(((nerr > 0) & valid).sum())]
    assert got[2] == frames
    assert abs(got[1] - want[1]) <= max(2, 0.002 * frames)
    assert abs(got[0] - want[0]) <= max(50, 0.01 * want[0])
    assert abs(got[3] - want[3]) <= 2


def test_layers_forward_match_reference_golden():
    g = load_golden("mapping_layers")
    t = lambda k: torch.from_numpy(g[k]).to(DEV)
    cidx, vidx = t("z4_check").long(), t("z4_var").long()
    c2v = CheckLayer()(t("lay_x"), cidx)
    assert np.array_equal(c2v.cpu().numpy(), g["lay_c2v"])
    v2c = VariableLayer()(t("lay_llr"), t("lay_c2v"), vidx)
    np.testing.assert_allclose(v2c.cpu().numpy(), g["lay_v2c"], rtol=1e-5, atol=1e-5)
    res = ResidualLayer(g["lay_x"].shape[1], depth_L=2).to(DEV)
    with torch.no_grad():
        res.w_ch.copy_(t("lay_wch"))
        res.w_res.copy_(t("lay_wres"))
    r = res(t("lay_llr"), t("lay_c2v"), list(t("lay_prev")))
    np.testing.assert_allclose(r.detach().cpu().numpy(), g["lay_res"], rtol=1e-6, atol=1e-6)
    soft, ml = OutputLayer()(t("lay_res") * 0.1, t("lay_llr"), t("lay_gt"))
    np.testing.assert_allclose(soft.cpu().numpy(), g["lay_soft"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ml.cpu().numpy(), g["lay_maxloss"], rtol=1e-5)
    soft2, none = OutputLayer()(t("lay_res") * 0.1, t("lay_llr"))
    assert none is None and torch.equal(soft2, soft)


def test_layers_backward_match_reference_autograd():
    g = load_golden("mapping_layers")
    t = lambda k: torch.from_numpy(g[k]).to(DEV)
    cidx, vidx = t("z4_check").long(), t("z4_var").long()
    x = t("lay_x").requires_grad_(True)
    llr = t("lay_llr").requires_grad_(True)
    res = ResidualLayer(g["lay_x"].shape[1], depth_L=2).to(DEV)
    with torch.no_grad():
        res.w_ch.copy_(t("lay_wch"))
        res.w_res.copy_(t("lay_wres"))
    prev = t("lay_prev")
    c2v = CheckLayer()(x, cidx)
    v2c = VariableLayer()(llr, c2v, vidx)
    r = res(llr, c2v, [v2c, prev[1]])
    soft, ml = OutputLayer()(r * 0.1, llr, t("lay_gt"))
    np.testing.assert_allclose(ml.detach().cpu().numpy(), g["lay_g_maxloss"], rtol=1e-5)
    ml.sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["lay_grad_x"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(llr.grad.cpu().numpy(), g["lay_grad_llr"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(res.w_ch.grad.cpu().numpy(), g["lay_grad_wch"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(res.w_res.grad.cpu().numpy(), g["lay_grad_wres"], rtol=1e-4, atol=1e-5)


def test_layers_at_bg2_z32_against_oracle():
    """One check+variable pass at the headline code size (E = 6304, K = 9 / 22)."""
    from ldpc_b200.utils import create_LLR_mapping
    code = QCCode.nr_2_0(32)
    _, cidx, vidx, _ = create_LLR_mapping(code.dense().T)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(37, code.E)).astype(np.float32) * 4
    x[3, ::7] = 0.0
    llr = rng.normal(size=(37, code.E)).astype(np.float32)
    c2v = CheckLayer()(torch.from_numpy(x).to(DEV), cidx.to(DEV))
    ref_c2v = oracle.check_layer(x, cidx.numpy())
    assert np.array_equal(c2v.cpu().numpy(), ref_c2v)
    v2c = VariableLayer()(torch.from_numpy(llr).to(DEV), c2v, vidx.to(DEV))
    np.testing.assert_allclose(v2c.cpu().numpy(), oracle.variable_layer(llr, ref_c2v, vidx.numpy()), rtol=1e-5, atol=1e-4)
    with pytest.raises(RuntimeError):
        CheckLayer()(torch.from_numpy(x), cidx)


def test_sweep_driver_is_independent_of_sharding():
    """sim.simulate_fer: counters of rank 0 + rank 1 (emulated, world=2) == world=1; frames are
    keyed by the global frame index, so FER does not depend on the GPU count."""
    from ldpc_b200.sim import simulate_fer
    code = QCCode.nr_2_0(32)
    snrs = [-3.0, -2.0]
    one = simulate_fer(code, snrs, 6001, iters=6, seed=5, device=DEV, max_frames_per_call=2500)
    parts = [simulate_fer(code, snrs, 6001, iters=6, seed=5, device=DEV, rank=r, world=2) for r in range(2)]
    for k, point in enumerate(one):
        assert point["frames"] == 6001
        for key in ("bit_errors", "frame_errors", "frames", "undetected"):
            assert point[key] == parts[0][k][key] + parts[1][k][key]
        lo, hi = point["fer_ci"]
        assert lo <= point["fer"] <= hi
    assert one[0]["fer"] > one[1]["fer"] > 0


@pytest.mark.parametrize("Z", [4, 32])
def test_encoder_kernel_and_round_trip(Z):
    """ldpc_encode == the oracle's dense GF(2) solve bit for bit; every codeword passes the engine's
    own syndrome check; round trip info -> encode -> channel (real bits) -> decode recovers the codeword; and
    min-sum is symmetric: flipping the channel by a codeword flips the decisions by that codeword."""
    from ldpc_b200.models import MinSumScaledDecoder
    from ldpc_b200.utils import SystematicEncoder
    code = QCCode.nr_2_0(Z)
    enc = SystematicEncoder(code)
    rng = np.random.default_rng(100 + Z)
    for B in (1, 37):
        info = rng.integers(0, 2, size=(B, code.K), dtype=np.uint8)
        cw = enc.encode(torch.from_numpy(info).to(DEV))
        assert cw.dtype == torch.float32 and cw.shape == (B, code.N)
        assert np.array_equal(cw.cpu().numpy().astype(np.uint8), oracle.encode_dense(code.shifts, Z, info))
    dec = MinSumScaledDecoder(code, 20, 0.75, early_stopping=False)
    assert bool(dec._check_valid_codeword(cw).all())
    one = enc.encode(torch.from_numpy(info[0]))                        # CPU tensor, un-batched: staged through the GPU
    assert one.shape == (code.N,) and torch.equal(one, cw[0].cpu())
    # round trip at a comfortable SNR, both channels, real (non-zero) codewords
    for ch in (AWGNChannel(seed=3), QPSKChannel(seed=3)):
        bits, _ = dec.decode(ch.transmit(cw, 2.0))
        assert torch.equal(bits, cw)
    # symmetry at a noisy SNR: decode(llr) == decode(llr with the codeword's signs removed) XOR codeword
    llr_c = AWGNChannel(seed=9).transmit(cw, -3.0)
    llr_0 = llr_c * (1.0 - 2.0 * cw)                                   # what the all-zero codeword would have received
    b_c, _ = dec.decode(llr_c)
    b_0, _ = dec.decode(llr_0)
    assert torch.equal(b_c, (b_0 + cw) % 2)
    with pytest.raises(ValueError):
        enc.encode(torch.zeros(2, code.K + 1))
