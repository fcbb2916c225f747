"""Evaluation driver on the GPU: the ComparativeEvaluator-shaped sweep against an
independent computation with the oracle on the same LLRs, and the per-codeword sweep's invariants."""
import numpy as np
import pytest
import torch

from oracle import oracle
import ldpc_b200
from ldpc_b200.evaluation import ComparativeEvaluator
from ldpc_b200.models import MinSumScaledDecoder, create_message_gnn_decoder
from ldpc_b200.utils import QCCode, expand_base_matrix

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class OracleQPSK:
    """Channel stub: the oracle's QPSK LLRs (CPU), so that engine and expectation decode the very same floats."""

    def __init__(self, seed):
        self.seed, self.next_frame, self.log = seed, 0, []

    def transmit(self, bits, snr_db):
        B, N = bits.shape
        llr = oracle.qpsk_llr(None, B, N, snr_db, self.seed, first_frame=self.next_frame)
        self.next_frame += B
        self.log.append((snr_db, llr))
        return torch.from_numpy(llr).to(bits.device)


def expected_sweep(code, Z, algo, alpha, llr_batches, max_iters):
    """Reference semantics (comparative_evaluation.py:129-160 over traditional_decoders.py:102-106 / 255-258): a
    batch stops at the first iteration after which EVERY codeword is valid; BER / FER of that iteration's bits."""
    tot_ber = tot_fer = 0.0
    tot_it = 0
    for llr in llr_batches:
        full = oracle.decode(code.shifts, Z, llr, max_iters, algo, alpha, want_mask=True)
        t = oracle.first_all_valid(full["valid_mask"], max_iters)
        its = max_iters if t is None else t + 1
        hard = oracle.decode(code.shifts, Z, llr, its, algo, alpha)["hard"]
        tot_ber += hard.mean()
        tot_fer += hard.any(axis=1).mean()
        tot_it += its
    n = len(llr_batches)
    return tot_ber / n, tot_fer / n, tot_it / n


def test_evaluate_all_equals_the_reference_procedure_on_the_same_llrs():
    Z, B, trials, snrs, max_iters = 4, 24, 3, [0.0, 3.0], 12
    code = QCCode.nr_2_0(Z)
    ch = OracleQPSK(seed=9)
    ev = ComparativeEvaluator(base_graph=code.shifts, Z=Z, device=DEV, channel=ch, max_iterations=max_iters)
    res = ev.evaluate_all(snrs, batch_size=B, num_trials=trials)
    assert res["snr_range"] == snrs and set(res) == {"snr_range", "belief_propagation", "min_sum_scaled"}
    batches = [llr for _, llr in ch.log]
    assert len(batches) == 2 * len(snrs) * trials                       # BP sweep first, then min-sum, as the reference
    for k, (name, algo, alpha) in enumerate((("belief_propagation", "bp", 1.0), ("min_sum_scaled", "minsum", 0.75))):
        for i in range(len(snrs)):
            mine = batches[(k * len(snrs) + i) * trials:(k * len(snrs) + i + 1) * trials]
            ber, fer, its = expected_sweep(code, Z, algo, alpha, mine, max_iters)
            assert res[name]["ber"][i] == pytest.approx(ber, abs=1e-12)
            assert res[name]["fer"][i] == pytest.approx(fer, abs=1e-12)
            assert res[name]["avg_iterations"][i] == pytest.approx(its, abs=1e-12)
    import os, tempfile
    with tempfile.TemporaryDirectory() as d:
        ev.save_results(os.path.join(d, "r.pt"))
        assert torch.load(os.path.join(d, "r.pt"))["min_sum_scaled"]["fer"] == res["min_sum_scaled"]["fer"]


def test_device_channel_sweep_and_per_codeword_exit():
    """Default fused QPSK channel: reproducible (seed), FER falls with SNR, per-codeword exit needs no more
    iterations than the batch-global rule and yields the same error counts on frames that converged."""
    Z = 4
    code = QCCode.nr_2_0(Z)
    H = expand_base_matrix(torch.from_numpy(np.asarray(code.shifts, dtype=np.float32)), Z)
    a = ComparativeEvaluator(H, device=DEV, seed=3).evaluate_all([-6.0, -1.0], batch_size=64, num_trials=4)
    b = ComparativeEvaluator(base_graph=code.shifts, Z=Z, device=DEV, seed=3).evaluate_all([-6.0, -1.0], batch_size=64, num_trials=4)
    assert a == b                                                       # dense-H construction == (base graph, Z); same seed
    for name in ("belief_propagation", "min_sum_scaled"):
        assert a[name]["fer"][0] > a[name]["fer"][1] and a[name]["avg_iterations"][0] >= a[name]["avg_iterations"][1]
    ev = ComparativeEvaluator(base_graph=code.shifts, Z=Z, device=DEV, seed=3)
    pc = ev.evaluate_decoder_per_codeword(ev.ms_decoder, [-6.0, -1.0], batch_size=2048, num_trials=2)
    assert pc["fer"][0] > pc["fer"][1] and pc["avg_iterations"][0] > pc["avg_iterations"][1]
    assert all(1.0 <= x <= 50.0 for x in pc["avg_iterations"]) and pc["avg_iterations"][1] < 10
    # same frames through the oracle: per-codeword stop = first valid iteration of that codeword
    llr = oracle.qpsk_llr(None, 256, code.N, -4.0, seed=21)
    bits, its, ok = ev.ms_decoder.decode_with_iterations(torch.from_numpy(llr).to(DEV))
    # (path "auto": the specialised early-exit kernel, compiled for every lifting size since round 2 -> the kernel's operation order)
    o = oracle.decode(code.shifts, Z, llr, 50, "minsum", 0.75, order="fast", stop_when_valid=True)
    assert np.array_equal(bits.cpu().numpy().astype(np.uint8), o["hard"]) and np.array_equal(its.cpu().numpy(), o["iters"])
    # against the reference operation order: every codeword that converges within 20 iterations does so at the same iteration
    # with the same bits; later than that the flooding iteration has amplified the rounding difference of the two orders (about
    # 2x per iteration) and a late converger may stop a few iterations apart, or in one order only
    r = oracle.decode(code.shifts, Z, llr, 50, "minsum", 0.75, stop_when_valid=True)
    early = (o["iters"] <= 20) | (r["iters"] <= 20)
    assert early.sum() > 50 and np.array_equal(o["iters"][early], r["iters"][early]) and np.array_equal(o["hard"][early], r["hard"][early])
    assert (o["iters"] != r["iters"]).mean() < 0.1


def test_neural_decoder_slot_takes_the_gnn():
    Z = 4
    code = QCCode.nr_2_0(Z)
    torch.manual_seed(0)
    dec, _ = create_message_gnn_decoder(None, num_iterations=2, hidden_dim=64, base_graph=code.shifts, Z=Z)
    ev = ComparativeEvaluator(base_graph=code.shifts, Z=Z, device=DEV, neural_decoder=dec, seed=1, max_iterations=5)
    res = ev.evaluate_all([2.0], batch_size=16, num_trials=2)
    assert set(res["neural_decoder"]) == {"ber", "fer"} and 0.0 <= res["neural_decoder"]["ber"][0] <= 1.0


def test_neural_decoder_slot_takes_the_unrolled_min_sum_decoder():
    """run_comparison.py:85-110 call shape: LDPCNeuralDecoder(num_nodes, num_iterations, depth_L) in the evaluator's
    neural slot, index tensors passed to evaluate_all; (B, N) LLRs in, (B, N) bits out through output_index_tensor."""
    from ldpc_b200.models import LDPCNeuralDecoder
    from ldpc_b200.utils import create_LLR_mapping
    Z = 4
    code = QCCode.nr_2_0(Z)
    _, cidx, vidx, oidx = create_LLR_mapping(code.dense().T)
    dec = LDPCNeuralDecoder(num_nodes=code.E, num_iterations=5, depth_L=2, output_index_tensor=oidx)
    ev = ComparativeEvaluator(base_graph=code.shifts, Z=Z, device=DEV, neural_decoder=dec, seed=1, max_iterations=5)
    res = ev.evaluate_all([4.0], batch_size=64, num_trials=1, variable_bit_length=code.N,
                          check_index_tensor=cidx, var_index_tensor=vidx)
    nd = res["neural_decoder"]
    assert set(nd) == {"ber", "fer"} and 0.0 <= nd["ber"][0] <= 1.0
    # at 4 dB the unit-weight network is plain min-sum and decodes the all-zero codeword: with the reference's
    # `soft > 0.5` rule (trainer.py:186) on sigmoid(LLR) that reads as all ones, i.e. BER = 1 against zeros
    assert nd["ber"][0] > 0.95
