"""Evaluation driver on the engine.

Mirror of the reference's `training/comparative_evaluation.py:ComparativeEvaluator` (:10-222) -- same constructor,
`evaluate_all` signature and result dictionary (:86-104) -- so that code written against it runs unchanged:

    results = {'snr_range': [...],
               'belief_propagation': {'ber': [...], 'fer': [...], 'avg_iterations': [...]},
               'min_sum_scaled':     {'ber': [...], 'fer': [...], 'avg_iterations': [...]},
               'neural_decoder':     {'ber': [...], 'fer': [...]}}          # only with a neural decoder

Per SNR point and trial (comparative_evaluation.py:129-160): all-zero codewords -> QPSK -> AWGN -> LLRs ->
`decoder.decode` (batch-global early stopping, max 50 iterations) -> `compute_ber_fer`; the lists hold the means over
the trials.  (The reference module itself cannot be imported: it needs `models/decoder.py`, which is absent from the
repository; the plotting methods are out of scope here, see `visualization/`.)

What differs, deliberately: the channel is the fused device generator (`QPSKChannel`: one kernel instead of the
reference's per-codeword Python loops, noise keyed by (seed, global frame index) so a sweep is reproducible and
independent of the batch split), every step runs on the GPU, and `evaluate_decoder_per_codeword` adds the sweep that
`run_comparison_all.py:320-345` wants (`decode_with_iterations`, per-codeword early exit).
"""
import torch

from .models.traditional_decoders import BeliefPropagationDecoder, MinSumScaledDecoder
from .utils.channel import QPSKChannel, compute_ber_fer, count_errors


class ComparativeEvaluator:
    def __init__(self, H=None, neural_decoder=None, device="cuda", base_graph=None, Z=None, seed=0, channel=None,
                 max_iterations=50, scaling_factor=0.75, path="auto"):
        """`H` (dense parity-check matrix, factored into QC shifts) or `(base_graph, Z)`.  `channel`: any object with
        `transmit(bits, snr_db) -> llrs`; default `QPSKChannel(seed)` with the reference's LLR scaling.  `path`: kernel
        policy of both classic decoders ("auto": specialised kernels for min-sum, exact kernel for BP; "fast": specialised
        kernels for both -- BP then uses CUDA's tanhf/atanhf instead of once-rounded double evaluation; "exact")."""
        if not torch.cuda.is_available() and channel is None:
            raise RuntimeError("the LDPC engine needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.H = H
        self.neural_decoder = neural_decoder
        if neural_decoder is not None and hasattr(neural_decoder, "eval"):
            neural_decoder.to(self.device)
            neural_decoder.eval()
        self.bp_decoder = BeliefPropagationDecoder(H, max_iterations=max_iterations, early_stopping=True,
                                                   base_graph=base_graph, Z=Z, path=path)
        self.ms_decoder = MinSumScaledDecoder(H, max_iterations=max_iterations, scaling_factor=scaling_factor,
                                              early_stopping=True, base_graph=base_graph, Z=Z, path=path)
        self.channel = channel if channel is not None else QPSKChannel(seed=seed)
        self.results = {}

    # comparative_evaluation.py:40-104
    def evaluate_all(self, snr_range, batch_size=32, num_trials=100, variable_bit_length=None, check_index_tensor=None,
                     var_index_tensor=None):
        if variable_bit_length is None:
            variable_bit_length = self.bp_decoder.code.N
        bp = self._evaluate_traditional_decoder(self.bp_decoder, snr_range, batch_size, num_trials, variable_bit_length)
        ms = self._evaluate_traditional_decoder(self.ms_decoder, snr_range, batch_size, num_trials, variable_bit_length)
        self.results = {
            "snr_range": snr_range,
            "belief_propagation": {"ber": bp[0], "fer": bp[1], "avg_iterations": bp[2]},
            "min_sum_scaled": {"ber": ms[0], "fer": ms[1], "avg_iterations": ms[2]},
        }
        if self.neural_decoder is not None:
            nb, nf = self._evaluate_neural_decoder(snr_range, batch_size, num_trials, variable_bit_length,
                                                   check_index_tensor, var_index_tensor)
            self.results["neural_decoder"] = {"ber": nb, "fer": nf}
        return self.results

    def _llrs(self, batch_size, variable_bit_length, snr_db):
        tx = torch.zeros((batch_size, variable_bit_length), device=self.device)
        return tx, self.channel.transmit(tx, snr_db).view(batch_size, -1)

    # comparative_evaluation.py:106-166
    def _evaluate_traditional_decoder(self, decoder, snr_range, batch_size, num_trials, variable_bit_length):
        ber_results, fer_results, avg_iterations = [], [], []
        for snr_db in snr_range:
            total_ber = total_fer = 0.0
            total_iterations = 0
            for _ in range(num_trials):
                tx, llrs = self._llrs(batch_size, variable_bit_length, snr_db)
                decoded_bits, iterations = decoder.decode(llrs)
                ber, fer = compute_ber_fer(tx, decoded_bits)
                total_ber += ber
                total_fer += fer
                total_iterations += iterations
            ber_results.append(total_ber / num_trials)
            fer_results.append(total_fer / num_trials)
            avg_iterations.append(total_iterations / num_trials)
        return ber_results, fer_results, avg_iterations

    # comparative_evaluation.py:168-222
    def _evaluate_neural_decoder(self, snr_range, batch_size, num_trials, variable_bit_length, check_index_tensor=None,
                                 var_index_tensor=None):
        extra = [t.to(self.device) for t in (check_index_tensor, var_index_tensor) if t is not None]
        ber_results, fer_results = [], []
        with torch.no_grad():
            for snr_db in snr_range:
                total_ber = total_fer = 0.0
                for _ in range(num_trials):
                    tx, llrs = self._llrs(batch_size, variable_bit_length, snr_db)
                    hard_bits = self.neural_decoder.decode(llrs, *extra)
                    ber, fer = compute_ber_fer(tx, hard_bits)
                    total_ber += ber
                    total_fer += fer
                ber_results.append(total_ber / num_trials)
                fer_results.append(total_fer / num_trials)
        return ber_results, fer_results

    # run_comparison_all.py:320-345: per-codeword early exit, integer counters, one synchronisation per SNR point
    def evaluate_decoder_per_codeword(self, decoder, snr_range, batch_size=4096, num_trials=1):
        """{'snr_range', 'ber', 'fer', 'avg_iterations', 'undetected'}: every codeword stops at its own first valid
        iteration (`decode_with_iterations`); `undetected` counts frames that passed the syndrome check with wrong
        bits.  Counters stay on the device until the SNR point is finished."""
        N = decoder.code.N
        out = {"snr_range": list(snr_range), "ber": [], "fer": [], "avg_iterations": [], "undetected": []}
        for snr_db in snr_range:
            counters = torch.zeros(4, dtype=torch.int64, device=self.device)
            it_sum = torch.zeros((), dtype=torch.int64, device=self.device)
            undetected = torch.zeros((), dtype=torch.int64, device=self.device)
            for _ in range(num_trials):
                tx, llrs = self._llrs(batch_size, N, snr_db)
                bits, its, ok = decoder.decode_with_iterations(llrs)
                count_errors(tx, bits, counters)
                it_sum += its.to(torch.int64).sum()
                undetected += (ok & (bits != 0).any(dim=1)).sum()
            c = counters.tolist()
            frames = max(c[2], 1)
            out["ber"].append(c[0] / (frames * N))
            out["fer"].append(c[1] / frames)
            out["avg_iterations"].append(it_sum.item() / frames)
            out["undetected"].append(int(undetected.item()))
        return out

    def save_results(self, path):
        """`torch.save` of the results dictionary (what plot_comparison.py / run_comparison*.py load)."""
        if not self.results:
            raise ValueError("No results to save. Run evaluate_all() first.")
        torch.save(self.results, path)
