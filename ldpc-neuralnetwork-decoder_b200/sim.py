"""Monte-Carlo BER/FER sweeps on the engine's fused simulation path.

Replaces the loop shape of ComparativeEvaluator._evaluate_traditional_decoder
(training/comparative_evaluation.py:108-166: SNR list x trials of all-zero codewords through
the channel and a decoder, averaged BER/FER) for the classic decoders: frames are generated,
decoded and counted inside one kernel (ldpc_sim_fer); only four integer counters leave the GPU.

Multi-GPU: frames are independent, so rank r of W takes a contiguous range of GLOBAL frame
indices; the Philox counter is the global index, so the totals do not depend on W.  The only
collective is one all-reduce (sum) of the int64 counters per SNR point.
"""
import math

import torch

from . import _native


def shard_range(total, rank, world):
    """Contiguous [first, first+count) of `total` frames owned by `rank` (sizes differ by <= 1)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(int(total), world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def reduce_counters(counters, group=None):
    """Sum the [bit errors, frame errors, frames, undetected] counters over all ranks (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def wilson_interval(k, n, z=1.96):
    """95 % Wilson score interval of a proportion (for 'statistically indistinguishable' FER claims)."""
    if n == 0:
        return 0.0, 1.0
    p = k / n
    d = 1 + z * z / n
    c = p + z * z / (2 * n)
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4 * n * n))
    return max(0.0, (c - h) / d), min(1.0, (c + h) / d)


def simulate_fer(code, snr_db_list, frames, algo="minsum", iters=10, alpha=0.75, seed=1234, device=None,
                 rank=0, world=1, group=None, max_frames_per_call=1 << 24):
    """Sweep `snr_db_list`; `frames` all-zero codewords per point in total over all ranks.
    Returns a list of dicts: snr_db, frames, bit_errors, frame_errors, undetected, ber, fer, fer_ci."""
    dev = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
    L = _native.lib()
    h = code.handle(dev)
    a = _native.ALGO_MINSUM if algo == "minsum" else _native.ALGO_BP
    first, count = shard_range(frames, rank, world)
    out = []
    for point, snr_db in enumerate(snr_db_list):
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
        done = 0
        with torch.cuda.device(dev):
            while done < count:
                n = min(max_frames_per_call, count - done)
                _native.check(L.ldpc_sim_fer(h, a, int(iters), float(alpha), float(snr_db), int(seed) + point,
                                             first + done, n, _native.ptr(counters), _native.stream_ptr(dev)))
                done += n
        reduce_counters(counters, group)
        be, fe, fr, und = (int(x) for x in counters.tolist())
        out.append(dict(snr_db=float(snr_db), frames=fr, bit_errors=be, frame_errors=fe, undetected=und,
                        ber=be / max(fr * code.N, 1), fer=fe / max(fr, 1), fer_ci=wilson_interval(fe, fr)))
    return out
