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


# ---- sweep checkpoint ---------------------------------------------------------------------------------------------
# A 10^9-frame sweep takes minutes per point; a killed job restarts from its last block, not from zero.  The state is
# (point, frames_done, counters) -- the Philox counter is the GLOBAL frame index, so nothing else is needed, and a sweep
# may resume on a different number of GPUs than it started on.
CHECKPOINT_VERSION = 1


def sweep_signature(code, snr_db_list, frames, algo, iters, alpha, seed):
    import hashlib
    return {"version": CHECKPOINT_VERSION, "rows": int(code.rows), "cols": int(code.cols), "Z": int(code.Z),
            "shifts_sha256": hashlib.sha256(code.shifts.astype("<i2").tobytes()).hexdigest(),
            "snr_db_list": [float(x) for x in snr_db_list], "frames": int(frames), "algo": str(algo), "iters": int(iters),
            "alpha": float(alpha), "seed": int(seed)}


def load_checkpoint(path, signature):
    """State of an interrupted sweep, or None.  A file written for a different sweep is an error, not a fresh start."""
    import json
    import os
    if not path or not os.path.exists(path):
        return None
    st = json.load(open(path))
    if st.get("signature") != signature:
        raise ValueError(f"checkpoint {path} belongs to a different sweep: {st.get('signature')} != {signature}")
    return st


def save_checkpoint(path, signature, points, point, frames_done, counters):
    import json
    import os
    tmp = f"{path}.tmp.{os.getpid()}"
    with open(tmp, "w") as f:
        json.dump({"signature": signature, "points": points,
                   "current": {"point": int(point), "frames_done": int(frames_done), "counters": [int(c) for c in counters]}}, f)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)                                  # atomic: a reader sees the old or the new state, never half


def _run_block(code, algo, iters, alpha, snr_db, seed, first, count, dev, max_frames_per_call):
    """[bit errors, frame errors, frames, undetected] of global frames [first, first+count) on this rank's GPU."""
    L = _native.lib()
    h = code.handle(dev)
    a = _native.ALGO_MINSUM if algo == "minsum" else _native.ALGO_BP
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    done = 0
    with torch.cuda.device(dev):
        while done < count:
            n = min(max_frames_per_call, count - done)
            _native.check(L.ldpc_sim_fer(h, a, int(iters), float(alpha), float(snr_db), int(seed), first + done, n,
                                         _native.ptr(counters), _native.stream_ptr(dev)))
            done += n
    return counters


def simulate_fer(code, snr_db_list, frames, algo="minsum", iters=10, alpha=0.75, seed=1234, device=None,
                 rank=0, world=1, group=None, max_frames_per_call=1 << 24, checkpoint=None, block_frames=1 << 26,
                 stop_after_blocks=None):
    """Sweep `snr_db_list`; `frames` all-zero codewords per point in total over all ranks.
    Returns a list of dicts: snr_db, frames, bit_errors, frame_errors, undetected, ber, fer, fer_ci.

    Every point is processed in global blocks of `block_frames` frames; a block is split over the ranks
    (`shard_range`), its counters are all-reduced, and with `checkpoint=<path>` rank 0 then records
    (point, frames_done, counters) atomically.  Calling again with the same arguments and path resumes after the last
    completed block -- on any number of ranks, since a frame's noise depends only on its global index -- and yields
    exactly the counters of an uninterrupted run.  `stop_after_blocks` (tests) ends the call early, as a kill would."""
    dev = None
    if device is not None or torch.cuda.is_available():
        dev = torch.device(device if device is not None else ("cuda", torch.cuda.current_device()))
    sig = sweep_signature(code, snr_db_list, frames, algo, iters, alpha, seed)
    state = load_checkpoint(checkpoint, sig) if checkpoint else None
    out = list(state["points"]) if state else []
    cur = state["current"] if state else None
    blocks_run = 0
    for point in range(len(out), len(snr_db_list)):
        snr_db = snr_db_list[point]
        done, total = 0, [0, 0, 0, 0]
        if cur and cur["point"] == point:
            done, total = int(cur["frames_done"]), [int(c) for c in cur["counters"]]
        while done < frames:
            nb = min(int(block_frames), frames - done)
            first, count = shard_range(nb, rank, world)
            blk = _run_block(code, algo, iters, alpha, snr_db, int(seed) + point, done + first, count, dev, max_frames_per_call)
            reduce_counters(blk, group)
            total = [t + int(c) for t, c in zip(total, blk.tolist())]
            done += nb
            blocks_run += 1
            if checkpoint and rank == 0:
                save_checkpoint(checkpoint, sig, out, point, done, total)
            if stop_after_blocks is not None and blocks_run >= stop_after_blocks and done < frames:
                return None
        be, fe, fr, und = total
        out.append(dict(snr_db=float(snr_db), frames=fr, bit_errors=be, frame_errors=fe, undetected=und,
                        ber=be / max(fr * code.N, 1), fer=fe / max(fr, 1), fer_ci=list(wilson_interval(fe, fr))))
        if checkpoint and rank == 0:
            save_checkpoint(checkpoint, sig, out, point + 1, 0, [0, 0, 0, 0])
        if stop_after_blocks is not None and blocks_run >= stop_after_blocks and point + 1 < len(snr_db_list):
            return None
    return out
