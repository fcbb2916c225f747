"""Multi-GPU host logic on CPU: frame-range sharding + counter all-reduce with 2 gloo ranks.
Each rank decodes ITS range of global frame indices (the oracle stands in for the GPU kernel,
same Philox keying) and the reduced counters must equal a single-rank run over all frames."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
import ldpc_b200
from ldpc_b200.sim import shard_range, reduce_counters, wilson_interval
from ldpc_b200.utils import QCCode

FRAMES, SNR_DB, ITERS, SEED = 301, -1.0, 5, 77


def local_counters(first, count):
    from oracle import oracle
    code = QCCode.nr_2_0(4)
    llr = oracle.awgn_llr(None, count, code.N, SNR_DB, SEED, first_frame=first)
    o = oracle.decode(code.shifts, 4, llr, ITERS, "minsum", 0.75, order="fast", want_mask=True)
    nerr = o["hard"].sum(axis=1)
    valid = ((o["valid_mask"][:, 0] >> np.uint64(ITERS - 1)) & np.uint64(1)).astype(bool)
    return torch.tensor([int(nerr.sum()), int((nerr > 0).sum()), count, int(((nerr > 0) & valid).sum())], dtype=torch.int64)


def worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(FRAMES, rank, world)
    c = local_counters(first, count)
    reduce_counters(c)
    if rank == 0:
        q.put(c.tolist())
    dist.destroy_process_group()


def grad_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import ldpc_b200  # noqa: F401
    from ldpc_b200.training import allreduce_gradients
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 1))
    for i, p in enumerate(m.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    list(m.parameters())[-1].grad = None if rank == 1 else torch.ones(1)     # a parameter without gradient on one rank
    assert allreduce_gradients(m) == world
    if rank == 0:
        q.put([p.grad.reshape(-1).tolist() for p in m.parameters()])
    dist.destroy_process_group()


def test_gradient_allreduce_two_gloo_ranks():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i, g in enumerate(got[:-1]):
        assert all(abs(x - 1.5 * (i + 1)) < 1e-6 for x in g)        # mean of rank 0 (x1) and rank 1 (x2)
    assert abs(got[-1][0] - 0.5) < 1e-6


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 301, 10 ** 9):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    lo, hi = wilson_interval(5, 1000)
    assert lo < 0.005 < hi and wilson_interval(0, 0) == (0.0, 1.0)


def test_two_rank_gloo_counters_equal_single_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = local_counters(0, FRAMES).tolist()
    assert got == want and got[2] == FRAMES


# ---- checkpointed sweep (sim.simulate_fer): killed after some blocks, resumed on another world size ----------------
def _oracle_block(code, algo, iters, alpha, snr_db, seed, first, count, dev, max_frames_per_call):
    """Stand-in for sim._run_block (the fused GPU kernel) with the same Philox keying, on the CPU oracle."""
    from oracle import oracle
    if count == 0:
        return torch.zeros(4, dtype=torch.int64)
    llr = oracle.awgn_llr(None, count, code.N, snr_db, seed, first_frame=first)
    o = oracle.decode(code.shifts, code.Z, llr, iters, algo, alpha, order="fast", want_mask=True)
    nerr = o["hard"].sum(axis=1)
    valid = ((o["valid_mask"][:, 0] >> np.uint64(iters - 1)) & np.uint64(1)).astype(bool)
    return torch.tensor([int(nerr.sum()), int((nerr > 0).sum()), count, int(((nerr > 0) & valid).sum())], dtype=torch.int64)


SWEEP = dict(snr_db_list=[-1.0, 0.5], frames=203, algo="minsum", iters=4, alpha=0.75, seed=5, block_frames=50)


def sweep_worker(rank, world, port, q, ckpt, stop_after):
    import sys
    sys.path.insert(0, ROOT)
    import ldpc_b200  # noqa: F401
    from ldpc_b200 import sim
    sim._run_block = _oracle_block
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = sim.simulate_fer(QCCode.nr_2_0(4), rank=rank, world=world, checkpoint=ckpt, stop_after_blocks=stop_after, **SWEEP)
    dist.barrier()
    if rank == 0:
        q.put(res)
    dist.destroy_process_group()


def test_sweep_checkpoint_resumes_on_another_world_size(tmp_path, monkeypatch):
    from ldpc_b200 import sim
    monkeypatch.setattr(sim, "_run_block", _oracle_block)
    code = QCCode.nr_2_0(4)
    want = sim.simulate_fer(code, **SWEEP)                                  # uninterrupted, one rank, no checkpoint
    assert [p["frames"] for p in want] == [203, 203] and want[0]["frame_errors"] > 0
    ckpt = str(tmp_path / "sweep.json")
    # 2 gloo ranks, "killed" after 3 of the 5 + 5 blocks
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=sweep_worker, args=(r, 2, port, q, ckpt, 3)) for r in range(2)]
    for p in procs:
        p.start()
    assert q.get(timeout=180) is None
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    st = sim.load_checkpoint(ckpt, sim.sweep_signature(code, SWEEP["snr_db_list"], 203, "minsum", 4, 0.75, 5))
    assert st["current"]["point"] == 0 and st["current"]["frames_done"] == 150 and st["points"] == []
    # resumed on ONE rank, killed again inside point 1, then finished: identical to the uninterrupted sweep
    assert sim.simulate_fer(code, checkpoint=ckpt, stop_after_blocks=4, **SWEEP) is None
    st = sim.load_checkpoint(ckpt, st["signature"])
    assert len(st["points"]) == 1 and st["current"] == {"point": 1, "frames_done": 100, "counters": st["current"]["counters"]}
    got = sim.simulate_fer(code, checkpoint=ckpt, **SWEEP)
    assert got == want
    assert sim.simulate_fer(code, checkpoint=ckpt, **SWEEP) == want         # a finished sweep replays from the file
    with pytest.raises(ValueError):                                         # a different sweep must not adopt this file
        sim.simulate_fer(code, checkpoint=ckpt, **dict(SWEEP, seed=6))
