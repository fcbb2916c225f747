#!/usr/bin/env python
"""Data-parallel GNN training step on N GPUs (BASELINE config 5): one rank per GPU, codewords
sharded by rank, ONE NCCL all-reduce of the flat fp32 gradient per step.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_bench.py
Checks on the way that (a) every rank holds the same averaged gradient after the all-reduce and (b) that gradient equals
the mean of the per-rank gradients (gathered and averaged on rank 0), then reports aggregate codewords/s."""
import argparse, json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ldpc_b200
from ldpc_b200 import _native
from ldpc_b200.models import create_message_gnn_decoder, LDPCNeuralDecoder
from ldpc_b200.training import train_step, train_step_neural, allreduce_gradients
from ldpc_b200.utils import QCCode, create_LLR_mapping

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=512); ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--model", choices=["gnn", "neural"], default="gnn", help="gnn: MessageGNNDecoder (config 5); neural: LDPCNeuralDecoder")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
code = QCCode.nr_2_0(32)
torch.manual_seed(0)                                                    # same initial weights on every rank
B = a.batch
llr = torch.empty((B, code.N), dtype=torch.float32, device=dev)
_native.check(_native.lib().ldpc_awgn_llr(None, B, code.N, -2.0, 1, rank * B, _native.ptr(llr), _native.stream_ptr(dev)))   # rank's own frames
if a.model == "gnn":
    dec, _ = create_message_gnn_decoder(code, 5, 64, base_graph=code.base_matrix(), Z=32)
    dec = dec.to(dev)
    gt = torch.zeros((B, code.N), device=dev)
    forward = lambda: dec(llr, None, None, None, None, ground_truth=gt)[1]
    step = lambda: train_step(dec, llr, gt, opt)
else:
    _, cidx, vidx, oidx = create_LLR_mapping(code.dense().T)
    cidx, vidx = cidx.to(dev), vidx.to(dev)
    dec = LDPCNeuralDecoder(code.E, 5, 2, output_index_tensor=oidx).to(dev)
    llr = llr * 0.125                      # keeps sigmoid / BCE out of saturation so that the max loss has a gradient
    gt = torch.ones((B, code.N), device=dev)   # sigmoid(LLR > 0) -> 1 is the reference's output convention (layers.py:198)
    forward = lambda: dec(llr, cidx, vidx, gt)[1].mean()
    step = lambda: train_step_neural(dec, llr, cidx, vidx, gt, opt)

# gradient check: local gradient -> all-reduce -> compare with the gathered mean
loss = forward()
loss.backward()
local_flat = torch.cat([p.grad.reshape(-1) for p in dec.parameters()]).clone()
allreduce_gradients(dec)
avg_flat = torch.cat([p.grad.reshape(-1) for p in dec.parameters()])
check = {"ranks_agree": True, "equals_mean": True}
if world > 1:
    gathered = [torch.empty_like(local_flat) for _ in range(world)]
    dist.all_gather(gathered, local_flat)
    mean = torch.stack(gathered).mean(0)
    check["equals_mean"] = bool(torch.allclose(avg_flat, mean, rtol=1e-5, atol=1e-7 * float(mean.abs().max())))
    ref = avg_flat.clone()
    dist.broadcast(ref, 0)
    check["ranks_agree"] = bool(torch.equal(ref, avg_flat))
    flags = torch.tensor([int(check["ranks_agree"]), int(check["equals_mean"])], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    check = {"ranks_agree": bool(flags[0].item()), "equals_mean": bool(flags[1].item())}

opt = torch.optim.SGD(dec.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-4)
for _ in range(2):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"model": a.model, "train_step_ms": float(ms.item()), "n_gpus": world, "codewords_per_step": B * world,
                      "codewords_per_s": B * world / float(ms.item()) * 1e3, "loss_rank0": float(loss), **check}))
if world > 1:
    dist.destroy_process_group()
