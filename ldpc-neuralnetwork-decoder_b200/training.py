"""One data-parallel training step of the message-GNN decoder (BASELINE.json config 5) and of the unrolled
neural min-sum decoder (`LDPCNeuralDecoder`).

The reference trains with SGD(momentum 0.9, weight decay 1e-4) on all-zero codewords
(training/trainer.py:70,231; the harness itself is unimportable).
Here a step is: engine forward with saved activations + engine backward (csrc/gnn_bwd.cuh) via
`loss.backward()`, ONE all-reduce of the flattened fp32 gradient (134 918 elements = 540 KB at
BG2 Z=32, latency-bound on NVLink), optimizer step.  Codewords are sharded by rank; the mean-BCE
loss is a per-rank mean, so gradients are averaged over ranks.
"""
import torch


def allreduce_gradients(module, group=None):
    """Average the parameters' gradients over all ranks with a single flat all-reduce."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world == 1:
        return 1
    params = [p for p in module.parameters() if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
    return world


def train_step(decoder, llr, ground_truth, optimizer, group=None):
    """forward + backward + gradient all-reduce + optimizer step.  Returns the local loss (tensor)."""
    optimizer.zero_grad(set_to_none=False)
    _, loss = decoder(llr, None, None, None, None, ground_truth=ground_truth)
    loss.backward()
    allreduce_gradients(decoder, group)
    optimizer.step()
    return loss.detach()


def train_step_neural(decoder, llr, check_index_tensor, var_index_tensor, ground_truth, optimizer, group=None):
    """The reference's training iteration for `LDPCNeuralDecoder` (training/trainer.py:95-110: forward with the
    index tensors and the transmitted bits, `loss.mean().backward()`, optimizer step), data-parallel: codewords
    sharded by rank, one flat all-reduce of the gradient of `w_ch` (E,) and `w_res` (L,).  Returns the local
    mean of the per-frame max loss."""
    optimizer.zero_grad(set_to_none=False)
    _, loss = decoder(llr, check_index_tensor, var_index_tensor, ground_truth)
    batch_loss = loss.mean()
    batch_loss.backward()
    allreduce_gradients(decoder, group)
    optimizer.step()
    return batch_loss.detach()
