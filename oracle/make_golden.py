#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE ONLY.  This script imports the reference implementation from
/root/reference (read-only, present only in the build container, never on the GPU
box), feeds it seeded inputs and stores inputs + outputs as small .npz files.  The
fixtures pin the oracle (oracle/ldpc_oracle.c, oracle/oracle.py) and, through it, the
CUDA path.  Nothing under the product package imports this file.

Reference entry points exercised (paths relative to /root/reference/ldpc_neural_decoder):
  models/traditional_decoders.py:137-260  MinSumScaledDecoder.decode
  models/traditional_decoders.py:4-109    BeliefPropagationDecoder.decode
  models/layers.py:5-210                  CheckLayer / VariableLayer / ResidualLayer / OutputLayer
  models/message_gnn_decoder.py:155-582   MessageGNNDecoder / TannerToMessageGraph / factory
  utils/ldpc_utils.py:5-146               load_base_matrix / expand_base_matrix / create_LLR_mapping
  utils/channel.py:156-231                compute_ber_fer / AWGNChannel.transmit

The reference never returns its soft beliefs (`var_beliefs` is a local of decode());
we read that local at function return with sys.setprofile, i.e. the reference code is
executed unchanged and merely observed.

Usage:  python oracle/make_golden.py [--only NAME ...]
Run time: ~6 minutes on 8 cores (the reference's Python loops dominate).
"""
import argparse
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

REF_ROOT = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, REF_ROOT)

from ldpc_neural_decoder.models import (  # noqa: E402
    BeliefPropagationDecoder, MinSumScaledDecoder, CheckLayer, VariableLayer,
    ResidualLayer, OutputLayer, create_message_gnn_decoder)
from ldpc_neural_decoder.utils.ldpc_utils import (  # noqa: E402
    load_base_matrix, expand_base_matrix, create_LLR_mapping)
from ldpc_neural_decoder.utils.channel import (  # noqa: E402
    AWGNChannel, compute_ber_fer, qpsk_modulate, awgn_channel, qpsk_demodulate)

TABLES = {4: os.path.join(REF_ROOT, "5G LDPC CODES", "NR_2_0_4.txt"),
          32: os.path.join(REF_ROOT, "5G LDPC CODES", "NR_2_0_32.txt")}


def run_capturing_beliefs(dec, llr):
    """Call dec.decode(llr) unchanged and return (bits, iters, var_beliefs-at-return)."""
    captured = {}
    code = type(dec).decode.__code__

    def prof(frame, event, arg):
        if event == "return" and frame.f_code is code:
            captured["beliefs"] = frame.f_locals["var_beliefs"].clone()

    sys.setprofile(prof)
    try:
        bits, iters = dec.decode(llr)
    finally:
        sys.setprofile(None)
    return bits, iters, captured["beliefs"]


def share_indices(src, dst_cls, **kw):
    """Build a second classic decoder without paying _precompute_indices twice."""
    dst = dst_cls.__new__(dst_cls)
    dst.H = src.H
    dst.check_to_var = src.check_to_var
    dst.var_to_check = src.var_to_check
    for k, v in kw.items():
        setattr(dst, k, v)
    return dst


def pack(bits):
    return np.packbits(bits.numpy().astype(np.uint8), axis=1)


def classic(Z, B, iters, snr_db, seed, alpha, tag):
    base = load_base_matrix(TABLES[Z])
    H = expand_base_matrix(base, Z)
    t0 = time.time()
    ms = MinSumScaledDecoder(H, max_iterations=iters, scaling_factor=alpha, early_stopping=False)
    t_pre = time.time() - t0
    bp = share_indices(ms, BeliefPropagationDecoder, max_iterations=iters, early_stopping=False)
    torch.manual_seed(seed)
    llr = AWGNChannel().transmit(torch.zeros(B, H.shape[1]), snr_db)
    t0 = time.time()
    ms_bits, ms_it, ms_bel = run_capturing_beliefs(ms, llr)
    t_ms = time.time() - t0
    t0 = time.time()
    bp_bits, bp_it, bp_bel = run_capturing_beliefs(bp, llr)
    t_bp = time.time() - t0
    ber, fer = compute_ber_fer(torch.zeros_like(ms_bits), ms_bits)
    np.savez_compressed(
        os.path.join(OUT, f"classic_{tag}.npz"),
        Z=Z, iters=iters, alpha=alpha, snr_db=snr_db, seed=seed,
        llr=llr.numpy(), ms_bits=pack(ms_bits), ms_beliefs=ms_bel.numpy(), ms_iters=ms_it,
        bp_bits=pack(bp_bits), bp_beliefs=bp_bel.numpy(), bp_iters=bp_it,
        ms_ber=ber, ms_fer=fer,
        ref_seconds=np.array([t_pre, t_ms, t_bp]), torch_version=torch.__version__)
    print(f"classic_{tag}: pre {t_pre:.1f}s ms {t_ms:.1f}s bp {t_bp:.1f}s  ms FER {fer:.3f} "
          f"bp nonfinite {(~torch.isfinite(bp_bel)).float().mean():.3f}")


def nonfinite(Z=4, B=24, iters=6, seed=21):
    """Non-finite channel LLRs through the unmodified reference (min-sum and BP): hard-decision inputs (+-inf), single
    huge / infinite entries, a NaN entry.  Pins how inf - inf, sign(NaN) and `mag < min_mag` behave in the reference
    (traditional_decoders.py:207-244) -- the engine's exact kernel and the oracle must reproduce the NaN/inf pattern."""
    base = load_base_matrix(TABLES[Z])
    H = expand_base_matrix(base, Z)
    ms = MinSumScaledDecoder(H, max_iterations=iters, scaling_factor=0.75, early_stopping=False)
    bp = share_indices(ms, BeliefPropagationDecoder, max_iterations=iters, early_stopping=False)
    torch.manual_seed(seed)
    llr = AWGNChannel().transmit(torch.zeros(B, H.shape[1]), 1.0)
    hd = torch.where(llr >= 0, torch.tensor(float("inf")), torch.tensor(float("-inf")))
    llr[:8] = hd[:8]                                   # frames 0-7: pure hard-decision input
    llr[8:12, ::3] = hd[8:12, ::3]                     # frames 8-11: every third LLR saturated
    llr[12, 5] = float("inf")
    llr[13, 7] = float("-inf")
    llr[14, 11] = float("nan")
    llr[15, 3] = 1e30
    llr[16, :4] = 0.0
    llr[16, 9] = float("inf")
    ms_bits, ms_it, ms_bel = run_capturing_beliefs(ms, llr)
    bp_bits, bp_it, bp_bel = run_capturing_beliefs(bp, llr)
    np.savez_compressed(os.path.join(OUT, f"nonfinite_z{Z}_b{B}.npz"), Z=Z, iters=iters, alpha=0.75, llr=llr.numpy(),
                        ms_bits=pack(ms_bits), ms_beliefs=ms_bel.numpy(), bp_bits=pack(bp_bits), bp_beliefs=bp_bel.numpy())
    print(f"nonfinite: ms nan {torch.isnan(ms_bel).float().mean():.3f} inf {torch.isinf(ms_bel).float().mean():.3f}; "
          f"bp nan {torch.isnan(bp_bel).float().mean():.3f} inf {torch.isinf(bp_bel).float().mean():.3f}")


def early_stop(Z, B, iters, snr_db, seed, tag):
    """early_stopping=True: batch-global stop rule (traditional_decoders.py:255-258)."""
    base = load_base_matrix(TABLES[Z])
    H = expand_base_matrix(base, Z)
    ms = MinSumScaledDecoder(H, max_iterations=iters, scaling_factor=0.75, early_stopping=True)
    bp = share_indices(ms, BeliefPropagationDecoder, max_iterations=iters, early_stopping=True)
    torch.manual_seed(seed)
    llr = AWGNChannel().transmit(torch.zeros(B, H.shape[1]), snr_db)
    ms_bits, ms_it, ms_bel = run_capturing_beliefs(ms, llr)
    bp_bits, bp_it, bp_bel = run_capturing_beliefs(bp, llr)
    valid = ms._check_valid_codeword(ms_bits)
    np.savez_compressed(
        os.path.join(OUT, f"earlystop_{tag}.npz"), Z=Z, iters=iters, alpha=0.75, snr_db=snr_db,
        llr=llr.numpy(), ms_bits=pack(ms_bits), ms_beliefs=ms_bel.numpy(), ms_iters=ms_it,
        bp_bits=pack(bp_bits), bp_beliefs=bp_bel.numpy(), bp_iters=bp_it, ms_valid=valid.numpy())
    print(f"earlystop_{tag}: ms iters {ms_it} bp iters {bp_it} valid {valid.float().mean():.2f}")


def mapping_and_layers():
    out = {}
    # notebook cell 5/7 toy H: the only golden vector the reference itself prints
    Htoy = torch.tensor([[1, 1, 0, 0], [0, 1, 1, 1], [1, 0, 0, 1]], dtype=torch.float32)
    m, c, v, o = create_LLR_mapping(Htoy.T)
    out.update(toy_H=Htoy.numpy(), toy_map=m.numpy(), toy_check=c.numpy(), toy_var=v.numpy(), toy_out=o.numpy())
    base = load_base_matrix(TABLES[4])
    H = expand_base_matrix(base, 4)
    m, c, v, o = create_LLR_mapping(H.T)
    out.update(z4_H=np.packbits(H.numpy().astype(np.uint8), axis=1), z4_base=base.numpy(),
               z4_map=m.numpy().astype(np.int32), z4_check=c.numpy().astype(np.int32),
               z4_var=v.numpy().astype(np.int32), z4_out=o.numpy().astype(np.int32))
    E = c.shape[0]
    torch.manual_seed(7)
    B = 8
    x = torch.randn(B, E) * 3
    x[0, :16] = 0.0           # exact zeros exercise the "+1e-10 / 0 -> 1e10" rules (layers.py:52-57)
    x[1] = 0.0                # an all-zero row yields 1e10 magnitudes
    llr_e = torch.randn(B, E) * 2
    cl, vl = CheckLayer(), VariableLayer()
    c2v = cl(x, c)
    v2c = vl(llr_e, c2v, v)
    res = ResidualLayer(E, depth_L=2)
    with torch.no_grad():
        res.w_ch.copy_(torch.rand(E) + 0.5)
        res.w_res.copy_(torch.tensor([0.3, -0.2]))
    prev = [torch.randn(B, E), torch.randn(B, E), torch.randn(B, E)]   # third is ignored (i < depth_L)
    r = res(llr_e, c2v, prev)
    gt = (torch.rand(B, E) > 0.5).float()
    soft, max_loss = OutputLayer()(r * 0.1, llr_e, gt)
    # gradients of a scalar through check->variable->residual->output (autograd of the reference ops)
    xg = x.clone().requires_grad_(True)
    lg = llr_e.clone().requires_grad_(True)
    c2v_g = cl(xg, c)
    v2c_g = vl(lg, c2v_g, v)
    rg = res(lg, c2v_g, [v2c_g, prev[1]])
    sg, ml = OutputLayer()(rg * 0.1, lg, gt)
    ml.sum().backward()
    out.update(lay_x=x.numpy(), lay_llr=llr_e.numpy(), lay_c2v=c2v.detach().numpy(), lay_v2c=v2c.detach().numpy(),
               lay_wch=res.w_ch.detach().numpy(), lay_wres=res.w_res.detach().numpy(),
               lay_prev=torch.stack(prev).numpy(), lay_res=r.detach().numpy(), lay_gt=gt.numpy(),
               lay_soft=soft.detach().numpy(), lay_maxloss=max_loss.detach().numpy(),
               lay_grad_x=xg.grad.numpy(), lay_grad_llr=lg.grad.numpy(),
               lay_grad_wch=res.w_ch.grad.numpy(), lay_grad_wres=res.w_res.grad.numpy(),
               lay_g_maxloss=ml.detach().numpy())
    # compute_ber_fer known answers (channel.py:156-190)
    tx = torch.zeros(4, 10)
    rx = torch.zeros(4, 10)
    rx[1, 3] = 1
    rx[3, :4] = 1
    ber, fer = compute_ber_fer(tx, rx)
    out.update(berfer_tx=tx.numpy(), berfer_rx=rx.numpy(), berfer=np.array([ber, fer]))
    np.savez_compressed(os.path.join(OUT, "mapping_layers.npz"), **out)
    print("mapping_layers: toy check\n", c.shape, "ber/fer", ber, fer)


def gnn(Z, B, snr_db, tag, with_grad):
    base = load_base_matrix(TABLES[Z])
    H = expand_base_matrix(base, Z)
    torch.manual_seed(0)
    t0 = time.time()
    dec, conv = create_message_gnn_decoder(H, num_iterations=5, hidden_dim=64, base_graph=base, Z=Z)
    t_build = time.time() - t0
    types = conv.get_message_types(base, Z)
    m2v = torch.tensor([v for v, _ in conv.messages], dtype=torch.long)   # the intended 1-D mapping
    torch.manual_seed(11)
    llr = AWGNChannel().transmit(torch.zeros(B, H.shape[1]), snr_db)
    sink = io.StringIO()
    t0 = time.time()
    with contextlib.redirect_stdout(sink):
        with torch.no_grad():
            probs = dec(llr, m2v, types, conv.var_to_check_adjacency, conv.check_to_var_adjacency)
            hard = dec.decode(llr, m2v, types, conv.var_to_check_adjacency, conv.check_to_var_adjacency)
    t_fwd = time.time() - t0
    sd = {k: v.detach().numpy() for k, v in dec.state_dict().items()}
    extra = {}
    if with_grad:
        gt = torch.zeros(B, H.shape[1])
        with contextlib.redirect_stdout(sink):
            p2, loss = dec(llr, m2v, types, conv.var_to_check_adjacency, conv.check_to_var_adjacency, ground_truth=gt)
        loss.backward()
        extra["loss"] = loss.detach().numpy()
        for k, p in dec.named_parameters():
            extra["grad." + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
            extra["hasgrad." + k] = np.array(p.grad is not None)
    np.savez_compressed(
        os.path.join(OUT, f"gnn_{tag}.npz"), Z=Z, snr_db=snr_db, llr=llr.numpy(), probs=probs.numpy(),
        hard=pack(hard), types=types.numpy().astype(np.int32), m2v=m2v.numpy().astype(np.int32),
        msg_check=np.array([c for _, c in conv.messages], dtype=np.int32),
        ref_seconds=np.array([t_build, t_fwd]), **{"sd." + k: v for k, v in sd.items()}, **extra)
    print(f"gnn_{tag}: build {t_build:.1f}s fwd {t_fwd:.1f}s params {sum(v.size for v in sd.values())}")


def qpsk():
    """The reference's QPSK chain (utils/channel.py:4-154) on fixed inputs: random bits of even and odd length,
    the symbols it maps them to, a fixed received block and the LLRs it demodulates, plus one pass through its
    awgn_channel under a fixed torch seed (statistics only: the engine's noise source is Philox)."""
    g = torch.Generator().manual_seed(2024)
    out = {}
    for tag, n in (("even", 208), ("odd", 51)):
        bits = torch.randint(0, 2, (6, n), generator=g).float()
        sym = qpsk_modulate(bits)
        out[f"bits_{tag}"] = bits.numpy().astype(np.uint8)
        out[f"sym_{tag}"] = torch.view_as_real(sym).numpy()
        rx = sym + torch.complex(torch.randn(sym.shape, generator=g) * 0.4, torch.randn(sym.shape, generator=g) * 0.4)
        for snr in (-2.0, 1.5, 6.0):
            out[f"llr_{tag}_snr{snr}"] = qpsk_demodulate(rx, snr).numpy()
        out[f"rx_{tag}"] = torch.view_as_real(rx).numpy()
    one = qpsk_demodulate(qpsk_modulate(torch.tensor([0., 1., 1.])), 0.0)        # un-batched, odd length
    out["llr_unbatched"] = one.numpy()
    torch.manual_seed(7)
    zeros = torch.zeros(512, 208)
    llr = qpsk_demodulate(awgn_channel(qpsk_modulate(zeros), 1.0), 1.0)
    out["chain_snr1_mean_std"] = np.array([llr.mean().item(), llr.std().item()], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "qpsk.npz"), **out)


def neural_decoder(Z=4, B=6, iters=4, depth_L=2):
    """The composition LDPCNeuralDecoder stands for (models/decoder.py is missing from the
    reference; notebook cell 11 is its prototype), run with the REFERENCE's own layer
    classes: forward, per-frame max loss, and autograd gradients of loss.mean()
    (training/trainer.py:102-107)."""
    base = load_base_matrix(TABLES[Z])
    H = expand_base_matrix(base, Z)
    _, c, v, o = create_LLR_mapping(H.T)
    E = c.shape[0]
    torch.manual_seed(11)
    bits = torch.zeros(B, H.shape[1])
    # scaled-down LLRs at low SNR keep sigmoid/BCE out of saturation, so the per-frame max loss has a gradient
    llr = AWGNChannel().transmit(bits, snr_db=-3.0) * 0.125
    llr[0, :3] = 0.0                                       # exact zeros: the "ignored" rule of CheckLayer
    llr_e = llr[:, o[0]].clone()
    # sigmoid(LLR > 0) -> 1 is the reference's output convention (layers.py:198), so the all-zero codeword's
    # target in that convention is all ones; a few flipped targets exercise both BCE branches
    gt_e = (torch.rand(B, H.shape[1]) > 0.05).float()[:, o[0]]
    cl, vl, ol = CheckLayer(), VariableLayer(), OutputLayer()
    res = ResidualLayer(E, depth_L=depth_L)
    with torch.no_grad():
        res.w_ch.copy_(torch.rand(E) * 0.5 + 0.75)
        res.w_res.copy_(torch.tensor([0.25, -0.125, 0.0625][:depth_L]))
    queue, x, c2v, xs = [], llr_e, None, []
    for l in range(iters):
        c2v = cl(x, c)
        if l == iters - 1:
            break
        s = vl(torch.zeros_like(c2v), c2v, v)
        x = res(llr_e, s, queue[:depth_L])
        queue.insert(0, x)
        xs.append(x.detach().numpy())
    final = vl(c2v, c2v, v)
    soft, max_loss = ol(final, llr_e, gt_e)
    max_loss.mean().backward()
    if Z == 32:
        # the index tables are create_LLR_mapping's (the test rebuilds them); intermediate activations are dropped: the
        # fixture pins the QC-structured kernels (csrc/neural_qc_kernel.cuh) directly to the reference's layer classes
        np.savez_compressed(
            os.path.join(OUT, f"neural_decoder_z{Z}.npz"),
            Z=Z, iters=iters, depth_L=depth_L, llr_e=llr_e.numpy(), gt_e=np.packbits(gt_e.numpy().astype(np.uint8), axis=1),
            w_ch=res.w_ch.detach().numpy(), w_res=res.w_res.detach().numpy(), soft=soft.detach().numpy(),
            max_loss=max_loss.detach().numpy(), grad_wch=res.w_ch.grad.numpy(), grad_wres=res.w_res.grad.numpy())
    else:
        np.savez_compressed(
            os.path.join(OUT, f"neural_decoder_z{Z}.npz"),
            Z=Z, iters=iters, depth_L=depth_L, check=c.numpy().astype(np.int32), var=v.numpy().astype(np.int32),
            out_index=o.numpy().astype(np.int32), llr=llr.numpy(), llr_e=llr_e.numpy(), gt_e=gt_e.numpy(),
            w_ch=res.w_ch.detach().numpy(), w_res=res.w_res.detach().numpy(), x=np.stack(xs),
            c2v=c2v.detach().numpy(), final=final.detach().numpy(), soft=soft.detach().numpy(),
            max_loss=max_loss.detach().numpy(), grad_wch=res.w_ch.grad.numpy(), grad_wres=res.w_res.grad.numpy())
    print("neural_decoder: E", E, "max_loss", max_loss.detach().numpy(), "grad_wres", res.w_res.grad.numpy(),
          "nonzero grad_wch", int((res.w_ch.grad != 0).sum()))


JOBS = {
    "mapping_layers": mapping_and_layers,
    # BASELINE.json config 1 (plumbing): Z=4, B=1024, 5 iters, alpha 0.75, snr_db 2.0, seed 1234
    "classic_z4": lambda: classic(4, 1024, 5, 2.0, 1234, 0.75, "z4_b1024_it5"),
    # low SNR at Z=4: many sign flips, exact zeros unlikely but ties/saturation exercised
    "classic_z4_low": lambda: classic(4, 256, 10, -2.0, 4321, 0.8, "z4_b256_it10_a08"),
    # headline code, BG2 Z=32, 10 iterations
    "classic_z32": lambda: classic(32, 16, 10, -2.0, 1234, 0.75, "z32_b16_it10"),
    "classic_z32_hi": lambda: classic(32, 8, 10, 0.0, 99, 0.75, "z32_b8_it10_snr0"),
    "nonfinite_z4": nonfinite,
    "earlystop_z4": lambda: early_stop(4, 8, 20, 1.0, 5, "z4_b8"),
    "gnn_z4": lambda: gnn(4, 4, 1.0, "z4_b4", True),
    "gnn_z32": lambda: gnn(32, 2, -2.0, "z32_b2", True),
    "qpsk": qpsk,
    "neural_decoder": neural_decoder,
    "neural_decoder_z32": lambda: neural_decoder(Z=32, B=5, iters=5, depth_L=2),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    for name, fn in JOBS.items():
        if args.only and name not in args.only:
            continue
        t0 = time.time()
        fn()
        print(f"[{name}] done in {time.time() - t0:.1f}s", flush=True)
