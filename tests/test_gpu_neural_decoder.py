"""LDPCNeuralDecoder (unrolled neural min-sum decoder) on the GPU against the
composition of the reference's own layer classes (golden) and against the oracle at Z=32.
Needs a B200 (-m gpu)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import oracle
import ldpc_b200  # noqa: F401
from ldpc_b200.models import LDPCNeuralDecoder
from ldpc_b200.utils import QCCode, create_LLR_mapping

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _decoder(g, fused, out_index=None):
    E = g["llr_e"].shape[1]
    dec = LDPCNeuralDecoder(num_nodes=E, num_iterations=int(g["iters"]), depth_L=int(g["depth_L"]),
                            output_index_tensor=out_index, fused=fused).to(DEV)
    with torch.no_grad():
        dec.residual_layer.w_ch.copy_(torch.from_numpy(g["w_ch"]))
        dec.residual_layer.w_res.copy_(torch.from_numpy(g["w_res"]))
    return dec


@pytest.mark.parametrize("fused", [True, False])
def test_forward_and_gradients_match_reference_layers(fused):
    g = load_golden("neural_decoder_z4")
    t = lambda k, dt=torch.float32: torch.from_numpy(g[k]).to(DEV).to(dt)
    dec = _decoder(g, fused)
    cidx, vidx = t("check", torch.int64), t("var", torch.int64)
    soft, ml = dec(t("llr_e"), cidx, vidx, t("gt_e"))
    np.testing.assert_allclose(soft.detach().cpu().numpy(), g["soft"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ml.detach().cpu().numpy(), g["max_loss"], rtol=1e-5)
    ml.mean().backward()                                             # trainer.py:105-107
    np.testing.assert_allclose(dec.residual_layer.w_ch.grad.cpu().numpy(), g["grad_wch"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(dec.residual_layer.w_res.grad.cpu().numpy(), g["grad_wres"], rtol=1e-4, atol=1e-5)
    # state_dict carries exactly the reference layer's parameter names
    assert set(dec.state_dict()) == {"residual_layer.w_ch", "residual_layer.w_res"}


def test_fused_variable_residual_kernel_is_bit_identical_to_the_composition():
    g = load_golden("neural_decoder_z4")
    t = lambda k, dt=torch.float32: torch.from_numpy(g[k]).to(DEV).to(dt)
    cidx, vidx = t("check", torch.int64), t("var", torch.int64)
    a = _decoder(g, True)._messages(t("llr_e"), cidx, vidx)
    b = _decoder(g, False)._messages(t("llr_e"), cidx, vidx)
    assert torch.equal(a, b)
    np.testing.assert_allclose(a.detach().cpu().numpy(), g["c2v"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("depth_L,iters", [(2, 4), (1, 3), (0, 2), (3, 6), (2, 1), (4, 7), (5, 7)])
def test_one_kernel_decoder_is_bit_identical_to_the_layer_chain(depth_L, iters):
    """ldpc_neural_decode (messages resident in shared memory) vs the per-layer kernels: same
    bits for soft outputs and per-frame max loss, for every queue depth / iteration count,
    on a ragged batch; and both equal the oracle."""
    g = load_golden("neural_decoder_z4")
    cidx = torch.from_numpy(g["check"]).to(DEV).long()
    vidx = torch.from_numpy(g["var"]).to(DEV).long()
    E = g["llr_e"].shape[1]
    rng = np.random.default_rng(depth_L * 10 + iters)
    B = 1201                                                         # > 4 * 2 * 148: the 4-row tiles, ragged tail
    llr_e = (rng.normal(size=(B, E)) * 0.6 + 0.3).astype(np.float32)
    llr_e[5, ::3] = 0.0
    gt_e = (rng.random((B, E)) > 0.1).astype(np.float32)
    w_ch = (rng.random(E) * 0.5 + 0.75).astype(np.float32)
    w_res = np.array([0.25, -0.125, 0.0625, 0.03125, -0.015625][:depth_L], np.float32)
    decs = []
    for fused in (True, False):
        d = LDPCNeuralDecoder(E, iters, depth_L, fused=fused).to(DEV)
        with torch.no_grad():
            d.residual_layer.w_ch.copy_(torch.from_numpy(w_ch))
            d.residual_layer.w_res.copy_(torch.from_numpy(w_res))
        decs.append(d)
    x, y = torch.from_numpy(llr_e).to(DEV), torch.from_numpy(gt_e).to(DEV)
    with torch.no_grad():
        for nb in (B, 300, 7):                                       # 4-, 2- and 1-row tiles
            s1, m1 = decs[0](x[:nb], cidx, vidx, y[:nb])             # one kernel
            s2, m2 = decs[1](x[:nb], cidx, vidx, y[:nb])             # four layers per iteration
            assert torch.equal(s1, s2) and torch.equal(m1, m2)
        s3, none = decs[0](x, cidx, vidx)
        assert none is None and torch.equal(s3, s1 if nb == B else decs[1](x, cidx, vidx)[0])
    ref = oracle.neural_minsum_forward(llr_e[:64], g["check"].astype(np.int64), g["var"].astype(np.int64),
                                       w_ch, w_res, iters, gt_e[:64])
    with torch.no_grad():
        s64, m64 = decs[0](x[:64], cidx, vidx, y[:64])
    np.testing.assert_allclose(s64.cpu().numpy(), ref["soft"], rtol=1e-4, atol=1e-6)
    # -log(1 - s) with s within a few ulp of 1 turns one ulp of the sigmoid (fp32 expf here, fp64 in the oracle)
    # into 0.3 of loss: compare the max loss where it is not saturated
    ok = ref["max_loss"] < 10.0
    np.testing.assert_allclose(m64.cpu().numpy()[ok], ref["max_loss"][ok], rtol=1e-3, atol=1e-6)


def test_variable_space_io_and_decode():
    g = load_golden("neural_decoder_z4")
    t = lambda k, dt=torch.float32: torch.from_numpy(g[k]).to(DEV).to(dt)
    cidx, vidx = t("check", torch.int64), t("var", torch.int64)
    dec = _decoder(g, True, out_index=torch.from_numpy(g["out_index"]))
    soft_v, none = dec(t("llr"), cidx, vidx)                         # (B, N) in -> (B, N) out
    assert none is None and soft_v.shape == g["llr"].shape
    first = np.full(g["llr"].shape[1], -1)
    for e, v in reversed(list(enumerate(g["out_index"][0]))):
        first[v] = e
    np.testing.assert_allclose(soft_v.detach().cpu().numpy(), g["soft"][:, first], rtol=1e-5, atol=1e-6)
    hard = dec.decode(t("llr"), cidx, vidx)
    assert torch.equal(hard, (soft_v > 0.5).float())
    with pytest.raises(ValueError):
        dec(t("llr")[:, :-1], cidx, vidx)
    with pytest.raises(RuntimeError):
        dec(torch.from_numpy(g["llr"]), cidx, vidx)                  # CPU tensor: no fallback


def test_z32_against_oracle_ragged_batch():
    code = QCCode.nr_2_0(32)
    _, cidx, vidx, oidx = create_LLR_mapping(code.dense().T)
    rng = np.random.default_rng(5)
    B, iters, L = 11, 5, 2
    llr = (rng.normal(size=(B, code.N)) * 0.5 + 0.4).astype(np.float32)
    llr[2, ::9] = 0.0
    llr_e = llr[:, oidx.numpy()[0]]
    w_ch = (rng.random(code.E) * 0.5 + 0.75).astype(np.float32)
    w_res = np.array([0.2, -0.1], np.float32)
    gt_e = np.ones_like(llr_e)
    dec = LDPCNeuralDecoder(code.E, iters, L).to(DEV)
    with torch.no_grad():
        dec.residual_layer.w_ch.copy_(torch.from_numpy(w_ch))
        dec.residual_layer.w_res.copy_(torch.from_numpy(w_res))
    soft, ml = dec(torch.from_numpy(llr_e).to(DEV), cidx.to(DEV), vidx.to(DEV), torch.from_numpy(gt_e).to(DEV))
    ref = oracle.neural_minsum_forward(llr_e, cidx.numpy(), vidx.numpy(), w_ch, w_res, iters, gt_e)
    np.testing.assert_allclose(soft.detach().cpu().numpy(), ref["soft"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ml.detach().cpu().numpy(), ref["max_loss"], rtol=1e-4, atol=1e-6)
    with torch.no_grad():                                            # one-kernel path, 2-row tiles at E = 6304
        big = torch.from_numpy(np.tile(llr_e, (30, 1))).to(DEV)
        sk, mk = dec(big, cidx.to(DEV), vidx.to(DEV), torch.ones_like(big))
        s1, m1 = dec(torch.from_numpy(llr_e).to(DEV), cidx.to(DEV), vidx.to(DEV), torch.from_numpy(gt_e).to(DEV))
    assert torch.equal(sk[:B], soft.detach()) and torch.equal(mk[-B:], ml.detach())
    assert torch.equal(s1, soft.detach()) and torch.equal(m1, ml.detach())
    plain = LDPCNeuralDecoder(code.E, iters, L, fused=False).to(DEV)
    plain.load_state_dict(dec.state_dict())
    args = (torch.from_numpy(llr_e).to(DEV), cidx.to(DEV), vidx.to(DEV))
    assert torch.equal(dec._messages(*args), plain._messages(*args))


def test_c_abi_identity_order_equals_sorted_order():
    """ldpc_neural_decode with cperm = vperm = NULL (columns in edge order) gives the same bits as the
    degree-sorted columns the Python class passes; padding may sit anywhere in the caller's rows."""
    from ldpc_b200 import _native
    from ldpc_b200.models.layers import packed_index
    g = load_golden("neural_decoder_z4")
    cidx = torch.from_numpy(g["check"]).to(DEV).long()
    vidx = torch.from_numpy(g["var"]).to(DEV).long()
    # move some padding to the FRONT of the rows: the kernel must not rely on compacted tables
    cidx = torch.flip(cidx, dims=[1]).contiguous()
    E = g["llr_e"].shape[1]
    B, iters, L = 77, 4, 2
    rng = np.random.default_rng(3)
    llr_e = torch.from_numpy((rng.normal(size=(B, E)) * 0.6 + 0.3).astype(np.float32)).to(DEV)
    w_ch = torch.from_numpy((rng.random(E) * 0.5 + 0.75).astype(np.float32)).to(DEV)
    w_res = torch.tensor([0.25, -0.125], device=DEV)
    pc, pv = packed_index(cidx), packed_index(vidx)
    outs = []
    for (ct, cp, cc), (vt, vp, vc) in ((pc.compacted(), pv.compacted()), (pc.sorted(), pv.sorted())):
        soft = torch.empty_like(llr_e)
        _native.check(_native.lib().ldpc_neural_decode(
            _native.ptr(llr_e), _native.ptr(ct), ct.shape[0], _native.ptr(cc), _native.ptr(cp), _native.ptr(vt),
            vt.shape[0], _native.ptr(vc), _native.ptr(vp), _native.ptr(w_ch), _native.ptr(w_res), L, iters, B, E, None,
            _native.ptr(soft), None, _native.stream_ptr(llr_e.device)))
        outs.append(soft)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    ref = oracle.neural_minsum_forward(llr_e.cpu().numpy(), cidx.cpu().numpy(), vidx.cpu().numpy(), w_ch.cpu().numpy(),
                                       w_res.cpu().numpy(), iters)
    np.testing.assert_allclose(outs[0].cpu().numpy(), ref["soft"], rtol=1e-4, atol=1e-6)


def test_training_steps_reduce_the_loss():
    """trainer.py:95-110 loop shape through train_step_neural: SGD on w_ch / w_res lowers the mean max-loss of a fixed batch."""
    from ldpc_b200.training import train_step_neural
    g = load_golden("neural_decoder_z4")
    t = lambda k, dt=torch.float32: torch.from_numpy(g[k]).to(DEV).to(dt)
    dec = _decoder(g, True)
    cidx, vidx = t("check", torch.int64), t("var", torch.int64)
    opt = torch.optim.SGD(dec.parameters(), lr=1e-2, momentum=0.9, weight_decay=1e-4)     # trainer.py:70
    losses = [float(train_step_neural(dec, t("llr_e"), cidx, vidx, t("gt_e"), opt)) for _ in range(8)]
    np.testing.assert_allclose(losses[0], g["max_loss"].mean(), rtol=1e-5)
    assert losses[-1] < losses[0]
    with torch.no_grad():                                            # the trained weights also run through the one-kernel path
        soft, ml = dec(t("llr_e"), cidx, vidx, t("gt_e"))
    assert torch.isfinite(soft).all() and float(ml.mean()) <= losses[0]


def test_unit_channel_weights_without_residuals_equal_classic_min_sum():
    """With w_ch = 1 and w_res = 0 the unrolled network IS flooding min-sum with scaling factor 1
    (x_{l+1} = llr + sum of the other check messages; output = llr + sum of all), so its decisions must equal
    MinSumScaledDecoder(scaling_factor=1.0) on the same LLRs: the edge-space path (generic index tables) checked
    against the QC-structured kernel family on the headline code."""
    from ldpc_b200.models import MinSumScaledDecoder
    code = QCCode.nr_2_0(32)
    _, cidx, vidx, oidx = create_LLR_mapping(code.dense().T)
    iters, B = 6, 256
    llr = torch.from_numpy(oracle.awgn_llr(None, B, code.N, snr_db=-2.0, seed=77)).to(DEV)
    dec = LDPCNeuralDecoder(code.E, iters, 2, output_index_tensor=oidx).to(DEV)
    with torch.no_grad():
        dec.residual_layer.w_res.zero_()
        soft, _ = dec(llr, cidx.to(DEV), vidx.to(DEV))               # sigmoid(posterior LLR), (B, N)
    classic = MinSumScaledDecoder(code, max_iterations=iters, scaling_factor=1.0, early_stopping=False, path="exact")
    beliefs, bits = classic.forward(llr)
    mine = (soft < 0.5).float()                                      # posterior < 0  <=>  bit 1
    differ = (mine != bits.float())
    # the two paths add the messages in different orders: decisions may differ only where the belief is ~0
    assert int(differ.sum()) <= 8 and (beliefs[differ].abs() < 1e-3).all()
    sure = beliefs.abs() < 10
    np.testing.assert_allclose(torch.logit(soft[sure].double()).cpu().numpy(), beliefs[sure].double().cpu().numpy(),
                               rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("depth_L,iters", [(2, 5), (1, 3), (0, 2), (2, 1), (2, 8)])
def test_qc_structured_kernel_is_bit_identical_to_the_table_driven_kernels(depth_L, iters):
    """csrc/neural_qc.cuh (neighbours implied by the base graph, state in Tensor Memory, no index tables) against the
    table-driven one-kernel decoder and the per-layer chain on BG2 Z=32: same bits for every soft output and per-frame
    max loss; per-edge LLRs that differ between the edges of one variable, zeros, a ragged batch."""
    code = QCCode.nr_2_0(32)
    _, cidx, vidx, oidx = create_LLR_mapping(code.dense().T)
    cidx, vidx = cidx.to(DEV), vidx.to(DEV)
    rng = np.random.default_rng(100 * depth_L + iters)
    B = 4 * 148 + 7                                                  # more codewords than warps in the grid, ragged tail
    llr_e = (rng.normal(size=(B, code.E)) * 0.8 + 0.4).astype(np.float32)      # edge space: independent value per edge
    llr_e[3, ::7] = 0.0
    llr_e[4, 5:900:3] = -1e-10
    llr_e[5] = np.abs(llr_e[5]) * 30.0
    w_ch = (rng.random(code.E) * 0.5 + 0.75).astype(np.float32)
    w_res = np.array([0.2, -0.1, 0.05][:max(depth_L, 1)], np.float32)
    gt_e = (rng.random((B, code.E)) < 0.5).astype(np.float32)

    def make(qc, fused=True):
        d = LDPCNeuralDecoder(code.E, iters, depth_L, fused=fused, qc=qc).to(DEV)
        with torch.no_grad():
            d.residual_layer.w_ch.copy_(torch.from_numpy(w_ch))
            d.residual_layer.w_res.copy_(torch.from_numpy(w_res[:depth_L]))
        return d
    x, y = torch.from_numpy(llr_e).to(DEV), torch.from_numpy(gt_e).to(DEV)
    with torch.no_grad():
        launches0 = ldpc_b200._native.lib().ldpc_launch_count()
        s_qc, m_qc = make(True)(x, cidx, vidx, y)
        assert ldpc_b200._native.lib().ldpc_launch_count() == launches0 + 1          # ONE kernel, no table packing
        s_tb, m_tb = make(False)(x, cidx, vidx, y)
        s_n, _ = make(True)(x, cidx, vidx)                                            # without ground truth
    assert torch.equal(s_qc, s_tb) and torch.equal(m_qc, m_tb) and torch.equal(s_n, s_qc)
    with torch.no_grad():                                            # soft targets: the two-logarithm form of the loss
        ysoft = torch.from_numpy(rng.random((64, code.E)).astype(np.float32)).to(DEV)
        assert torch.equal(make(True)(x[:64], cidx, vidx, ysoft)[1], make(False)(x[:64], cidx, vidx, ysoft)[1])
    chain = make(False, fused=False)
    s_ch, m_ch = chain(x[:64], cidx, vidx, y[:64])                                  # literal four-layer composition
    assert torch.equal(s_ch.detach(), s_qc[:64]) and torch.equal(m_ch.detach(), m_qc[:64])
    # tables that are NOT the code's (one entry changed) must not take the QC path
    other = cidx.clone()
    other[0, 0] = other[0, 1]
    with torch.no_grad():
        s_o, _ = make(True)(x[:8], other, vidx)
        s_t, _ = make(False)(x[:8], other, vidx)
    assert torch.equal(s_o, s_t)


@pytest.mark.parametrize("depth_L,iters", [(2, 5), (1, 4), (0, 3), (2, 2)])
def test_qc_structured_training_step_matches_the_per_layer_autograd(depth_L, iters):
    """One forward + one backward kernel on the QC structure (the trainer's `loss.mean().backward()`, trainer.py:105-107)
    against the per-layer kernels under torch.autograd (themselves pinned to the reference layers' autograd on the Z=4
    golden): same max_loss bits, gradients of w_ch and w_res within 1e-4 of each tensor's scale."""
    code = QCCode.nr_2_0(32)
    _, cidx, vidx, _ = create_LLR_mapping(code.dense().T)
    cidx, vidx = cidx.to(DEV), vidx.to(DEV)
    rng = np.random.default_rng(7 + depth_L)
    B = 4 * 148 + 5
    llr_e = (rng.normal(size=(B, code.E)) * 0.3 + 0.2).astype(np.float32) * 0.5      # keeps sigmoid / BCE out of saturation
    w_ch = (rng.random(code.E) * 0.5 + 0.75).astype(np.float32)
    w_res = np.array([0.2, -0.1][:depth_L], np.float32)
    gt_e = (rng.random((B, code.E)) < 0.7).astype(np.float32)
    x, y = torch.from_numpy(llr_e).to(DEV), torch.from_numpy(gt_e).to(DEV)

    def run(qc):
        d = LDPCNeuralDecoder(code.E, iters, depth_L, qc=qc).to(DEV)
        with torch.no_grad():
            d.residual_layer.w_ch.copy_(torch.from_numpy(w_ch))
            d.residual_layer.w_res.copy_(torch.from_numpy(w_res))
        n0 = ldpc_b200._native.lib().ldpc_launch_count()
        soft, ml = d(x, cidx, vidx, y)
        (ml * torch.linspace(0.5, 1.5, B, device=DEV)).mean().backward()            # non-uniform upstream gradient
        return soft.detach(), ml.detach(), d.residual_layer.w_ch.grad, d.residual_layer.w_res.grad, ldpc_b200._native.lib().ldpc_launch_count() - n0
    s_q, m_q, gw_q, gr_q, n_q = run(True)
    s_t, m_t, gw_t, gr_t, n_t = run(False)
    assert n_q == 2 and n_t > 2                                       # one forward + one backward kernel
    assert torch.equal(s_q, s_t) and torch.equal(m_q, m_t)
    assert float(gw_t.abs().max()) > 0
    assert float((gw_q - gw_t).abs().max()) <= 1e-4 * float(gw_t.abs().max())
    if depth_L:
        assert float((gr_q - gr_t).abs().max()) <= 1e-4 * max(float(gr_t.abs().max()), 1e-12), (gr_q, gr_t)


@pytest.mark.parametrize("Z", [4, 32])
def test_tied_decoder_equals_the_untied_one_with_expanded_weights(Z):
    """TiedNeuralLDPCDecoder(base_graph, Z, iterations, depth_L) (main.py:73-79): one channel weight per base-graph edge.
    Outputs equal LDPCNeuralDecoder with those weights copied to the Z lifted edges bit for bit; the gradient of a tied
    weight is the sum of the gradients of its copies."""
    from ldpc_b200.models import TiedNeuralLDPCDecoder
    code = QCCode.nr_2_0(Z)
    _, cidx, vidx, oidx = create_LLR_mapping(code.dense().T)
    cidx, vidx = cidx.to(DEV), vidx.to(DEV)
    rng = np.random.default_rng(Z)
    B, iters, L = 37, 4, 2
    tied = TiedNeuralLDPCDecoder(code.base_matrix(), Z, iters, L).to(DEV)
    plain = LDPCNeuralDecoder(code.E, iters, L, output_index_tensor=oidx).to(DEV)
    wt = torch.from_numpy((rng.random(code.base_edges) * 0.5 + 0.75).astype(np.float32)).to(DEV)
    wr = torch.tensor([0.2, -0.1], device=DEV)
    with torch.no_grad():
        tied.w_ch_tied.copy_(wt)
        tied.residual_layer.w_res.copy_(wr)
        plain.residual_layer.w_ch.copy_(wt[tied.edge_cell])
        plain.residual_layer.w_res.copy_(wr)
    assert {n for n, _ in tied.named_parameters()} == {"w_ch_tied", "residual_layer.w_res"}
    llr = torch.from_numpy((rng.normal(size=(B, code.N)) * 0.3 + 0.1).astype(np.float32)).to(DEV)     # variable-space input
    gt = torch.from_numpy((rng.random((B, code.N)) < 0.6).astype(np.float32)).to(DEV)
    s_t, m_t = tied(llr, ground_truth=gt)                       # tables owned by the decoder
    s_p, m_p = plain(llr, cidx, vidx, gt)
    assert s_t.shape == (B, code.N) and torch.equal(s_t.detach(), s_p.detach()) and torch.equal(m_t.detach(), m_p.detach())
    m_t.mean().backward()
    m_p.mean().backward()
    want = torch.zeros(code.base_edges, device=DEV).index_add_(0, tied.edge_cell, plain.residual_layer.w_ch.grad)
    assert float(want.abs().max()) > 0
    assert float((tied.w_ch_tied.grad - want).abs().max()) <= 1e-5 * float(want.abs().max())
    assert torch.allclose(tied.residual_layer.w_res.grad, plain.residual_layer.w_res.grad, rtol=1e-5, atol=1e-8)
    with torch.no_grad():
        assert torch.equal(tied.decode(llr), plain.decode(llr, cidx, vidx))
        assert torch.equal(tied(llr, cidx, vidx)[0], s_p.detach())           # reference-style call with explicit tables


def test_qc_structured_kernels_against_the_reference_layers_at_z32():
    """tests/golden/neural_decoder_z32.npz: the reference's OWN CheckLayer / VariableLayer / ResidualLayer / OutputLayer objects in
    the decoder's composition on BG2 Z=32 (5 iterations, L = 2; oracle/make_golden.py:neural_decoder), forward and autograd of
    loss.mean().  The QC-structured forward and backward kernels against it: soft outputs and max loss to 1e-5, parameter
    gradients to 1e-4 -- the same criteria the table-driven path meets on the Z=4 fixture."""
    g = load_golden("neural_decoder_z32")
    code = QCCode.nr_2_0(32)
    _, cidx, vidx, _ = create_LLR_mapping(code.dense().T)
    cidx, vidx = cidx.to(DEV), vidx.to(DEV)
    dec = LDPCNeuralDecoder(code.E, int(g["iters"]), int(g["depth_L"])).to(DEV)
    with torch.no_grad():
        dec.residual_layer.w_ch.copy_(torch.from_numpy(g["w_ch"]))
        dec.residual_layer.w_res.copy_(torch.from_numpy(g["w_res"]))
    x = torch.from_numpy(g["llr_e"]).to(DEV)
    y = torch.from_numpy(np.unpackbits(g["gt_e"], axis=1)[:, :code.E].astype(np.float32)).to(DEV)
    n0 = ldpc_b200._native.lib().ldpc_launch_count()
    soft, ml = dec(x, cidx, vidx, y)
    ml.mean().backward()                                             # trainer.py:105-107
    assert ldpc_b200._native.lib().ldpc_launch_count() - n0 == 2      # QC forward + QC backward, nothing else
    np.testing.assert_allclose(soft.detach().cpu().numpy(), g["soft"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ml.detach().cpu().numpy(), g["max_loss"], rtol=1e-5)
    gw, gr = dec.residual_layer.w_ch.grad.cpu().numpy(), dec.residual_layer.w_res.grad.cpu().numpy()
    assert np.count_nonzero(g["grad_wch"]) > 0
    np.testing.assert_allclose(gw, g["grad_wch"], rtol=1e-4, atol=1e-4 * np.abs(g["grad_wch"]).max())
    np.testing.assert_allclose(gr, g["grad_wres"], rtol=1e-4, atol=1e-4 * np.abs(g["grad_wres"]).max())
    with torch.no_grad():                                            # inference launch of the same kernel
        s2, m2 = dec(x, cidx, vidx, y)
    assert torch.equal(s2, soft.detach()) and torch.equal(m2, ml.detach())


@pytest.mark.parametrize("depth_L,iters", [(2, 5), (1, 3), (0, 2)])
def test_per_variable_qc_path_is_bit_identical_to_the_expanded_edge_space_path(depth_L, iters):
    """The trainer's call shape (trainer.py:95-110,180-187): (B, N) LLRs and targets.  ldpc_neural_decode_qc_var /
    ldpc_neural_backward_qc_var expand a variable's value to its edges on chip; against the edge-space kernels fed with
    llr[:, edge_to_var] / gt[:, edge_to_var]: soft outputs (at each variable's first edge) and max_loss EQUAL, in inference
    and in training, gradients within rounding of the accumulation order (atomics), and exactly two kernel launches per
    training step with no index_select expansion."""
    code = QCCode.nr_2_0(32)
    _, cidx, vidx, oidx = create_LLR_mapping(code.dense().T)
    cidx, vidx = cidx.to(DEV), vidx.to(DEV)
    etv = torch.as_tensor(oidx).reshape(-1).to(torch.int64)
    rng = np.random.default_rng(31 + depth_L)
    B = 4 * 148 + 3
    llr = torch.from_numpy(((rng.normal(size=(B, code.N)) * 0.3 + 0.2) * 0.5).astype(np.float32)).to(DEV)
    gt = torch.from_numpy((rng.random((B, code.N)) < 0.7).astype(np.float32)).to(DEV)
    w_ch = torch.from_numpy((rng.random(code.E) * 0.5 + 0.75).astype(np.float32))
    w_res = torch.tensor([0.2, -0.1][:depth_L])

    def make(with_map):
        d = LDPCNeuralDecoder(code.E, iters, depth_L, output_index_tensor=oidx if with_map else None).to(DEV)
        with torch.no_grad():
            d.residual_layer.w_ch.copy_(w_ch)
            d.residual_layer.w_res.copy_(w_res)
        return d
    dv, de = make(True), make(False)
    llr_e, gt_e = llr[:, etv.to(DEV)].contiguous(), gt[:, etv.to(DEV)].contiguous()
    first = dv.var_first_edge.to(DEV)
    lib = ldpc_b200._native.lib()
    # inference, without and with targets
    with torch.no_grad():
        n0 = lib.ldpc_launch_count()
        s_v, none = dv(llr, cidx, vidx)
        assert lib.ldpc_launch_count() - n0 == 1 and none is None and s_v.shape == llr.shape
        s_e, _ = de(llr_e, cidx, vidx)
        assert torch.equal(s_v, s_e[:, first])
        s_v, m_v = dv(llr, cidx, vidx, gt)
        s_e, m_e = de(llr_e, cidx, vidx, gt_e)
        assert torch.equal(s_v, s_e[:, first]) and torch.equal(m_v, m_e)
    # training step
    up = torch.linspace(0.5, 1.5, B, device=DEV)
    n0 = lib.ldpc_launch_count()
    s_v, m_v = dv(llr, cidx, vidx, gt)
    (m_v * up).mean().backward()
    assert lib.ldpc_launch_count() - n0 == 2
    s_e, m_e = de(llr_e, cidx, vidx, gt_e)
    (m_e * up).mean().backward()
    assert torch.equal(s_v.detach(), s_e.detach()[:, first]) and torch.equal(m_v.detach(), m_e.detach())
    gv, ge = dv.residual_layer.w_ch.grad, de.residual_layer.w_ch.grad
    assert float(ge.abs().max()) > 0 and float((gv - ge).abs().max()) <= 1e-5 * float(ge.abs().max())
    if depth_L:
        rv, re_ = dv.residual_layer.w_res.grad, de.residual_layer.w_res.grad
        assert float((rv - re_).abs().max()) <= 1e-5 * max(float(re_.abs().max()), 1e-12)
    # targets in edge space with LLRs per variable: not this path's shape -> the expansion path, same numbers
    with torch.no_grad():
        s_m, m_m = dv(llr, cidx, vidx, gt_e)
    assert torch.equal(s_m, s_v.detach()) and torch.equal(m_m, m_v.detach())


@pytest.mark.parametrize("Z,depth_L,iters,B", [(16, 2, 5, 2 * 4 * 148 + 3), (16, 1, 3, 1), (16, 0, 2, 2), (16, 2, 8, 45),
                                                (8, 2, 5, 4 * 4 * 148 + 5), (8, 1, 2, 3), (4, 2, 5, 8 * 4 * 148 + 9), (4, 0, 3, 7)])
def test_qc_structured_kernel_at_the_default_lifting_factor_16(Z, depth_L, iters, B):
    """Z = 16 is the reference's default --lifting_factor (main.py:38).  The QC-structured forward kernel then holds TWO codewords
    per warp (lane = 16 * sub + r; rotations inside the 16 lanes of a codeword).  Same bits as the table-driven kernel for every
    soft output and max loss -- edge-space and per-variable I/O, odd batch sizes (a warp with one live codeword), zeros.
    Z = 8 and Z = 4 (the shipped NR_2_0_4.txt): four and eight codewords per warp, same kernels."""
    code = QCCode.nr_2_0(Z)
    _, cidx, vidx, oidx = create_LLR_mapping(code.dense().T)
    cidx, vidx = cidx.to(DEV), vidx.to(DEV)
    etv = torch.as_tensor(oidx).reshape(-1).to(torch.int64).to(DEV)
    rng = np.random.default_rng(7 * depth_L + iters)
    llr_e = (rng.normal(size=(B, code.E)) * 0.8 + 0.4).astype(np.float32)
    llr_e[0, ::7] = 0.0
    llr_e[B // 2, 5:900:3] = -1e-10
    w_ch = (rng.random(code.E) * 0.5 + 0.75).astype(np.float32)
    w_res = np.array([0.2, -0.1][:depth_L], np.float32)
    gt_e = (rng.random((B, code.E)) < 0.5).astype(np.float32)

    def make(qc, out=None):
        d = LDPCNeuralDecoder(code.E, iters, depth_L, output_index_tensor=out, qc=qc).to(DEV)
        with torch.no_grad():
            d.residual_layer.w_ch.copy_(torch.from_numpy(w_ch))
            d.residual_layer.w_res.copy_(torch.from_numpy(w_res))
        return d
    x, y = torch.from_numpy(llr_e).to(DEV), torch.from_numpy(gt_e).to(DEV)
    lib = ldpc_b200._native.lib()
    with torch.no_grad():
        n0 = lib.ldpc_launch_count()
        s_qc, m_qc = make(True)(x, cidx, vidx, y)
        assert lib.ldpc_launch_count() == n0 + 1
        s_tb, m_tb = make(False)(x, cidx, vidx, y)
        assert torch.equal(s_qc, s_tb) and torch.equal(m_qc, m_tb)
        assert torch.equal(make(True)(x, cidx, vidx)[0], s_qc)
        # per-variable I/O
        llr_v = torch.from_numpy((rng.normal(size=(B, code.N)) * 0.8 + 0.4).astype(np.float32)).to(DEV)
        gt_v = torch.from_numpy((rng.random((B, code.N)) < 0.5).astype(np.float32)).to(DEV)
        dv = make(True, oidx)
        n0 = lib.ldpc_launch_count()
        s_v, m_v = dv(llr_v, cidx, vidx, gt_v)
        assert lib.ldpc_launch_count() == n0 + 1 and s_v.shape == llr_v.shape
        s_e, m_e = make(False)(llr_v[:, etv].contiguous(), cidx, vidx, gt_v[:, etv].contiguous())
        assert torch.equal(s_v, s_e[:, dv.var_first_edge.to(DEV)]) and torch.equal(m_v, m_e)
    # training at Z = 16: one forward + one backward kernel, against the per-layer kernels under autograd
    up = torch.linspace(0.5, 1.5, B, device=DEV)
    xs, ys = x * 0.25, y                                     # keeps sigmoid / BCE out of saturation
    dq, dt = make(True), make(False)
    n0 = lib.ldpc_launch_count()
    _, m_q = dq(xs, cidx, vidx, ys)
    (m_q * up).mean().backward()
    assert lib.ldpc_launch_count() == n0 + 2
    _, m_t = dt(xs, cidx, vidx, ys)
    (m_t * up).mean().backward()
    assert torch.equal(m_q.detach(), m_t.detach())
    gq, gt_ = dq.residual_layer.w_ch.grad, dt.residual_layer.w_ch.grad
    assert float(gt_.abs().max()) > 0 and float((gq - gt_).abs().max()) <= 1e-4 * float(gt_.abs().max())
    if depth_L:
        rq, rt = dq.residual_layer.w_res.grad, dt.residual_layer.w_res.grad
        assert float((rq - rt).abs().max()) <= 1e-4 * max(float(rt.abs().max()), 1e-12)
    # per-variable training step: same gradients as the edge-space step on the expanded arrays
    dvq = make(True, oidx)
    _, m_v = dvq(llr_v * 0.25, cidx, vidx, gt_v)
    (m_v * up).mean().backward()
    dte = make(False)
    _, m_e2 = dte((llr_v * 0.25)[:, etv].contiguous(), cidx, vidx, gt_v[:, etv].contiguous())
    (m_e2 * up).mean().backward()
    assert torch.equal(m_v.detach(), m_e2.detach())
    gv, ge = dvq.residual_layer.w_ch.grad, dte.residual_layer.w_ch.grad
    assert float((gv - ge).abs().max()) <= 1e-4 * max(float(ge.abs().max()), 1e-12)
