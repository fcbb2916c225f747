"""Parity of the CUDA decoders (through the drop-in classes -> ctypes -> C ABI) with the
oracle and with the golden vectors of the unmodified reference.  Needs a B200 (-m gpu)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, unpack
from test_oracle_golden import CLASSIC, fast_order_tolerance_ok
from oracle import oracle
import ldpc_b200
from ldpc_b200 import _native
from ldpc_b200.models import MinSumScaledDecoder, BeliefPropagationDecoder
from ldpc_b200.utils import QCCode, expand_base_matrix

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def run(dec, llr):
    soft, hard = dec.forward(torch.from_numpy(np.ascontiguousarray(llr)).to(dev()))
    torch.cuda.synchronize()
    return soft.cpu().numpy(), hard.cpu().numpy().astype(np.uint8)


# ---- golden vectors of the reference -------------------------------------------------------
@pytest.mark.parametrize("name", CLASSIC)
def test_minsum_exact_path_bit_identical_to_reference(name):
    g = load_golden(name)
    Z = int(g["Z"])
    code = QCCode.nr_2_0(Z)
    dec = MinSumScaledDecoder(code, int(g["iters"]), float(g["alpha"]), early_stopping=False, path="exact")
    soft, hard = run(dec, g["llr"])
    assert np.array_equal(hard, unpack(g["ms_bits"], code.N))
    assert np.array_equal(soft, g["ms_beliefs"])          # the reference's `var_beliefs`, bit for bit


@pytest.mark.parametrize("name", CLASSIC)
def test_minsum_fast_path_against_reference(name):
    g = load_golden(name)
    Z = int(g["Z"])
    code = QCCode.nr_2_0(Z)
    it, alpha = int(g["iters"]), float(g["alpha"])
    assert code.has_fast_path(dev())
    dec = MinSumScaledDecoder(code, it, alpha, early_stopping=False, path="fast")
    soft, hard = run(dec, g["llr"])
    assert np.array_equal(hard, unpack(g["ms_bits"], code.N))
    r = oracle.decode(code.shifts, Z, g["llr"], it, "minsum", alpha, want_mask=True)
    conv = ((r["valid_mask"][:, (it - 1) >> 6] >> np.uint64((it - 1) & 63)) & np.uint64(1)).astype(bool)
    assert fast_order_tolerance_ok(soft, g["ms_beliefs"], conv)
    # and bit-identical to the oracle restated in the kernel's own operation order
    o = oracle.decode(code.shifts, Z, g["llr"], it, "minsum", alpha, order="fast")
    assert np.array_equal(soft, o["beliefs"])


@pytest.mark.parametrize("name", CLASSIC)
def test_bp_against_reference_and_oracle(name):
    g = load_golden(name)
    Z = int(g["Z"])
    code = QCCode.nr_2_0(Z)
    dec = BeliefPropagationDecoder(code, int(g["iters"]), early_stopping=False, path="exact")
    soft, hard = run(dec, g["llr"])
    ref = g["bp_beliefs"]
    assert np.array_equal(hard, unpack(g["bp_bits"], code.N))
    assert np.array_equal(np.isnan(ref), np.isnan(soft))
    assert np.array_equal(np.isposinf(ref), np.isposinf(soft))
    assert np.array_equal(np.isneginf(ref), np.isneginf(soft))
    fin = np.isfinite(ref)
    assert np.all(np.abs(soft[fin] - ref[fin]) <= 2e-4 * np.maximum(np.abs(ref[fin]), 1.0))
    # same correctly rounded tanh/atanh as the oracle: every bit equal (NaN == NaN)
    o = oracle.decode(code.shifts, Z, g["llr"], int(g["iters"]), "bp")
    assert np.array_equal(soft, o["beliefs"], equal_nan=True)


def test_reference_early_stopping_rule():
    g = load_golden("earlystop_z4_b8")
    code = QCCode.nr_2_0(4)
    llr = torch.from_numpy(g["llr"]).to(dev())
    for cls, key, kw in ((MinSumScaledDecoder, "ms", dict(scaling_factor=0.75)), (BeliefPropagationDecoder, "bp", {})):
        dec = cls(code, max_iterations=int(g["iters"]), early_stopping=True, path="exact", **kw)
        bits, iters = dec.decode(llr)
        assert iters == int(g[key + "_iters"])
        assert bits.dtype == torch.float32 and bits.shape == llr.shape
        assert np.array_equal(bits.cpu().numpy().astype(np.uint8), unpack(g[key + "_bits"], code.N))
        assert bool(dec._check_valid_codeword(bits).all())
        soft = dec.forward(llr)[0].cpu().numpy()
        if key == "ms":
            assert np.array_equal(soft, g[key + "_beliefs"])
        else:   # BP: same inf/NaN pattern, finite values within the tanh/atanh 1-ulp band
            ref = g[key + "_beliefs"]
            fin = np.isfinite(ref)
            assert np.array_equal(fin, np.isfinite(soft)) and np.array_equal(np.isnan(ref), np.isnan(soft))
            assert np.all(np.abs(soft[fin] - ref[fin]) <= 2e-4 * np.maximum(np.abs(ref[fin]), 1.0))
        auto = cls(code, max_iterations=int(g["iters"]), early_stopping=True, **kw)     # fast kernel for pass 2
        bits2, iters2 = auto.decode(llr)
        assert iters2 == iters and torch.equal(bits2, bits)
        ref = g[key + "_beliefs"]
        fin = np.isfinite(ref)
        np.testing.assert_allclose(auto.forward(llr)[0].cpu().numpy()[fin], ref[fin], rtol=2e-4, atol=2e-4)


# ---- seeded inputs against the oracle ------------------------------------------------------
@pytest.mark.parametrize("Z,B,iters,snr_db,alpha", [(32, 97, 10, -2.0, 0.75), (32, 33, 3, -3.0, 0.8), (4, 1001, 5, 2.0, 0.75),
                                                     (4, 5, 7, -4.0, 0.9), (32, 1, 1, 0.0, 0.75), (16, 40, 6, -1.0, 0.75),
                                                     (16, 515, 10, -2.0, 0.75), (8, 129, 8, -1.5, 0.8), (2, 64, 5, 1.0, 0.75)])
def test_minsum_paths_vs_oracle(Z, B, iters, snr_db, alpha):
    code = QCCode.nr_2_0(Z)
    llr = oracle.awgn_llr(None, B, code.N, snr_db, seed=Z * 1000 + B)
    o = oracle.decode(code.shifts, Z, llr, iters, "minsum", alpha)
    soft, hard = run(MinSumScaledDecoder(code, iters, alpha, early_stopping=False, path="exact"), llr)
    assert np.array_equal(soft, o["beliefs"]) and np.array_equal(hard, o["hard"])
    if code.has_fast_path(dev()):
        of = oracle.decode(code.shifts, Z, llr, iters, "minsum", alpha, order="fast")
        soft, hard = run(MinSumScaledDecoder(code, iters, alpha, early_stopping=False, path="fast"), llr)
        assert np.array_equal(soft, of["beliefs"]) and np.array_equal(hard, of["hard"])
    else:
        assert Z == 2                                    # every other lifting size of BG2 set 0 (4, 8, 16, 32) is specialised
    assert code.has_fast_path(dev()) == (Z in (4, 8, 16, 32))
    if Z in (8, 16):                                     # sum-product on the new specialisations: decisions equal the oracle's
        ob = oracle.decode(code.shifts, Z, llr, iters, "bp")
        _, hb = run(BeliefPropagationDecoder(code, iters, early_stopping=False, path="fast"), llr)
        assert (hb != ob["hard"]).mean() <= 1e-4
        # and the fused simulation (on-chip channel) runs on them too
        from ldpc_b200.sim import simulate_fer
        pt = simulate_fer(code, [snr_db], 4096, iters=iters, alpha=alpha, seed=3, device=dev())[0]
        llr2 = oracle.awgn_llr(None, 4096, code.N, snr_db, 3)
        o2 = oracle.decode(code.shifts, Z, llr2, iters, "minsum", alpha, order="fast")
        assert abs(pt["frame_errors"] - int((o2["hard"].sum(axis=1) > 0).sum())) <= max(2, 0.002 * 4096)


def test_exact_zeros_ties_and_negative_zero():
    """sign(0)=0 (reference :217), ties between equal magnitudes, -0.0 < 0 is False."""
    code = QCCode.nr_2_0(32)
    rng = np.random.default_rng(3)
    llr = rng.choice(np.array([-2.0, -1.0, -0.0, 0.0, 1.0, 1.0, 2.0, 3.0], dtype=np.float32), size=(24, code.N))
    llr[0] = 0.0
    llr[1] = -0.0
    llr[2] = 1.5
    for path, order in (("exact", "reference"), ("fast", "fast")):
        o = oracle.decode(code.shifts, 32, llr, 6, "minsum", 0.75, order=order)
        soft, hard = run(MinSumScaledDecoder(code, 6, 0.75, early_stopping=False, path=path), llr)
        assert np.array_equal(soft, o["beliefs"]), path
        assert np.array_equal(hard, o["hard"]), path
        assert not hard[0].any() and not hard[1].any()


def test_bp_inf_nan_semantics():
    """Saturated tanh -> atanh(1)=inf, inf-inf=NaN, NaN<0 False -> bit 0."""
    code = QCCode.nr_2_0(32)
    llr = oracle.awgn_llr(None, 32, code.N, 4.0, seed=5)          # high SNR: |llr| ~ 5..30, saturates
    llr[0, :64] = -llr[0, :64]
    o = oracle.decode(code.shifts, 32, llr, 10, "bp")
    soft, hard = run(BeliefPropagationDecoder(code, 10, early_stopping=False, path="exact"), llr)
    assert np.isinf(soft).any()
    assert np.array_equal(soft, o["beliefs"], equal_nan=True) and np.array_equal(hard, o["hard"])
    assert not hard[np.isnan(soft)].any()


def test_non_qc_toy_matrix_and_dense_H_constructor():
    """Decoder(H) with the notebook's 3x4 H (Z=1 tables) and with a dense BG2 Z=4 H."""
    H = torch.tensor([[1, 1, 0, 0], [0, 1, 1, 1], [1, 0, 0, 1]], dtype=torch.float32)
    llr = np.array([[0.0, 0.0, 0.0, 0.0], [-0.3, 0.7, -0.9, 0.2], [2.0, -1.0, 0.5, 0.1]], dtype=np.float32)
    shifts = QCCode.from_dense(H).shifts
    for cls, algo in ((MinSumScaledDecoder, "minsum"), (BeliefPropagationDecoder, "bp")):
        dec = cls(H, max_iterations=4, early_stopping=False)
        soft, hard = run(dec, llr)
        o = oracle.decode(shifts, 1, llr, 4, algo, 0.75)
        assert np.array_equal(soft, o["beliefs"], equal_nan=True) and np.array_equal(hard, o["hard"])
    code = QCCode.nr_2_0(4)
    dec = MinSumScaledDecoder(code.dense(), max_iterations=3, early_stopping=False)
    assert dec.code.Z == 4 and np.array_equal(dec.code.shifts, code.shifts)


def test_output_formats_syndrome_and_per_codeword_stop():
    code = QCCode.nr_2_0(32)
    B, iters = 70, 12
    llr = oracle.awgn_llr(None, B, code.N, -1.5, seed=11)
    llr_t = torch.from_numpy(llr).to(dev())
    o = oracle.decode(code.shifts, 32, llr, iters, "minsum", 0.75, want_mask=True, stop_when_valid=True)
    for path, order in (("exact", "reference"), ("fast", "fast")):
        of = oracle.decode(code.shifts, 32, llr, iters, "minsum", 0.75, order=order, want_mask=True)
        dec = MinSumScaledDecoder(code, iters, 0.75, early_stopping=False, path=path)
        for dtype in (_native.HARD_F32, _native.HARD_U8, _native.HARD_PACKED):
            _, hard, syn, its, _ = dec._launch(llr_t, dev(), iters, hard_dtype=dtype, syndrome=True, iters_out=True)
            h = hard.cpu().numpy()
            if dtype == _native.HARD_PACKED:
                h = np.unpackbits(h.view(np.uint8), axis=1, bitorder="little")[:, :code.N]
            assert np.array_equal(h.astype(np.uint8), of["hard"]), (path, dtype)
            last = ((of["valid_mask"][:, 0] >> np.uint64(iters - 1)) & np.uint64(1)).astype(np.uint8)
            assert np.array_equal(syn.cpu().numpy(), last)
            assert (its.cpu().numpy() == iters).all()
    # per-codeword early exit (exact path): iteration counts and frozen outputs as the oracle
    dec = MinSumScaledDecoder(code, iters, 0.75, early_stopping=True)
    bits, its, ok = dec.decode_with_iterations(llr_t)
    assert np.array_equal(its.cpu().numpy(), o["iters"])
    assert np.array_equal(bits.cpu().numpy().astype(np.uint8), o["hard"])
    assert np.array_equal(ok.cpu().numpy(), o["iters"] < iters) or ok.cpu().numpy()[o["iters"] == iters].any()
    # validity mask of the C ABI == oracle's
    _, _, _, _, mask = MinSumScaledDecoder(code, iters, 0.75, early_stopping=False)._launch(llr_t, dev(), iters, mask=True)
    ref_mask = oracle.decode(code.shifts, 32, llr, iters, "minsum", 0.75, want_mask=True)["valid_mask"]
    assert np.array_equal(mask.cpu().numpy().view(np.uint64), ref_mask)


def test_empty_ragged_and_cpu_inputs():
    code = QCCode.nr_2_0(4)
    dec = MinSumScaledDecoder(code, 3, 0.75, early_stopping=False)
    bits, it = dec.decode(torch.zeros((0, code.N), device=dev()))
    assert bits.shape == (0, code.N) and it == 3
    with pytest.raises(ValueError):
        dec.decode(torch.zeros((2, code.N + 1), device=dev()))
    llr = oracle.awgn_llr(None, 13, code.N, 1.0, seed=2)      # 13 is not a multiple of 8 codewords/warp
    bits_cpu, _ = dec.decode(torch.from_numpy(llr))           # CPU tensor: staged through the GPU
    assert not bits_cpu.is_cuda
    o = oracle.decode(code.shifts, 4, llr, 3, "minsum", 0.75, order="fast")
    assert np.array_equal(bits_cpu.numpy().astype(np.uint8), o["hard"])
    # float64 / non-contiguous input is converted like the reference's float tensors
    bits2, _ = dec.decode(torch.from_numpy(llr.astype(np.float64)).to(dev()))
    assert np.array_equal(bits2.cpu().numpy().astype(np.uint8), o["hard"])


def test_host_buffer_entry_point():
    """ldpc_decode_host (the e2e path of bench.py): pinned host LLRs in, host bits out."""
    import ctypes as C
    code = QCCode.nr_2_0(32)
    B, iters = 3000, 10
    llr = oracle.awgn_llr(None, B, code.N, -2.0, seed=21)
    pinned = torch.from_numpy(llr).pin_memory()
    hard = torch.empty((B, (code.N + 31) // 32), dtype=torch.int32).pin_memory()
    soft = torch.empty((B, code.N), dtype=torch.float32).pin_memory()
    _native.check(_native.lib().ldpc_decode_host(code.handle(dev()), _native.ALGO_MINSUM, _native.ptr(pinned), B, iters, 0.75,
                                                 _native.PATH_AUTO, _native.ptr(soft), _native.ptr(hard),
                                                 _native.HARD_PACKED, 1024))
    o = oracle.decode(code.shifts, 32, llr, iters, "minsum", 0.75, order="fast")
    h = np.unpackbits(hard.numpy().view(np.uint8), axis=1, bitorder="little")[:, :code.N]
    assert np.array_equal(h, o["hard"]) and np.array_equal(soft.numpy(), o["beliefs"])


@pytest.mark.parametrize("fmt", ["i8", "f16"])
def test_host_buffer_quantised_llrs(fmt):
    """ldpc_decode_host_q: 8-bit / half-precision host LLRs decode bit-identically to the fp32 values they stand for
    (llr = q * scale), on batches that exercise the chunking (ragged last chunk) and both output kinds."""
    code = QCCode.nr_2_0(32)
    B, iters = 2500, 10
    llr = oracle.awgn_llr(None, B, code.N, -2.0, seed=33)
    if fmt == "i8":
        scale = np.float32(0.25)
        q = np.clip(np.rint(llr / scale), -127, 127).astype(np.int8)
        deq = q.astype(np.float32) * scale
        raw, code_fmt = torch.from_numpy(q).pin_memory(), _native.LLR_I8
    else:
        scale = np.float32(1.0)
        q = llr.astype(np.float16)
        deq = q.astype(np.float32)
        raw, code_fmt = torch.from_numpy(q).pin_memory(), _native.LLR_F16
    hard = torch.empty((B, (code.N + 31) // 32), dtype=torch.int32).pin_memory()
    soft = torch.empty((B, code.N), dtype=torch.float32).pin_memory()
    _native.check(_native.lib().ldpc_decode_host_q(code.handle(dev()), _native.ALGO_MINSUM, _native.ptr(raw), code_fmt, float(scale),
                                                   B, iters, 0.75, _native.PATH_AUTO, _native.ptr(soft), _native.ptr(hard),
                                                   _native.HARD_PACKED, 1024))
    o = oracle.decode(code.shifts, 32, deq, iters, "minsum", 0.75, order="fast")
    h = np.unpackbits(hard.numpy().view(np.uint8), axis=1, bitorder="little")[:, :code.N]
    assert np.array_equal(h, o["hard"]) and np.array_equal(soft.numpy(), o["beliefs"])
    with pytest.raises(_native.LdpcError):
        _native.check(_native.lib().ldpc_decode_host_q(code.handle(dev()), _native.ALGO_MINSUM, _native.ptr(raw), 7, 1.0, B, iters,
                                                       0.75, _native.PATH_AUTO, None, _native.ptr(hard), _native.HARD_PACKED, 0))


def test_linearity_at_full_batch_size():
    """Size-independent property at a bench-sized batch: min-sum is odd-symmetric under a
    codeword flip -- decoding llr*(1-2c) for a codeword c gives the decisions XOR c -- and
    the all-zero codeword decodes to zero with a valid syndrome at high SNR."""
    code = QCCode.nr_2_0(32)
    B, iters = 1 << 15, 10
    ch_llr = torch.empty((B, code.N), dtype=torch.float32, device=dev())
    _native.check(_native.lib().ldpc_awgn_llr(None, B, code.N, 1.0, 99, 0, _native.ptr(ch_llr), _native.stream_ptr(dev())))
    dec = MinSumScaledDecoder(code, iters, 0.75, early_stopping=False)
    _, hard, syn, _, _ = dec._launch(ch_llr, dev(), iters, hard_dtype=_native.HARD_U8, syndrome=True)
    assert int(hard.sum()) == 0 and bool(syn.all())
    # a non-zero codeword from the decoder itself: decode noise-free +-1 pattern found by flipping
    base = oracle.awgn_llr(None, 8, code.N, -3.5, seed=77)
    o = oracle.decode(code.shifts, 32, base, 50, "minsum", 0.75, want_mask=True)
    valid = [b for b in range(8) if ((o["valid_mask"][b, 0] >> np.uint64(49)) & np.uint64(1)) and o["hard"][b].any()]
    if valid:
        c = torch.from_numpy(o["hard"][valid[0]].astype(np.float32)).to(dev())
        flipped = ch_llr[:4096] * (1.0 - 2.0 * c)
        _, h2, s2, _, _ = dec._launch(flipped.contiguous(), dev(), iters, hard_dtype=_native.HARD_U8, syndrome=True)
        assert bool((h2.float() == c).all()) and bool(s2.all())


def test_syndrome_check_entry_point():
    """_check_valid_codeword (reference :262-284) through ldpc_syndrome_check, all bit formats."""
    code = QCCode.nr_2_0(32)
    dec = MinSumScaledDecoder(code, 10, 0.75, early_stopping=False)
    llr = oracle.awgn_llr(None, 50, code.N, -2.5, seed=4)
    o = oracle.decode(code.shifts, 32, llr, 10, "minsum", 0.75, order="fast", want_mask=True)
    want = ((o["valid_mask"][:, 0] >> np.uint64(9)) & np.uint64(1)).astype(bool)
    assert want.any() and not want.all()
    bits = torch.from_numpy(o["hard"].astype(np.float32)).to(dev())
    assert np.array_equal(dec._check_valid_codeword(bits).cpu().numpy(), want)
    assert np.array_equal(dec._check_valid_codeword(bits.to(torch.uint8)).cpu().numpy(), want)
    assert np.array_equal(dec._check_valid_codeword(bits.cpu()).numpy(), want)
    toy = MinSumScaledDecoder(torch.tensor([[1, 1, 0, 0], [0, 1, 1, 1], [1, 0, 0, 1]], dtype=torch.float32), 2, early_stopping=False)
    cw = torch.tensor([[0, 0, 0, 0], [1, 1, 1, 0], [1, 1, 0, 1], [1, 0, 0, 0]], dtype=torch.float32, device=dev())
    assert toy._check_valid_codeword(cw).tolist() == [True, False, True, False]


@pytest.mark.parametrize("name", CLASSIC)
def test_bp_fast_path_against_reference(name):
    """Specialised sum-product kernel (path="fast": CUDA tanhf/atanhf, 2-3 ulp, and the (F, K)
    finite-sum / non-finite-count variable update).  Criterion ("BP numerics"):
    hard decisions identical to the reference, identical inf/NaN pattern except where a product
    sits within a few ulp of the saturation point, finite beliefs within 1e-3 relative."""
    g = load_golden(name)
    Z = int(g["Z"])
    code = QCCode.nr_2_0(Z)
    dec = BeliefPropagationDecoder(code, int(g["iters"]), early_stopping=False, path="fast")
    soft, hard = run(dec, g["llr"])
    ref = g["bp_beliefs"]
    ref_hard = unpack(g["bp_bits"], code.N)
    assert (hard != ref_hard).mean() <= 1e-4, int((hard != ref_hard).sum())
    cls = lambda x: np.where(np.isnan(x), 3, np.where(np.isposinf(x), 1, np.where(np.isneginf(x), 2, 0)))
    assert (cls(soft) != cls(ref)).mean() <= 2e-3, int((cls(soft) != cls(ref)).sum())
    fin = np.isfinite(ref) & np.isfinite(soft)
    if fin.any():
        rel = np.abs(soft[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1.0)
        assert np.quantile(rel, 0.999) <= 1e-3 and rel.max() <= 5e-2, (float(np.quantile(rel, 0.999)), float(rel.max()))


def test_bp_fast_path_inf_nan_and_frames():
    """High-SNR batch (saturating messages, inf-inf -> NaN -> bit 0) and a low-SNR batch: frame and
    bit error counts of the fast kernel equal the oracle's."""
    code = QCCode.nr_2_0(32)
    for snr_db, seed in ((4.0, 5), (-3.0, 6), (-2.0, 8)):
        llr = oracle.awgn_llr(None, 256, code.N, snr_db, seed=seed)
        if snr_db > 0:
            llr[0, :64] = -llr[0, :64]
        o = oracle.decode(code.shifts, 32, llr, 10, "bp")
        soft, hard = run(BeliefPropagationDecoder(code, 10, early_stopping=False, path="fast"), llr)
        assert not hard[np.isnan(soft)].any()
        assert (hard.sum(axis=1) > 0).sum() == (o["hard"].sum(axis=1) > 0).sum()
        assert abs(int(hard.sum()) - int(o["hard"].sum())) <= 2


def test_bench_line_contract_small_run():
    """bench.py end to end on a small batch: the JSON line carries every key of the contract (the driver's run uses
    the default 2^20-codeword batch; this guards the plumbing: events, e2e legs, clocks sampler, roofline, CPU arm)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "3", "--warmup", "3", "--batch", "65536",
                          "--e2e-steps", "1", "--cpu-seconds", "2", "--no-pytorch-baseline"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-3000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["value"] > 1.0 and d["gpu_launches"] >= 3 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"]) and 0 < d["roofline"]["frac"] <= 1.0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0 and d["fer"]["frames"] == 3 * 65536
    assert d["roofline"]["counts"]["file"] == "profiles/r2_minsum_counts.json"
    assert 0 < d["e2e"]["roofline"]["frac"] <= 1.05 and d["e2e"]["roofline"]["h2d_peak_gbs_at_N"] > 1.0
    par = d["parity"]
    assert par["frames"] == 65536 and par["converged_frames"] + par["nonconverged_frames"] == 65536
    assert par["frames_over_1e-4_converged"] == 0 and par["frames_with_hard_mismatch"] <= par["nonconverged_frames"]


@pytest.mark.parametrize("algo,Z", [("minsum", 32), ("bp", 32), ("minsum", 16), ("bp", 16), ("minsum", 8), ("minsum", 4)])
def test_fast_kernel_per_codeword_early_exit(algo, Z):
    """The specialised kernels' early-exit variant: every codeword stops after its first valid iteration (with Z < 32 the 32/Z
    codewords of a warp freeze their decisions individually).  Hard decisions, iteration counts and the syndrome flag equal
    the oracle run with the same stopping rule (min-sum: in the fast kernel's operation order), on a ragged batch that mixes
    early, late and never-converging codewords.  Z = 16 with 50 iterations and early stopping is the reference's default."""
    code = QCCode.nr_2_0(Z)
    B, iters = 301, 20
    lo, hi = (-3.6, -1.0) if Z == 32 else ((-2.5, 0.5) if Z >= 8 else (-1.0, 3.0))
    llr = np.concatenate([oracle.awgn_llr(None, 150, code.N, lo, seed=41), oracle.awgn_llr(None, 151, code.N, hi, seed=42)])
    rng = np.random.default_rng(Z)
    llr = llr[rng.permutation(B)]                       # early and late codewords share warps
    if algo == "minsum":
        dec = MinSumScaledDecoder(code, iters, 0.75, early_stopping=True, check_finite=False)   # auto -> specialised kernel
        o = oracle.decode(code.shifts, Z, llr, iters, "minsum", 0.75, order="fast", stop_when_valid=True)
    else:
        dec = BeliefPropagationDecoder(code, iters, early_stopping=True, path="fast")
        o = oracle.decode(code.shifts, Z, llr, iters, "bp", 1.0, stop_when_valid=True)
    assert len(set(o["iters"].tolist())) > 2
    n0 = _native.lib().ldpc_launch_count()
    bits, its, ok = dec.decode_with_iterations(torch.from_numpy(llr).to(dev()))
    assert _native.lib().ldpc_launch_count() - n0 == 1                                   # ONE launch of the specialised kernel
    assert np.array_equal(its.cpu().numpy(), o["iters"])
    assert np.array_equal(bits.cpu().numpy().astype(np.uint8), o["hard"])
    valid = np.asarray(dec._check_valid_codeword(bits).cpu().numpy())
    assert np.array_equal(ok.cpu().numpy(), valid)
    assert 0 < int((its.cpu().numpy() < iters).sum()) and int((~valid).sum()) > 0          # the batch really is mixed
    # the exact kernel (reference operation order) under the same rule: every codeword that converges does so at the same
    # iteration with the same decisions; frames that never converge are chaotic (see fast_order_tolerance_ok) and their
    # decisions after 20 iterations may differ between the two operation orders
    ex = type(dec)(code, iters, early_stopping=True, path="exact", **({"scaling_factor": 0.75} if algo == "minsum" else {}))
    bits_x, its_x, ok_x = ex.decode_with_iterations(torch.from_numpy(llr).to(dev()))
    assert torch.equal(ok_x, ok)
    assert torch.equal(its_x[ok_x], its[ok_x]) and torch.equal(bits_x[ok_x], bits[ok_x])


def test_batch_global_early_stopping_fast_route():
    """decode() with the reference's batch-global rule on the specialised kernels (two passes) == the exact
    validity-mask route, for a batch that converges (stop = slowest codeword) and one that does not (stop = max)."""
    code = QCCode.nr_2_0(32)
    for snr, seed in ((0.0, 7), (-3.6, 8)):
        llr = torch.from_numpy(oracle.awgn_llr(None, 96, code.N, snr, seed=seed)).to(dev())
        fast = MinSumScaledDecoder(code, 30, 0.75, early_stopping=True)
        exact = MinSumScaledDecoder(code, 30, 0.75, early_stopping=True, path="exact")
        bf, tf = fast.decode(llr)
        bx, tx = exact.decode(llr)
        conv_t = exact._check_valid_codeword(bx)
        assert tf == tx and torch.equal(bf[conv_t], bx[conv_t])              # never-converging frames: chaotic, see above
        assert (tf < 30) == (snr == 0.0) and (snr != 0.0 or bool(conv_t.all()))
        sf, hf = fast.forward(llr)
        sx, hx = exact.forward(llr)
        conv = np.asarray(conv_t.cpu().numpy())
        assert torch.equal(hf[conv_t], hx[conv_t])
        if snr == 0.0:       # quick convergence: the two operation orders stay within the fast path's soft tolerance
            assert fast_order_tolerance_ok(sf.cpu().numpy(), sx.cpu().numpy(), conv)


def test_fast_kernel_against_reference_order_kernel_at_scale():
    """the headline kernel (v2c = posterior - own message) against the REFERENCE operation order
    (path="exact", bit-identical to the reference's beliefs on the golden fixtures above) on 2^18 bench frames --
    the same comparison bench.py prints for its 2^20 frames as `parity`.  Criteria: soft outputs within 1e-4 relative
    on every frame that converged; hard decisions may differ only on frames that did not converge (the flooding
    iteration is chaotic there), and the two frame-error rates are statistically indistinguishable."""
    import bench
    code = QCCode.nr_2_0(32)
    B = 1 << 18
    L = _native.lib()
    llr = torch.empty((B, code.N), dtype=torch.float32, device=dev())
    _native.check(L.ldpc_awgn_llr(None, B, code.N, bench.SNR_DB, 1234, 0, _native.ptr(llr), _native.stream_ptr(dev())))
    par = bench.parity_block(L, code.handle(dev()), code, llr, "minsum", dev())
    assert par["frames"] == B and par["nonfinite_class_mismatches"] == 0
    assert par["frames_over_1e-4_converged"] == 0, par
    assert par["frames_with_hard_mismatch"] <= par["nonconverged_frames"], par
    assert par["hard_bit_mismatches"] <= 1e-6 * B * code.N, par
    lo_e, hi_e = par["fer_exact_wilson95"]
    lo_f, hi_f = par["fer_fast_wilson95"]
    assert max(lo_e, lo_f) <= min(hi_e, hi_f), par
    assert abs(par["frame_errors_exact"] - par["frame_errors_fast"]) <= 2, par


def test_non_finite_llrs_against_the_reference_golden():
    """The reference's own outputs on +-inf / NaN / huge LLRs (tests/golden/nonfinite_z4_b24.npz): the default path
    (auto -> routed to the exact kernel) reproduces its min-sum beliefs bit for bit, NaN pattern included, and its BP
    hard bits and inf/NaN classes."""
    g = load_golden("nonfinite_z4_b24")
    code = QCCode.nr_2_0(4)
    it = int(g["iters"])
    soft, hard = run(MinSumScaledDecoder(code, it, 0.75, early_stopping=False), g["llr"])
    assert np.array_equal(soft, g["ms_beliefs"], equal_nan=True)
    assert np.array_equal(hard, unpack(g["ms_bits"], code.N))
    soft, hard = run(BeliefPropagationDecoder(code, it, early_stopping=False), g["llr"])
    ref = g["bp_beliefs"]
    assert np.array_equal(hard, unpack(g["bp_bits"], code.N))
    for f in (np.isnan, np.isposinf, np.isneginf):
        assert np.array_equal(f(ref), f(soft))
    fin = np.isfinite(ref)
    assert np.all(np.abs(soft[fin] - ref[fin]) <= 2e-4 * np.maximum(np.abs(ref[fin]), 1.0))


def test_auto_path_routes_non_finite_llrs_to_the_reference_order_kernel():
    """+-inf channel LLRs (hard-decision inputs) make posterior - own message an inf - inf on the specialised
    kernel; the drop-in class sends such batches to the exact kernel, so `auto` equals `exact` bit for bit (NaN == NaN)
    and equals the reference-order oracle."""
    code = QCCode.nr_2_0(32)
    llr = oracle.awgn_llr(None, 64, code.N, 1.0, seed=11)
    llr = np.where(llr >= 0, np.inf, -np.inf).astype(np.float32)          # hard-decision input
    llr[1, :] = oracle.awgn_llr(None, 1, code.N, 1.0, seed=12)[0]           # one ordinary frame in the batch
    llr[2, 5] = 1e30
    auto = MinSumScaledDecoder(code, 10, 0.75, early_stopping=False)
    exact = MinSumScaledDecoder(code, 10, 0.75, early_stopping=False, path="exact")
    sa, ha = run(auto, llr)
    se, he = run(exact, llr)
    assert np.array_equal(sa, se, equal_nan=True) and np.array_equal(ha, he)
    o = oracle.decode(code.shifts, 32, llr, 10, "minsum", 0.75)
    assert np.array_equal(se, o["beliefs"], equal_nan=True) and np.array_equal(he, o["hard"])
    # finite batches still take the specialised kernel
    fin = oracle.awgn_llr(None, 64, code.N, -2.0, seed=13)
    sf, _ = run(auto, fin)
    assert np.array_equal(sf, oracle.decode(code.shifts, 32, fin, 10, "minsum", 0.75, order="fast")["beliefs"])


@pytest.mark.parametrize("Z,B,iters", [(64, 37, 6), (96, 9, 4), (48, 50, 5)])
def test_lifting_factors_above_32_decode_like_the_oracle(Z, B, iters):
    """Z > 32 through the classic decoders' (base_graph, Z) constructor: the code is held as an equivalent code with a smaller lifting
    factor (tests/test_host_logic.py) and inputs / outputs are renumbered, so beliefs, decisions and validity must equal the
    oracle's run on the natural code -- bit for bit for min-sum (the operation order is preserved), decisions for BP."""
    rng = np.random.default_rng(Z)
    support = QCCode.nr_2_0(32).shifts >= 0
    base = np.where(support, rng.integers(0, Z, size=support.shape), -1)
    llr = oracle.awgn_llr(None, B, 52 * Z, 1.0, seed=Z)
    o = oracle.decode(base, Z, llr, iters, "minsum", 0.75)
    dec = MinSumScaledDecoder(base_graph=torch.from_numpy(base.astype(np.float32)), Z=Z, max_iterations=iters, scaling_factor=0.75,
                              early_stopping=False)
    soft, hard = run(dec, llr)
    assert np.array_equal(soft, o["beliefs"]) and np.array_equal(hard, o["hard"])
    ok = dec._check_valid_codeword(torch.from_numpy(o["hard"]).to(dev())).cpu().numpy()
    H = dec.code.dense().numpy()
    assert np.array_equal(ok, ((o["hard"].astype(np.int64) @ H.T) % 2 == 0).all(axis=1))
    bits, its, syn = dec.decode_with_iterations(torch.from_numpy(llr).to(dev()))
    assert np.array_equal(bits.cpu().numpy(), o["hard"]) and np.array_equal(syn.cpu().numpy(), ok)
    ob = oracle.decode(base, Z, llr, iters, "bp")
    _, hb = run(BeliefPropagationDecoder(base_graph=torch.from_numpy(base.astype(np.float32)), Z=Z, max_iterations=iters,
                                         early_stopping=False), llr)
    assert np.array_equal(hb, ob["hard"])
    # early stopping: the reference's batch-global rule on a batch that converges
    llr_hi = oracle.awgn_llr(None, 8, 52 * Z, 6.0, seed=Z + 1)
    oe = oracle.decode(base, Z, llr_hi, 20, "minsum", 0.75, want_mask=True)
    t = oracle.first_all_valid(oe["valid_mask"], 20)
    hard_e, n_it = MinSumScaledDecoder(base_graph=torch.from_numpy(base.astype(np.float32)), Z=Z, max_iterations=20).decode(
        torch.from_numpy(llr_hi).to(dev()))
    assert t is not None and n_it == t + 1
    assert np.array_equal(hard_e.cpu().numpy(), oracle.decode(base, Z, llr_hi, t + 1, "minsum", 0.75)["hard"])


def test_codes_too_large_for_shared_memory_use_the_global_workspace():
    """Z = 384 = 12 x 32 (2 364 + 2 x 624 cell rows of 128 bytes per codeword) and a 3 328-bit code without quasi-cyclic structure
    (Z = 1: 12 608 + 2 x 3 328 rows per warp of 32 codewords) do not fit an SM's shared memory: the table-driven kernel then
    keeps its per-warp state in a global workspace.  Same code, same operation order: beliefs bit-identical to the oracle.
    The second case is also the reference's own call shape -- a dense H handed to the constructor (main.py:92,152)."""
    rng = np.random.default_rng(5)
    support = QCCode.nr_2_0(32).shifts >= 0
    base = np.where(support, rng.integers(0, 384, size=support.shape), -1)
    llr = oracle.awgn_llr(None, 5, 52 * 384, 1.0, seed=9)
    o = oracle.decode(base, 384, llr, 3, "minsum", 0.75)
    dec = MinSumScaledDecoder(base_graph=torch.from_numpy(base.astype(np.float32)), Z=384, max_iterations=3, early_stopping=False)
    soft, hard = run(dec, llr)
    assert np.array_equal(soft, o["beliefs"]) and np.array_equal(hard, o["hard"])
    ok = dec._check_valid_codeword(torch.from_numpy(o["hard"]).to(dev())).cpu().numpy()
    assert np.array_equal(ok, (oracle.decode(base, 384, llr, 3, "minsum", 0.75, want_mask=True)["valid_mask"][:, 0] >> np.uint64(2)) & np.uint64(1) == 1)
    # dense H = BG2 lifted with Z = 64: recognised as such (renumbered 32-circulants, shared memory)
    from ldpc_b200.utils.ldpc_utils import expand_base_matrix
    base64 = np.where(support, rng.integers(0, 64, size=support.shape), -1)
    H = expand_base_matrix(torch.from_numpy(base64.astype(np.float32)), 64)
    llr = oracle.awgn_llr(None, 33, 52 * 64, 0.5, seed=10)
    o = oracle.decode(base64, 64, llr, 4, "minsum", 0.75)
    dec = MinSumScaledDecoder(H, max_iterations=4, early_stopping=False)
    assert dec.code.lift_Z == 64 and dec.code.Z == 32
    soft, hard = run(dec, llr)
    assert np.array_equal(soft, o["beliefs"]) and np.array_equal(hard, o["hard"])
    # the same H with its columns shuffled: no quasi-cyclic structure left, Z = 1, global workspace; BP too
    perm = rng.permutation(H.shape[1])
    Hp = H[:, perm]
    dec = MinSumScaledDecoder(Hp, max_iterations=4, early_stopping=False)
    assert dec.code.Z == 1 and dec.code.lift_Z is None
    shifts1 = np.where(Hp.numpy() > 0, 0, -1)
    o = oracle.decode(shifts1, 1, llr[:, perm], 4, "minsum", 0.75)
    soft, hard = run(dec, llr[:, perm])
    assert np.array_equal(soft, o["beliefs"]) and np.array_equal(hard, o["hard"])
    ob = oracle.decode(shifts1, 1, llr[:, perm], 4, "bp")
    _, hb = run(BeliefPropagationDecoder(Hp, max_iterations=4, early_stopping=False), llr[:, perm])
    assert np.array_equal(hb, ob["hard"])
