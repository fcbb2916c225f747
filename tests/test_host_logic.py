"""Host-side logic of the drop-in package: tables, QC factorisation, edge numbering. CPU only."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import load_golden, ROOT
import ldpc_b200
from ldpc_b200.utils import QCCode, create_LLR_mapping, expand_base_matrix, load_base_matrix, as_code

REF_TABLES = "/root/reference/5G LDPC CODES"


def test_notebook_cell7_known_answer():
    """The only golden vector the reference itself prints (EE4002R_2025.ipynb cell 7)."""
    H = torch.tensor([[1, 1, 0, 0], [0, 1, 1, 1], [1, 0, 0, 1]], dtype=torch.float32)
    m, c, v, o = create_LLR_mapping(H.T)
    assert m.tolist() == [[0, 2, -1, -1], [-1, 3, 4, 5], [1, -1, -1, 6]]
    assert c.tolist() == [[2, -1], [6, -1], [0, -1], [4, 5], [3, 5], [3, 4], [1, -1]]
    assert v.tolist() == [[1], [0], [3], [2], [-1], [6], [5]]
    assert o.tolist() == [[0, 0, 1, 1, 2, 3, 3]]


def test_mapping_matches_reference_on_bg2_z4():
    g = load_golden("mapping_layers")
    code = QCCode.nr_2_0(4)
    m, c, v, o = create_LLR_mapping(code.dense().T)
    assert np.array_equal(m.numpy(), g["z4_map"])
    assert np.array_equal(c.numpy(), g["z4_check"])
    assert np.array_equal(v.numpy(), g["z4_var"])
    assert np.array_equal(o.numpy(), g["z4_out"])
    # dense H of the engine's table == the reference's expand_base_matrix output
    H = np.unpackbits(g["z4_H"], axis=1)[:, :code.N]
    assert np.array_equal(code.dense().numpy().astype(np.uint8), H)
    assert np.array_equal(code.shifts, g["z4_base"].astype(np.int16))


@pytest.mark.parametrize("Z", [4, 32])
def test_builtin_table_equals_reference_file(Z):
    path = os.path.join(REF_TABLES, f"NR_2_0_{Z}.txt")
    if not os.path.exists(path):
        pytest.skip("reference tables are only present in the build container")
    base = load_base_matrix(path)
    assert np.array_equal(QCCode.nr_2_0(Z).shifts, base.numpy().astype(np.int16))
    assert np.array_equal(QCCode.from_file(path, Z).shifts, QCCode.nr_2_0(Z).shifts)


@pytest.mark.parametrize("Z", [1, 3, 4, 7, 32])
def test_factor_dense_roundtrip(Z):
    rng = np.random.default_rng(Z)
    base = rng.integers(-1, Z, size=(5, 9))
    base[rng.random(base.shape) < 0.5] = -1
    base[0, 0] = Z - 1
    H = expand_base_matrix(torch.tensor(base, dtype=torch.float32), Z)
    for i, j in zip(*np.nonzero(base >= 0)):   # lifting convention ldpc_utils.py:121-123
        r = int(rng.integers(0, Z))
        assert H[i * Z + r, j * Z + (r + base[i, j]) % Z] == 1
    code = QCCode.from_dense(H, Z)
    assert code.Z == Z and np.array_equal(code.shifts, base)
    assert np.array_equal(code.dense().numpy(), H.numpy())
    chk, var = code.edges()
    assert np.array_equal(H.numpy()[chk, var], np.ones(len(chk), dtype=np.float32)) and len(chk) == int(H.sum())
    assert np.all(np.diff(chk) >= 0)


def test_factor_autodetects_and_rejects():
    code = QCCode.nr_2_0(32)
    assert QCCode.from_dense(code.dense()).Z == 32
    H = torch.tensor([[1, 1, 0, 0], [0, 1, 1, 1], [1, 0, 0, 1]], dtype=torch.float32)
    assert QCCode.from_dense(H).Z == 1                       # any binary H is QC with Z=1
    with pytest.raises(ValueError):
        QCCode.from_dense(code.dense(), Z=16)                # not QC with this lifting factor
    with pytest.raises(ValueError):
        QCCode.from_dense(torch.full((4, 4), 2.0))
    with pytest.raises(ValueError):
        QCCode(np.array([[0, 5]]), 4)
    with pytest.raises(ValueError):
        QCCode(np.array([[0, 1]]), 64)
    with pytest.raises(ValueError):
        as_code(None, base_graph=np.zeros((2, 2)))
    assert as_code(code) is code


def test_oracle_is_not_reachable_from_the_product():
    """The product package must never import or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "ldpc-neuralnetwork-decoder_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle/|ldpc_oracle", re.M)
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                hits = [m.group(0) for m in pat.finditer(text)]
                # comments may cite the oracle file by name; imports/paths may not appear in code
                code_lines = [l for l in text.splitlines() if pat.search(l) and not l.strip().startswith(("//", "#", "*", "\"\"\""))
                              and "oracle/ldpc_oracle.c restates" not in l and "restated in oracle" not in l]
                assert not code_lines, (f, code_lines)


def test_cpu_tensor_without_gpu_raises():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ldpc_b200.models import MinSumScaledDecoder
    dec = MinSumScaledDecoder(QCCode.nr_2_0(4), max_iterations=2, early_stopping=False)
    with pytest.raises(RuntimeError):
        dec.decode(torch.zeros(1, 208))


def test_qpsk_helpers_equal_the_reference():
    """qpsk_modulate / qpsk_demodulate (utils/channel.py:4-60, 91-154) against golden vectors produced by the
    reference itself (oracle/make_golden.py:qpsk): bit-identical, including odd lengths and un-batched input."""
    from ldpc_b200.utils import qpsk_modulate, qpsk_demodulate, awgn_channel
    g = load_golden("qpsk")
    for tag in ("even", "odd"):
        bits = torch.from_numpy(g[f"bits_{tag}"]).float()
        sym = qpsk_modulate(bits)
        assert np.array_equal(torch.view_as_real(sym).numpy(), g[f"sym_{tag}"])
        rx = torch.view_as_complex(torch.from_numpy(g[f"rx_{tag}"]).contiguous())
        for snr in (-2.0, 1.5, 6.0):
            assert np.array_equal(qpsk_demodulate(rx, snr).numpy(), g[f"llr_{tag}_snr{snr}"])
    one = qpsk_demodulate(qpsk_modulate(torch.tensor([0., 1., 1.])), 0.0)
    assert one.shape == (4,) and np.array_equal(one.numpy(), g["llr_unbatched"])
    # the noisy chain is statistical: same mean / spread as the reference's run (512 x 208 LLRs)
    torch.manual_seed(11)
    llr = qpsk_demodulate(awgn_channel(qpsk_modulate(torch.zeros(512, 208)), 1.0), 1.0)
    m, sd = g["chain_snr1_mean_std"]
    assert abs(llr.mean().item() - m) < 0.02 * abs(m) and abs(llr.std().item() - sd) < 0.02 * sd


def test_oracle_qpsk_generator_statistics():
    """oracle_qpsk_llr (the CPU restatement the device kernel is checked against): mean 2*a/noise_var, standard
    deviation 2*sqrt(noise_var/2)/noise_var, sign flip for 1-bits, true_llr = sqrt(2) x the reference scaling."""
    from oracle import oracle
    snr_db = 1.0
    nv = 1 / (10 ** (snr_db / 10))
    a = oracle.qpsk_llr(None, 512, 208, snr_db, seed=3)
    assert abs(a.mean() - 2 * (1 / np.sqrt(2)) / nv) < 0.02 and abs(a.std() - 2 * np.sqrt(nv / 2) / nv) < 0.02
    g = load_golden("qpsk")
    m, sd = g["chain_snr1_mean_std"]
    assert abs(a.mean() - m) < 0.02 * abs(m) and abs(a.std() - sd) < 0.02 * sd
    bits = np.ones((4, 208), dtype=np.uint8)
    b = oracle.qpsk_llr(bits, 4, 208, snr_db, seed=3)
    z = oracle.qpsk_llr(None, 4, 208, snr_db, seed=3)
    amp2 = np.float32(2) * np.float32(1 / np.sqrt(2))
    assert np.allclose(z - b, 2 * amp2 / np.float32(nv), rtol=1e-5)
    t = oracle.qpsk_llr(None, 4, 208, snr_db, seed=3, true_llr=True)
    assert np.allclose(t, z * np.sqrt(2), rtol=1e-6)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm): one JSON line on rank 0 with the
    contract's keys, runs without a GPU, and the other ranks print nothing."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "Gbit/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["ms_per_step"] > 0 and d["config"]["workload"].startswith("minsum_bg2_z32_it10")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    other = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=root, env=env)
    assert other.returncode == 0 and other.stdout.strip() == ""


@pytest.mark.parametrize("Z", [4, 32])
def test_encoder_plan_and_oracle_encoder(Z):
    """Host side of the systematic encoder: the structure discovered from the BG2 table (10 information
    blocks, 4 core parity blocks, 38 extension rows), B^-1 really inverts the core block, and the oracle's dense
    GF(2) solve yields systematic codewords in the null space of H."""
    from oracle import oracle
    from ldpc_b200.utils import QCCode, SystematicEncoder
    code = QCCode.nr_2_0(Z)
    enc = SystematicEncoder(code)
    assert (enc.g, enc.kb) == (4, 10) and enc.plan[2] == (4 * Z + 31) // 32
    core_rows, core_cols, ext = enc.plan[3:7], enc.plan[7:11], enc.plan[11:]
    assert list(core_rows) == [0, 1, 2, 3] and list(core_cols) == [10, 11, 12, 13]
    assert list(ext[:4]) == [-1] * 4 and list(ext[4:]) == list(range(14, 52))
    # B * B^-1 = I over GF(2)
    n = 4 * Z
    binv = ((enc.binv_packed[:, np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(np.int64)
    Bm = np.zeros((n, n), dtype=np.int64)
    r = np.arange(Z)
    for a in range(4):
        for b in range(4):
            s = int(code.shifts[a, 10 + b])
            if s >= 0:
                Bm[a * Z + r, b * Z + (r + s) % Z] = 1
    assert np.array_equal((Bm @ binv) & 1, np.eye(n, dtype=np.int64))
    rng = np.random.default_rng(Z)
    info = rng.integers(0, 2, size=(5, code.K), dtype=np.uint8)
    cw = oracle.encode_dense(code.shifts, Z, info)                     # asserts H c = 0 itself
    assert cw.shape == (5, code.N) and np.array_equal(cw[:, :code.K], info)
    assert not oracle.encode_dense(code.shifts, Z, np.zeros((1, code.K), dtype=np.uint8)).any()
    with pytest.raises(ValueError):
        SystematicEncoder(QCCode(np.array([[0, 0, -1], [0, -1, 0]]), 2))      # every row owns a degree-1 column: no core block


def test_sorted_pack_plan_of_the_neighbour_tables():
    """models/layers.py:sort_plan (host side of the sorted-pack tables the layer / neural-decoder kernels read):
    row t holds exactly the valid neighbours of edge perm[t] in the caller's order, counts are right and
    non-increasing, perm is a permutation; padding anywhere in the caller's rows is tolerated."""
    from ldpc_b200.models.layers import sort_plan
    g = load_golden("mapping_layers")
    for name in ("z4_check", "z4_var", "toy_check", "toy_var"):
        src = torch.from_numpy(g[name].astype(np.int64))
        if name == "z4_check":
            src = torch.flip(src, dims=[1])                          # padding first
        rows, perm, cnt = sort_plan(src)
        E, K = src.shape
        assert sorted(perm.tolist()) == list(range(E))
        c = cnt.to(torch.int64)
        assert bool((c[:-1] >= c[1:]).all())
        for t in range(E):
            want = [int(v) for v in src[perm[t]] if v >= 0]
            assert rows[t, :len(want)].tolist() == want and int(c[t]) == len(want)
            assert bool((rows[t, len(want):] == 0).all())
        rows_id, perm_id, cnt_id = sort_plan(src, sort=False)
        assert perm_id.tolist() == list(range(E)) and int(cnt_id.to(torch.int64).sum()) == int((src >= 0).sum())


def test_base_matrix_shifts_are_reduced_mod_z_like_the_reference_lift():
    """the reference's expand_base_matrix rolls by the raw shift, i.e. mod Z (ldpc_utils.py:121-123), so
    NR_2_0_32.txt lifted with --lifting_factor 16 works upstream; the (base_graph, Z) constructors must accept it too
    and describe the same dense H.  The raw constructor keeps the strict range check."""
    base32 = QCCode.nr_2_0(32).base_matrix()                       # shifts up to 31
    code = QCCode.from_base_matrix(base32, 16)
    assert code.Z == 16 and int(code.shifts.max()) < 16
    assert np.array_equal(code.shifts, QCCode.nr_2_0(16).shifts)
    assert torch.equal(code.dense(), expand_base_matrix(base32, 16))
    assert np.array_equal(as_code(base_graph=base32, Z=16).shifts, code.shifts)
    with pytest.raises(ValueError):
        QCCode(base32.numpy(), 16)


def test_gnn_rejects_other_hidden_widths_in_the_constructor():
    from ldpc_b200.models import MessageGNNDecoder
    with pytest.raises(ValueError, match="hidden_dim"):
        MessageGNNDecoder(788, 5, hidden_dim=32)


def test_bench_roofline_counts_match_the_kernel_sources():
    """bench.py computes roofline.frac from per-codeword counts measured with ncu (tools/ncu_to_profile.py).  The count
    file records the SHA-256 of the kernel sources it was captured on: if the kernel changed since, the counts are
    stale and this test says so (re-capture: tools/ncu_to_profile.py, DESIGN.md 'Measurement')."""
    import hashlib
    import json
    import bench
    for workload, rel in bench.COUNTS_FILES.items():
        path = os.path.join(ROOT, rel)
        if not os.path.exists(path):
            assert workload != "minsum", "the headline workload needs its count file"
            continue
        c = json.load(open(path))
        assert c["inst_per_cw"] > 0 and c["dram_bytes_per_cw"] > 0 and c["codewords"] > 0
        assert c["source_sha256"], "count file does not say which sources it was measured on"
        for src, sha in c["source_sha256"].items():
            now = hashlib.sha256(open(os.path.join(ROOT, src), "rb").read()).hexdigest()
            assert now == sha, f"{src} changed after the ncu capture behind {rel}: re-run the capture"
        k = bench.kernel_counts(workload)
        assert k["inst_per_cw"] == c["inst_per_cw"] and k["file"] == rel


def test_rate_match_tables_hand_computed_case_and_oracle():
    """TS 38.212 5.4.2 on BG2 Z=2 (N = 104, d has 100 bits, K = 20): payload 16 bits -> fillers c[16..19] = d[12..15].
    rv 0, E = 10, Qm = 2: e = d[0..9] and f = [e0 e5 e1 e6 e2 e7 e3 e8 e4 e9]; full-codeword index = d index + 2Z."""
    from ldpc_b200.utils import rate_match_tables
    from oracle import oracle
    code = QCCode.nr_2_0(2)
    sel, kind = rate_match_tables(code, 10, payload_bits=16, rv=0, Qm=2)
    assert sel.tolist() == [4 + x for x in (0, 5, 1, 6, 2, 7, 3, 8, 4, 9)]
    assert kind[:4].tolist() == [1] * 4 and kind[16:20].tolist() == [2] * 4 and int((kind == 0).sum()) == 104 - 8
    # E = 20 crosses the filler block: d[12..15] are skipped
    sel, _ = rate_match_tables(code, 20, payload_bits=16, rv=0, Qm=1)
    assert sel.tolist() == [4 + x for x in list(range(12)) + list(range(16, 24))]
    # rv 2: k0 = floor(25 * 100 / 100) * 2 = 50; wrap-around with repetition when E exceeds the 96 usable bits
    sel, _ = rate_match_tables(code, 100, payload_bits=16, rv=2, Qm=1)
    usable = [x for x in range(100) if not 12 <= x < 16]
    start = usable.index(50)
    want = [usable[(start + t) % 96] + 4 for t in range(100)]
    assert sel.tolist() == want and sel[96] == sel[0]
    # against the specification's loops (oracle), bit by bit, over rv / Qm / limited buffers / both lifting sizes
    rng = np.random.default_rng(0)
    for Z, Kp, E, rv, Qm, Ncb in ((2, 16, 10, 0, 2, None), (2, 20, 96, 1, 4, None), (2, 13, 300, 3, 6, 80), (4, 33, 128, 2, 8, None),
                                  (32, 300, 2048, 0, 2, None), (32, 320, 1600, 3, 1, 1200)):
        c = QCCode.nr_2_0(Z)
        sel, kind = rate_match_tables(c, E, payload_bits=Kp, rv=rv, Qm=Qm, Ncb=Ncb)
        cw = rng.integers(0, 2, c.N).astype(np.uint8)
        assert np.array_equal(cw[sel], oracle.rate_match_38212(Z, c.cols, c.K, Kp, cw, E, rv, Qm, Ncb))
        assert not (kind[sel] != 0).any()                 # punctured and filler positions are never transmitted
    with pytest.raises(ValueError):
        rate_match_tables(code, 11, payload_bits=16, Qm=2)
    with pytest.raises(ValueError):
        rate_match_tables(code, 10, payload_bits=3)


def test_neural_qc_schedule_tables_cover_the_base_graph_exactly_once():
    """csrc/nq_tables.h (generated by csrc/gen_tables.py:emit_nq) drives the rolled QC neural kernels: every base row, core
    column, degree-1 cell and (member, cell) of the backward's flattened list must be owned exactly once, the row / column
    classes must match the shapes the kernel bodies are compiled for, and the packed chunk table must decode to the
    variable-major edge numbering of create_LLR_mapping."""
    text = open(os.path.join(ROOT, "ldpc-neuralnetwork-decoder_b200", "csrc", "nq_tables.h")).read()

    def arr(name):
        m = re.search(r"\b%s(\[[^=]*)= \{([^}]*)\}" % name, text)
        assert m, name
        dims = [int(x) for x in re.findall(r"\[(\d+)\]", m.group(1))]
        return np.array([int(x) for x in m.group(2).split(",")]).reshape(dims)
    code = QCCode.nr_2_0(32)
    deg_c = (code.shifts >= 0).sum(axis=0)
    core = [j for j in range(code.cols) if deg_c[j] > 1]
    row_meta, row_ext = arr("row_meta"), arr("row_ext")
    sched_rows, row_ptr = arr("sched_rows"), arr("sched_row_ptr")
    classes = [(2, 1), (3, 1), (4, 1), (5, 1), (8, 0), (10, 0)]
    seen_rows = []
    for m in range(4):
        for c, (nc, ne) in enumerate(classes):
            for t in range(row_ptr[m][c], row_ptr[m][c + 1]):
                i = int(sched_rows[m][t])
                seen_rows.append(i)
                cols = np.nonzero(code.shifts[i] >= 0)[0]
                assert sum(deg_c[j] > 1 for j in cols) == nc and sum(deg_c[j] == 1 for j in cols) == ne
                assert (row_ext[i] != 255) == bool(ne)
    assert sorted(seen_rows) == list(range(code.rows))
    # every core cell appears in exactly one row's metadata with the right shift
    vm0 = np.concatenate([[0], np.cumsum(deg_c)])[:-1]
    cells = []
    for i in range(code.rows):
        k_in_col = {j: int((code.shifts[:i, j] >= 0).sum()) for j in range(code.cols)}
        metas = [int(x) for x in row_meta[i][:sum(1 for j in np.nonzero(code.shifts[i] >= 0)[0] if deg_c[j] > 1)]]
        want = [(vm0[j] + k_in_col[j]) | (int(code.shifts[i, j]) << 8) for j in np.nonzero(code.shifts[i] >= 0)[0] if deg_c[j] > 1]
        assert metas == [int(w) for w in want]
        cells += [w & 0xff for w in metas]
    assert sorted(cells) == list(range(int(deg_c[core].sum())))
    sched_cols, col_ptr, col_b0, col_d = arr("sched_cols"), arr("sched_col_ptr"), arr("col_b0"), arr("col_d")
    cmax = [6, 8, 10, 13, 16, 23]
    seen_cols = []
    for m in range(4):
        for c in range(6):
            for t in range(col_ptr[m][c], col_ptr[m][c + 1]):
                j = int(sched_cols[m][t])
                seen_cols.append(j)
                assert (cmax[c - 1] if c else 0) < col_d[j] <= cmax[c] and col_b0[j] == vm0[j] and col_d[j] == deg_c[j]
    assert sorted(seen_cols) == core
    sched_ext, ext_cnt = arr("sched_ext"), arr("sched_ext_cnt")
    assert sorted(int(x) for m in range(4) for x in sched_ext[m][:ext_cnt[m]]) == list(range(code.cols - len(core)))
    sc, scnt = arr("sched_cells"), arr("sched_cell_cnt")
    assert sorted(int(x) for m in range(4) for x in sc[m][:scnt[m]] if x != 255) == list(range(int(deg_c[core].sum())))
    chunk = arr("chunk_meta")
    for m, w in enumerate(chunk):
        D, d, inv = int(w) & 0xff, (int(w) >> 8) & 0x1f, int(w) >> 13
        j = int(np.searchsorted(vm0, m, side="right") - 1)
        assert D == vm0[j] and d == deg_c[j]
        assert all(((off * inv) >> 16) == off // d for off in range(32 * d))


@pytest.mark.parametrize("Z", [33, 36, 48, 64, 96, 384])
def test_lifting_factors_above_32_are_rewritten_as_an_equivalent_code(Z):
    """The reference lifts with any --lifting_factor (main.py:38, expand_base_matrix).  QCCode.from_base_matrix(..., allow_split=True)
    holds Z > 32 as an equivalent QC code with Zs = the largest divisor of Z <= 32 and renumbered variables / checks: its expanded
    H must be the reference's expanded H with rows and columns permuted by exactly that renumbering, and cell order inside a
    base row / column must be preserved (so the decoders keep the reference's operation order)."""
    import torch
    from ldpc_b200.utils import QCCode
    from ldpc_b200.utils.ldpc_utils import expand_base_matrix
    rng = np.random.default_rng(Z)
    support = QCCode.nr_2_0(32).shifts >= 0
    base = np.where(support, rng.integers(0, 4 * Z, size=support.shape), -1)          # raw shifts beyond Z: taken mod Z as upstream
    code = QCCode.from_base_matrix(base, Z, allow_split=True)
    assert code.lift_Z == Z and code.Z <= 32 and Z % code.Z == 0 and code.N == 52 * Z and code.M == 42 * Z
    m = Z // code.Z
    H_nat = expand_base_matrix(torch.from_numpy(base.astype(np.float32)), Z).numpy()
    H_eng = expand_base_matrix(torch.from_numpy(code.shifts.astype(np.float32)), code.Z).numpy()
    ib, a = np.divmod(np.arange(code.M), code.Z)
    i, b = np.divmod(ib, m)
    chk_old = i * Z + m * a + b
    assert np.array_equal(H_eng, H_nat[np.ix_(chk_old, code.var_old_of_new)])
    assert np.array_equal(code.var_new_of_old[code.var_old_of_new], np.arange(code.N))
    assert np.array_equal(code.dense().numpy(), H_nat)                                # dense() speaks the caller's numbering
    # base row / column majority of the renumbering: the engine index of a variable grows with its base column, of a check with its base row
    assert np.array_equal(code.var_old_of_new // Z, np.arange(code.N) // Z)
    with pytest.raises(ValueError):
        QCCode.from_base_matrix(base, Z)                                              # other consumers: Z <= 32 only
