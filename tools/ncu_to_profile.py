#!/usr/bin/env python
"""Turn an ncu report (.ncu-rep, or the CSV of `ncu --page raw --csv`) into the small, tracked
files under profiles/ that bench.py and the judge read.

  python tools/ncu_to_profile.py gpurun_out/minsum_r2.ncu-rep --codewords 262144 \
         --kernel 'decode_fast_kernel<BG2Z32, 8, 0, 0>' \
         --sources ldpc-neuralnetwork-decoder_b200/csrc/decode_fast_kernel.cuh ... \
         --out profiles/r2_minsum_counts.json [--md profiles/r2_ncu_minsum.md]

The JSON carries the per-codeword work counts bench.py's roofline is computed from
(`inst_per_cw`, `dram_bytes_per_cw`), the utilisations that say which unit binds, and the
SHA-256 of the kernel sources the capture was taken from: tests/test_host_logic.py fails when
the sources have changed since (a stale count would silently falsify `roofline.frac`).
"""
import argparse
import csv
import hashlib
import io
import json
import os
import subprocess
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}           # bytes; milliseconds

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_read.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def sha256(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def source_hashes(paths):
    return {os.path.relpath(os.path.abspath(p), ROOT): sha256(p) for p in paths}


def read_rows(path):
    if path.endswith(".csv"):
        text = open(path).read()
    else:
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    # ncu prints "==PROF==" banner lines before the CSV when reading a log; keep from the header on
    start = text.find('"ID"')
    rows = list(csv.reader(io.StringIO(text[start if start >= 0 else 0:])))
    head, units, body = rows[0], rows[1], rows[2:]
    return head, units, body


def pick(head, units, body, kernel, index):
    k = head.index("Kernel Name")
    hits = [r for r in body if (kernel is None or kernel in r[k])]
    if not hits:
        raise SystemExit(f"no launch matching {kernel!r}; kernels: {sorted(set(r[k] for r in body))}")
    row = hits[index]
    out = {"kernel": row[k]}
    for i, name in enumerate(head):
        if not name or i >= len(row) or row[i] == "":
            continue
        try:
            v = float(row[i].replace(",", ""))
        except ValueError:
            continue
        out[name] = v * UNIT_SCALE.get(units[i], 1.0) if (units[i] in UNIT_SCALE and ("bytes" in name or "time_duration" in name)) else v
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default=None, help="substring of the kernel name (default: first launch)")
    ap.add_argument("--index", type=int, default=0, help="which matching launch")
    ap.add_argument("--codewords", type=int, required=True, help="codewords the captured launch processed")
    ap.add_argument("--sources", nargs="*", default=[])
    ap.add_argument("--out", required=True)
    ap.add_argument("--md", default=None)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    head, units, body = read_rows(a.report)
    m = pick(head, units, body, a.kernel, a.index)
    cw = a.codewords
    res = {
        "kernel": m["kernel"], "codewords": cw, "ncu_report": os.path.basename(a.report), "note": a.note,
        "generated": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
        "inst_per_cw": m["smsp__inst_executed.sum"] / cw,
        "dram_bytes_per_cw": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / cw,
        "issue_active": m.get("sm__issue_active.avg.pct_of_peak_sustained_elapsed", 0.0) / 100.0,
        "time_ms_under_ncu": m.get("gpu__time_duration.sum"),
        "registers": m.get("launch__registers_per_thread"),
        "metrics": {k: m[k] for k in KEEP if k in m},
        "stalls_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: round(v, 4) for k, v in m.items()
                             if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")},
        "source_sha256": source_hashes(a.sources),
    }
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    if a.md:
        with open(a.md, "w") as f:
            f.write(f"# ncu --set full: `{m['kernel']}`\n\n{a.note}\n\nreport `{os.path.basename(a.report)}`, {cw} codewords in the captured launch; "
                    "per-launch values under ncu are cold-cache and serialised (compare shares and utilisations, not absolute times).\n\n"
                    "| metric | value |\n|---|---|\n")
            f.write(f"| warp-instructions per codeword | {res['inst_per_cw']:.1f} |\n| DRAM bytes per codeword | {res['dram_bytes_per_cw']:.1f} |\n")
            for k, v in res["metrics"].items():
                f.write(f"| `{k}` | {v:.6g} |\n")
            f.write("\n## warp stall reasons (per issue)\n\n| reason | ratio |\n|---|---|\n")
            for k, v in sorted(res["stalls_per_issue"].items()):
                f.write(f"| {k} | {v:.3f} |\n")
    print(json.dumps({k: res[k] for k in ("kernel", "inst_per_cw", "dram_bytes_per_cw", "issue_active", "registers")}))


if __name__ == "__main__":
    main()
