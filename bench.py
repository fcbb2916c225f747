#!/usr/bin/env python
"""bench.py -- decoded information Gbit/s of the LDPC hot path on B200.

Workload (BASELINE.json metric): 5G NR BG2, Z=32 (N=1664, K=320 info
bits), scaled min-sum alpha=0.75, 10 flooding iterations, fixed iteration count, all-zero
codeword over BPSK-AWGN at snr_db=-2 (Eb/N0 ~ 2.15 dB), B codewords per GPU per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload minsum|bp]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the CPU oracle port on the host cores

One JSON line on stdout (rank 0).  `value`: LLRs resident in HBM, CUDA-event time of K decode
launches, max over ranks.  `e2e`: the same through ldpc_decode_host with pinned HOST buffers
(H2D + decode + D2H inside the timed region).  `roofline`: see DESIGN.md "Measurement".
Only this file's cpu_baseline / --impl reference legs execute oracle/.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_INFO, N_BITS, BASE_EDGES, Z = 320, 1664, 197, 32
ITERS, ALPHA, SNR_DB = 10, 0.75, -2.0
ISSUE_SLOTS_PER_EDGE_ITER = 18          # 14 ALU + 4 MIO lane-ops per edge-iteration
SM_COUNT = 148
# Binding resource of the min-sum kernel (ncu capture r1e, profiles/r1_ncu_minsum_fast.md): the
# instruction ISSUE rate -- sm__issue_active 74.6 % is the highest utilisation (ALU pipe 70 %, shared/
# shuffle pipe 48 %, FMA pipe 25 %, TMEM 3 %, DRAM 3 %).  Work per codeword = 19 075 warp-instructions
# (smsp__inst_executed.sum / codewords): 10 iterations x 1 870 (FMNMX/FMNMX3 374, FADD 318, LOP3 292,
# SHFL 290, FSETP 159, IMAD 190, FMUL 84, LDTM+STTM 80, LDS 14, ...) + 375 per codeword for load,
# message zeroing and output.  Peak: 4 warp-instructions per clock per SM.
ISSUE_INSTR_PER_CW = 19075
# earlier binding units, still reported: ALU pipe (825 ALU-pipe instructions per codeword-iteration at
# 2 per clock per SM) and, before the messages moved to Tensor Memory, the shared-memory/shuffle pipe
# (622 wavefronts per codeword-iteration at 1 per clock per SM).
ALU_OPS_PER_CW_ITER = 825
SMEM_WAVEFRONTS_PER_CW_ITER = 159 * 2 + 145 * 2 + 14
# DRAM bytes per codeword of the decode kernel measured by `ncu --set full` (profiles/r1_ncu_minsum_fast.md,
# capture r1e: dram__bytes_read.sum + dram__bytes_write.sum = 1.8010 GB for 262144 codewords)
NCU_DRAM_BYTES_PER_CW = 6870.0


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


ALL_CPUS = os.sched_getaffinity(0)          # before any rank-local binding (bind_to_gpu_numa_node)


def cpu_baseline(workload, seconds=12.0):
    """The oracle port (faithful O(d^2) restatement of the reference loops, OpenMP over
    codewords) on a bounded sample of the same workload."""
    import numpy as np
    from oracle import oracle
    import ldpc_b200
    from ldpc_b200.utils import QCCode
    code = QCCode.nr_2_0(Z)
    os.sched_setaffinity(0, ALL_CPUS)                    # undo the GPU-local binding of the e2e leg: the CPU arm gets every core
    threads = os.cpu_count() or oracle.num_threads()     # torchrun exports OMP_NUM_THREADS=1: ask for all cores explicitly
    probe = 64 * threads
    llr = oracle.awgn_llr(None, probe, code.N, SNR_DB, seed=1)
    t0 = time.perf_counter()
    oracle.decode(code.shifts, Z, llr, ITERS, workload, ALPHA, threads=threads)
    rate = probe / (time.perf_counter() - t0)
    sample = int(max(probe, min(rate * seconds, 1 << 20)))
    llr = oracle.awgn_llr(None, sample, code.N, SNR_DB, seed=2)
    t0 = time.perf_counter()
    oracle.decode(code.shifts, Z, llr, ITERS, workload, ALPHA, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": sample * K_INFO / dt / 1e9, "unit": "Gbit/s", "cores": threads, "kind": "port",
            "sample": f"{sample} codewords of the same workload (BG2 Z=32, {ITERS} it, snr_db {SNR_DB}), "
                      f"oracle/ldpc_oracle.c with {threads} OpenMP threads, {dt:.1f} s",
            "codewords_per_s": sample / dt, "sample_codewords": sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    from oracle import oracle
    oracle.build()
    per_step = max(4.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals, step_ms = [], []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(args.workload, seconds=per_step)
        if i >= args.warmup:
            vals.append(base["value"])
            cw = base["value"] * 1e9 / K_INFO                          # codewords/s of this step's sample
            step_ms.append(base["sample_codewords"] / cw * 1e3)
    v = sum(vals) / len(vals)
    base["value"] = v
    out = {"impl": "reference", "metric": "decoded info Gbit/s, 5G BG2 Z=32 10 iters", "value": v, "unit": "Gbit/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sum(step_ms) / len(step_ms),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": dict(workload_config(args, base["sample_codewords"]),
                          note="each step is a bounded sample of the workload on the host cores (rank 0 only)"),
           "cpu_baseline": base,
           "e2e": {"value": v, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def algo_bytes_fn(B):
    return B * (N_BITS * 4 + ((N_BITS + 31) // 32) * 4)


def workload_config(args, B):
    return {"workload": f"{args.workload}_bg2_z32_it{ITERS}_alpha{ALPHA}_snr{SNR_DB}dB_fixed_iters",
            "codewords_per_gpu_per_step": B, "N": N_BITS, "K": K_INFO, "iters": ITERS,
            "l2_policy": "inputs (6.6 KB/codeword x batch) far larger than the 126 MB L2",
            "sharding": "independent codeword ranges per GPU, counters all-reduced once"}


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function), so
    that the pinned staging buffers of the e2e path are first-touched on the GPU's NUMA node and the H2D copies do
    not cross the socket interconnect.  Best effort: returns the cpulist used, or None."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        cpulist = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpulist
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="minsum", choices=["minsum", "bp"])
    ap.add_argument("--batch", type=int, default=1 << 20)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="size of the cpu_baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    import ldpc_b200
    from ldpc_b200 import _native
    from ldpc_b200.utils import QCCode

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _native.lib()
    code = QCCode.nr_2_0(Z)
    h = code.handle(dev)
    B = args.batch
    algo = _native.ALGO_MINSUM if args.workload == "minsum" else _native.ALGO_BP
    NW = (code.N + 31) // 32

    # synthetic inputs, generated on the device by the engine's own channel (Philox), resident in HBM
    llr = torch.empty((B, code.N), dtype=torch.float32, device=dev)
    _native.check(L.ldpc_awgn_llr(None, B, code.N, SNR_DB, 1234, rank * B, _native.ptr(llr), _native.stream_ptr(dev)))
    hard = torch.empty((B, NW), dtype=torch.int32, device=dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    st = _native.stream_ptr(dev)

    def step():
        if algo == _native.ALGO_MINSUM:
            rc = L.ldpc_minsum_decode(h, _native.ptr(llr), B, ITERS, ALPHA, 0, 0, None, _native.ptr(hard),
                                      _native.HARD_PACKED, None, None, None, 0, st)
        else:
            rc = L.ldpc_bp_decode(h, _native.ptr(llr), B, ITERS, 0, _native.PATH_FAST, None, _native.ptr(hard),
                                  _native.HARD_PACKED, None, None, None, 0, st)
        _native.check(rc)
        _native.check(L.ldpc_count_errors(_native.ptr(hard), _native.HARD_PACKED, None, B, code.N, _native.ptr(counters), st))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.ldpc_launch_count()
    counters.zero_()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record()
    for i in range(args.steps):
        kev[i][0].record()
        step()
        kev[i][1].record()
    if world > 1:
        dist.all_reduce(counters)           # the path's only collective: error counters, once per sweep point
    ev[1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = L.ldpc_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev[0].elapsed_time(ev[1])
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    step_ms = sorted(a.elapsed_time(b) for a, b in kev)
    total_cw = B * world * args.steps
    value = total_cw * K_INFO / (elapsed_ms * 1e-3) / 1e9

    # ---- e2e: host buffers through the C ABI (H2D + decode + D2H in the timed region) ----
    Be = min(B, 1 << 19)
    llr_host = torch.empty((Be, code.N), dtype=torch.float32).pin_memory()
    llr_host.copy_(llr[:Be])
    hard_host = torch.empty((Be, NW), dtype=torch.int32).pin_memory()

    def e2e_step():
        _native.check(L.ldpc_decode_host(h, algo, _native.ptr(llr_host), Be, ITERS, ALPHA,
                                         _native.PATH_FAST if algo == _native.ALGO_BP else _native.PATH_AUTO, None,
                                         _native.ptr(hard_host), _native.HARD_PACKED, 1 << 15))
    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = Be * world * args.e2e_steps * K_INFO / float(te.item()) / 1e9

    # the same call with 8-bit LLRs (what a demapper produces): 4x fewer PCIe bytes, same fp32 arithmetic on q * scale
    q_host = torch.empty((Be, code.N), dtype=torch.int8).pin_memory()
    q_host.copy_(torch.clamp(torch.round(llr[:Be] * 4.0), -127, 127).to(torch.int8))

    def e2e_q_step():
        _native.check(L.ldpc_decode_host_q(h, algo, _native.ptr(q_host), _native.LLR_I8, 0.25, Be, ITERS, ALPHA,
                                           _native.PATH_FAST if algo == _native.ALGO_BP else _native.PATH_AUTO, None,
                                           _native.ptr(hard_host), _native.HARD_PACKED, 1 << 15))
    e2e_q_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_q_step()
    torch.cuda.synchronize()
    tq = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tq, op=dist.ReduceOp.MAX)
    e2e_q_value = Be * world * args.e2e_steps * K_INFO / float(tq.item()) / 1e9
    frame_errors = int(counters[1].item())
    frames = int(counters[2].item())

    if rank == 0:
        peaks, peak_src = measured_peaks()
        kernel_ms = step_ms[len(step_ms) // 2]                    # median launch (decode + tiny count kernel)
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        achieved = B * ISSUE_INSTR_PER_CW / (kernel_ms * 1e-3) / 1e9               # G warp-instructions / s, one GPU
        peak = SM_COUNT * 4 * sm_max * 1e6 / 1e9                                   # 4 warp-instr / clk / SM
        alu_ach = B * ALU_OPS_PER_CW_ITER * ITERS / (kernel_ms * 1e-3) / 1e9
        alu_peak = SM_COUNT * 2 * sm_max * 1e6 / 1e9
        smem_ach = B * SMEM_WAVEFRONTS_PER_CW_ITER * ITERS / (kernel_ms * 1e-3) / 1e9
        smem_peak = SM_COUNT * sm_max * 1e6 / 1e9
        edge_iters = B * BASE_EDGES * Z * ITERS
        issue_ach = edge_iters / (kernel_ms * 1e-3) / 1e9
        issue_peak = SM_COUNT * 4 * 32 * sm_max * 1e6 / ISSUE_SLOTS_PER_EDGE_ITER / 1e9
        algo_bytes = B * (code.N * 4 + NW * 4)
        hbm_ach = algo_bytes / (kernel_ms * 1e-3) / 1e9
        out = {
            "metric": "decoded info Gbit/s, 5G BG2 Z=32 10 iters", "value": value, "unit": "Gbit/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, B),
            "codewords_per_s": total_cw / (elapsed_ms * 1e-3),
            "fer": {"frame_errors": frame_errors, "frames": frames},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Gbit/s", "h2d_bytes_per_step": Be * code.N * 4,
                    "d2h_bytes_per_step": Be * NW * 4, "codewords_per_step": Be, "steps": args.e2e_steps,
                    "api": "ldpc_decode_host (pinned host fp32 LLRs -> packed hard bits)", "host_cpus_rank0": numa},
            "e2e_int8_llr": {"value": e2e_q_value, "unit": "Gbit/s", "h2d_bytes_per_step": Be * code.N,
                             "d2h_bytes_per_step": Be * NW * 4,
                             "api": "ldpc_decode_host_q (pinned host int8 LLRs, scale 0.25 -> packed hard bits)"},
            "roofline": {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "Gwarp-instr/s",
                         "frac": achieved / peak,
                         "traffic": (NCU_DRAM_BYTES_PER_CW * B if args.workload == "minsum" else None),
                         "traffic_unit": "bytes of DRAM per launch (ncu), algorithmic = %d" % algo_bytes_fn(B),
                         "model": f"{ISSUE_INSTR_PER_CW} warp-instructions per codeword (ncu, 10 x 1870 + 375; DESIGN.md 3.1), "
                                  f"148 SMs x 4 warp-instr/clk at {sm_max:.0f} MHz ({peak_src} max clock)",
                         "frac_at_measured_clock": (achieved / (peak * clocks["sm_mhz"] / sm_max)) if clocks and clocks.get("sm_mhz") else None,
                         "kernel_ms": kernel_ms,
                         "alu_pipe_model": {"achieved": alu_ach, "peak": alu_peak, "unit": "Gwarp-instr/s", "frac": alu_ach / alu_peak,
                                            "note": "825 ALU-pipe instructions per codeword-iteration at 2/clk/SM"},
                         "smem_pipe_model": {"achieved": smem_ach, "peak": smem_peak, "unit": "Gwavefront/s", "frac": smem_ach / smem_peak,
                                             "note": "622 shared-memory/shuffle wavefronts per codeword-iteration had the messages "
                                                     "stayed in shared memory (the bound of the pre-TMEM kernel)"},
                         "survey_issue_model": {"achieved": issue_ach, "peak": issue_peak, "unit": "Gedge-iter/s",
                                                "frac": issue_ach / issue_peak,
                                                "note": "SURVEY 8d estimate of 18 issue slots per edge-iteration; the kernel "
                                                        "issues 9.4, so this fraction can exceed 1"},
                         "hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": hbm_ach / peaks["hbm_gbs"], "algorithmic_bytes_per_codeword": code.N * 4 + NW * 4,
                                 "peak_source": peak_src}},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.workload, seconds=args.cpu_seconds)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
