#!/usr/bin/env python3
"""Benchmark of the hot path: (k-1)-mer counting -> filter -> de Bruijn graph build -> CSR.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c1] [--impl reference]

One "step" = one full pass of the path over the workload's synthetic reads.
  value   occurrences/s with the packed reads already resident in HBM (CUDA events)
  e2e     the same through the host-buffer entry: ASCII reads in pinned host memory -> H2D ->
          pack -> count -> filter -> build -> CSR -> D2H of the CSR arrays, all timed
  roofline  algorithmic bytes of the dominant kernel / its CUDA-event duration vs measured HBM peak
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, copied by oracle/make_ref.py) through its own CLI on a bounded sample
`--impl reference` times that reference run alone, K timed steps after W warm-up steps (the pure-Python port of
oracle/py_oracle.py stands in only when oracle/_ref/ is missing).
Workloads are synthetic stand-ins of BASELINE.json's configs ("Measurement").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "genome-assembler_b200")
for path in (ROOT, PKG):
    if path not in sys.path:
        sys.path.insert(0, path)

# name: (genome bp, reads (pairs when paired), read length, paired, k, F, description)
WORKLOADS = {
    "c1": (112031, 19000, 100, True, 28, 3, "C1-shaped: 112,031 bp genome, 19,000 read-pairs x 100 bp, k=28, F=3"),
    "c2": (2872769, 861831, 100, False, 31, 3, "C2-shaped (S. aureus size): 2,872,769 bp genome, 861,831 reads x 100 bp (30x), k=31, F=3"),
    "c3": (4641652, 700000, 100, True, 29, 3, "C3-shaped (E. coli size): 4,641,652 bp genome, 700,000 read-pairs x 100 bp, k=29, F=3"),
    "c4": (50000000, 100000000, 150, False, 31, 3, "C4: 50 Mbp genome, 100M reads x 150 bp, 1% substitutions, k=31, F=3"),
}
SEED = 2026
# algorithmic bytes per (k-1)-mer occurrence, 64-bit keys (SURVEY 8d / DESIGN.md)
B_TOTAL = 40.0
B_TOTAL_SKETCH = 100.0     # -c route: + d = 10 rows x (u16 RMW on update 4 B + 2 B read on estimate) per occurrence
# per kernel: packed bases (0.35) + what the kernel itself must touch per occurrence
B_KERNEL = {"count_full": 16.35,   # v1 full table: key probe 8 + count RMW 8
            "prefilter": 1.35,     # sketch cell read-modify-write (4-bit cell, counted as 1 B)
            "count": 16.85,        # cell read 0.5 + key probe 8 + count RMW 8
            "build": 22.35,        # key probe 8 + id/epoch word 4 + edge + node stamp RMW 10
            "build_tail": 22.35,   # same pass, the reads after the prefix
            # bucketed path (csrc/ga_superkmer.cu): the two base reads of the reference's two loops
            # happen in the scatter, every probe / count / stamp access in the bucket kernel
            # -c route: the sketch is poured from the exact table (one update per DISTINCT window and row) and asked
            # once per distinct window: SURVEY 8d charges 40 B (update) + 20 B (estimate) per OCCURRENCE
            "sketch_update": 40.0, "select_solid": 20.0,
            "sk_scatter1": 0.7, "sk_scatter2": 0.0, "sk_push": 0.0,
            "sk_bucket": 38.0}     # count probe + RMW 16, build probe + count read 12, stamp RMW 10


def measured_traffic(workload, kernel, launch_occ):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of this workload
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per occurrence), scaled to the
    occurrences one launch of this run covers; None when the kernel was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    per_occ = json.load(open(path)).get(workload, {}).get(kernel)
    return None if per_occ is None else per_occ * launch_occ


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML from a thread (cheap
    per query); spawning `nvidia-smi -lms` next to a 25 ms timed region stalls the driver for
    longer than the region itself, so it is only the fallback."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.nvml = index, [], threading.Event(), None
        self.period = float(os.environ.get("GA_CLOCK_PERIOD", "0.1"))   # seconds between NVML samples
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def start(self):
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()
        return self

    def _pump(self):
        if self.nvml is None:
            return
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((sm, reasons))
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def stop(self):
        self.stop_flag.set()
        if self.nvml is None:
            return self._smi_once()
        self.thread.join(timeout=2)
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(r[0] for r in self.rows)
        seen = set()
        for _, mask in self.rows:
            seen |= {name for name, bit in bits.items() if mask & bit}
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(seen),
                "samples": len(sm), "source": "nvml, %d ms period, during the timed region" % int(self.period * 1e3)}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index),
                                  "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
            cells = [c.strip() for c in out.strip().split(",")]
            return {"sm_mhz": float(cells[0]), "sm_max_mhz": float(cells[1]), "reasons": [cells[2]], "samples": 1,
                    "source": "nvidia-smi once after the timed region (NVML python module missing)"}
        except Exception as exc:          # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable: %s" % exc], "samples": 0}


def workload_occ(n_reads, read_len, paired, k):
    return n_reads * (2 if paired else 1) * (read_len - k + 2)


# ------------------------------------------------------------------------------------ CPU arm
REF_DIR = os.path.join(ROOT, "oracle", "_ref")          # the unmodified reference (oracle/make_ref.py)


def cpu_port_rate(reads, k, F, paired):
    """Occurrences/s of the oracle's pure-Python port (count + build), one core.  Only used when the
    reference itself is not available (oracle/_ref/ missing)."""
    from oracle import py_oracle as po
    t0 = time.perf_counter()
    tally = (po.count_paired if paired else po.count_unpaired)(k, reads)
    (po.build_paired if paired else po.build_unpaired)(tally, reads, k, F)
    dt = time.perf_counter() - t0
    occ = sum(tally.values())
    return occ / dt, dt, occ


def host_sample_reads(genome_size, n_reads, sample, read_len, paired):
    """A scaled-down instance of the workload as Python strings: `sample` reads over a genome
    shrunk by the same factor, so coverage (and with it the share of k-mers that pass the filter
    and reach the build) is that of the full workload.  Same generator arithmetic as the device
    (oracle/readgen.py)."""
    from oracle import readgen
    small_genome = max(4 * read_len, int(round(genome_size * (sample / float(n_reads)))))
    genome = readgen.splitmix_genome_codes(small_genome, SEED)
    mates = sample * (2 if paired else 1)
    codes = readgen.splitmix_reads_codes(genome, read_len, 0, mates, SEED, 100, paired, 125)
    strings = readgen.codes_to_strings(codes)
    return list(zip(strings[0::2], strings[1::2])) if paired else strings


def write_reference_input(path, reads, paired):
    """The stdin format of the reference's IOHandler.read_input (assemble.py:40-71)."""
    with open(path, "w") as fh:
        fh.write("%d\n" % len(reads))
        if paired:
            fh.write("".join("%s|%s|125\n" % pair for pair in reads))
        else:
            fh.write("\n".join(reads) + "\n")


def reference_cli_step(path, k, F, sketch=False):
    """One run of the UNMODIFIED reference through its own CLI and stock code path
    (`python assemble.py --stdout -t -k K -f F [-c] < reads`, assemble.py:172-194); stage times from its own
    `-t` prints (debug_graph.py:20-85).  Returns (count + [sketch] + build seconds, wall seconds)."""
    cmd = [sys.executable, os.path.join(REF_DIR, "assemble.py"), "--stdout", "-t", "-k", str(k), "-f", str(F)]
    if sketch:
        cmd.append("-c")
    t0 = time.perf_counter()
    with open(path) as fh:
        proc = subprocess.run(cmd, stdin=fh, capture_output=True, text=True, cwd=REF_DIR)
    wall = time.perf_counter() - t0
    if proc.returncode != 0:
        raise RuntimeError("reference CLI failed: " + proc.stderr[-400:])
    stamp = {}
    for line in proc.stdout.split("\n"):
        for key in ("STARTING TO COUNT KMERS", "FINISHED BUILDING GRAPH"):
            if key in line:
                stamp[key] = float(line.split("T =")[1].split()[0].rstrip("-"))
    return stamp["FINISHED BUILDING GRAPH"] - stamp["STARTING TO COUNT KMERS"], wall


def cpu_arm_sample(args, budget_s, runs):
    """Reads of the CPU sample: the whole workload when `runs` passes of it fit the time budget at the
    reference's measured ~0.9 M k-mers/s (large samples; small ones run up to 2 M/s), otherwise a coverage-scaled subsample that does."""
    genome_size, n_reads, read_len, paired, k, F, _ = WORKLOADS[args.workload]
    per_read = (2 if paired else 1) * (read_len - k + 2)
    fit = int(budget_s / max(runs, 1) * 0.9e6 / per_read)
    sample = max(1000, min(n_reads, fit, args.sample_reads if args.sample_reads else fit))
    return sample, host_sample_reads(genome_size, n_reads, sample, read_len, paired)


def cpu_arm_measure(args, reads, runs, warmup):
    """(k-mers/s, mean seconds per run, occurrences, kind, wall seconds per run) of the CPU arm."""
    _, _, read_len, paired, k, F, _ = WORKLOADS[args.workload]
    occ = len(reads) * (2 if paired else 1) * (read_len - k + 2)
    if os.path.exists(os.path.join(REF_DIR, "assemble.py")):
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "reads.txt")
            write_reference_input(path, reads, paired)
            times, walls = [], []
            for step in range(warmup + runs):
                dt, wall = reference_cli_step(path, k, F, sketch=getattr(args, "sketch", False))
                if step >= warmup:
                    times.append(dt)
                    walls.append(wall)
        mean = sum(times) / len(times)
        return occ / mean, mean, occ, "reference", sum(walls) / len(walls)
    times = []
    for step in range(warmup + runs):
        _, dt, occ = cpu_port_rate(reads, k, F, paired)
        if step >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return occ / mean, mean, occ, "port", mean


def cpu_sample_text(kind, sample, n_reads, paired, mean, wall):
    what = ("the unmodified reference through its CLI (oracle/_ref/assemble.py --stdout -t%s, stock code path, "
            "single-threaded pure Python); count + build seconds from its own -t prints, %.2f s of %.2f s wall per run"
            % ("", mean, wall)) if kind == "reference" else \
           "oracle/py_oracle.py (pure-Python port; oracle/_ref/ not present), %.2f s per run" % mean
    scope = "the whole workload" if sample >= n_reads else \
        "%d reads%s over a genome scaled to keep the workload's coverage" % (sample, " pairs" if paired else "")
    return "%s; %s" % (scope, what)


def workload_config(args, world):
    """The `config` object both arms print (identical keys and values)."""
    genome_size, n_reads, read_len, paired, k, F, desc = WORKLOADS[args.workload]
    if args.reads:
        n_reads = args.reads
    mates = 2 if paired else 1
    stride = (read_len + 31) // 32
    return {"workload": desc, "k": k, "filter": F, "paired": paired, "reads": n_reads,
            "occurrences": workload_occ(n_reads, read_len, paired, k), "sketch": bool(getattr(args, "sketch", False)),
            "sharding": "reads by index, k-mers by hash" if world > 1 else "none",
            "l2": "inputs larger than L2: %d MB of packed reads (and the record stream cut from them) per "
                  "step against a 126 MB L2; every table is rebuilt each step"
                  % (int(n_reads * mates * stride * 8) >> 20)}


def run_reference_arm(args):
    genome_size, n_reads, read_len, paired, k, F, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    runs = args.warmup + args.steps
    sample, reads = cpu_arm_sample(args, 150.0, runs)
    value, mean, occ, kind, wall = cpu_arm_measure(args, reads, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "k-mers/sec counted+filtered+graph-built", "value": value,
            "unit": "k-mers/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
            "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": 1, "kind": kind,
                             "host_cores": os.cpu_count(),
                             "sample": cpu_sample_text(kind, sample, n_reads, paired, mean, wall)},
            "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import ga_native as gn
    import ga_device as gd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    genome_size, n_reads, read_len, paired, k, F, desc = WORKLOADS[args.workload]
    if args.reads:
        n_reads = args.reads
    if args.genome:
        genome_size = args.genome
    sketch_widths = None
    if args.sketch_widths:          # C5: sketch width sweep over the reference's three prime tables
        from countminsketch import CountMinSketch
        sketch_widths = {"6e7": CountMinSketch.primes_6_10_7, "1e7": CountMinSketch.primes_1_10_7,
                         "5e6": CountMinSketch.primes_5_10_6}[args.sketch_widths]
    L = gn.lib()
    mates = 2 if paired else 1
    stride = (read_len + 31) // 32
    occ_total = workload_occ(n_reads, read_len, paired, k)

    # shard reads by index (contiguous ranges); every rank generates its own shard on the device
    lo = n_reads * rank // world
    hi = n_reads * (rank + 1) // world
    n_local = hi - lo
    genome = torch.empty(genome_size, dtype=torch.uint8, device=dev)
    gn.check(L.ga_gen_genome(gn.ptr(genome), genome_size, SEED, None))
    words = torch.empty(max(1, n_local * mates * stride), dtype=torch.int64, device=dev)
    gn.check(L.ga_gen_reads(gn.ptr(genome), genome_size, lo * mates, n_local * mates, read_len, SEED, 100,
                            gn.ptr(words), stride, int(paired), 125, None))
    torch.cuda.synchronize()
    reads = gd.DeviceReads.from_packed(words, n_local * mates, read_len, paired, first_read=lo,
                                       estride=read_len)
    if world > 1:
        import ga_multi
        step_fn = lambda timers=None: ga_multi.sharded_step(reads, k, F, timers=timers)   # noqa: E731
    else:
        step_fn = lambda timers=None: gd.device_step(reads, k, F, timers=timers,            # noqa: E731
                                                     sketch_rows=args.sketch_rows if args.sketch else 0,
                                                     sketch_widths=sketch_widths)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None     # NVML initialised outside the timed region
    for _ in range(args.warmup):
        step_fn()
    barrier()
    launches0 = L.ga_launch_count()
    if sampler:
        sampler.start()
    timers = {}
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    result = None
    for _ in range(args.steps):
        result = None      # one graph alive at a time: the CSR buffers of the previous step are reused
        result = step_fn(timers)
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = L.ga_launch_count() - launches0
    mem_after_steps = {"torch_reserved": round(torch.cuda.memory_reserved() / 1e9, 1),
                       "device_free": round(torch.cuda.mem_get_info()[0] / 1e9, 1),
                       "alloc_retries": torch.cuda.memory_stats().get("num_alloc_retries", 0)}
    clocks = sampler.stop() if sampler else None
    # fingerprint of the last step's CSR (outside the timed region): lines at different N, or of different exchange
    # routes, must agree -- the driver's own scaling runs then show bit-identity across 1/2/4/8 GPUs
    graph_sum = None
    if result is not None:
        try:
            graph_sum = result.checksum()
        except Exception as exc:      # noqa: BLE001  (never lose the bench line to the fingerprint)
            graph_sum = "unavailable: %r" % (exc,)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = occ_total / (ms_per_step * 1e-3)

    # dominant-kernel roofline from the per-launch CUDA events recorded inside the timed steps
    # per kernel: total ms per step, launches per step, occurrences per launch (rank 0's shard)
    host_notes = timers.pop("_host", [])
    if rank == 0 and host_notes and os.environ.get("GA_TRACE"):
        print("host: " + " | ".join(host_notes), file=sys.stderr)
    marks = timers.pop("_marks", [])
    stage_ms = {}
    for (_, ev0), (name, ev1) in zip(marks, marks[1:]):
        if name != "step begin":
            stage_ms[name] = stage_ms.get(name, 0.0) + ev0.elapsed_time(ev1) / args.steps
    kernel_ms = {name: sum(a.elapsed_time(b) for a, b, _ in spans) / args.steps for name, spans in timers.items()}
    per_rank = None
    if world > 1:       # every rank's own kernel and stage times (rank 0 alone runs the serial tail; the others wait for it)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"kernel_ms": {a: round(b, 2) for a, b in kernel_ms.items()},
                                          "stage_ms": {a: round(b, 2) for a, b in stage_ms.items()}})
    dominant = max(kernel_ms, key=kernel_ms.get)
    spans = timers[dominant]
    launch_ms = sum(a.elapsed_time(b) for a, b, _ in spans) / len(spans)
    launch_occ = sum(o for _, _, o in spans) / len(spans)
    alg_bytes = launch_occ * B_KERNEL.get(dominant, B_TOTAL)
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
    b_total = B_TOTAL_SKETCH if args.sketch else B_TOTAL
    traffic = measured_traffic(args.workload, dominant, launch_occ)
    ncu_notes = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))) \
        if os.path.exists(os.path.join(ROOT, "profiles", "traffic.json")) else {}
    limiter = ncu_notes.get("limiter", {}).get(dominant)
    roofline = {"bound": "hbm", "kernel": dominant + "_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                # `achieved` is an EQUIVALENT bandwidth (the bytes the roofline model charges per occurrence / time):
                # what the kernel really moves through DRAM, and what ncu says holds it back, stand next to it
                "dram_gbs_measured": traffic / (launch_ms * 1e-3) / 1e9 if traffic and launch_ms > 0 else None,
                # "bound" names the roofline the algorithmic bytes are held against (the contract's hbm | tensor);
                # "limited_by" is what ncu says actually holds the kernel back
                "limited_by": ncu_notes.get("limited_by", {}).get(dominant), "limiter": limiter,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per occurrence (profiles/traffic.json) "
                                  "x occurrences of this launch",
                "peak_source": peak_src,
                "kernel_ms_per_step": kernel_ms, "stage_ms_per_step": stage_ms, "per_rank": per_rank, "launch_ms": launch_ms, "launches_per_step": len(spans) / args.steps,
                "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_occurrence": B_KERNEL.get(dominant, B_TOTAL),
                "whole_path": {"achieved": occ_total / world * b_total / (ms_per_step * 1e-3) / 1e9,
                               "frac": occ_total / world * b_total / (ms_per_step * 1e-3) / 1e9 / peak,
                               "bytes_per_occurrence": b_total}}

    # end to end through the host-buffer entry (ASCII reads in pinned memory -> CSR on the host)
    e2e = None
    if world > 1 and not paired:
        import ga_multi
        ascii_dev = torch.empty(n_local * read_len, dtype=torch.uint8, device=dev)
        gn.check(L.ga_unpack_reads(gn.ptr(words), n_local, read_len, stride, 2, gn.ptr(reads.alphabet.inv_dev),
                                   gn.ptr(ascii_dev), None))
        pinned = torch.empty(ascii_dev.shape, dtype=torch.uint8, pin_memory=True)
        pinned.copy_(ascii_dev)
        torch.cuda.synchronize()
        del ascii_dev
        for _ in range(max(1, min(args.warmup, 2))):
            ga_multi.sharded_host_step(pinned, n_local, read_len, lo, k, F)
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        graph = None
        for _ in range(args.steps):
            graph = None      # drop the previous result first: its pinned CSR buffers are then reused
            graph = ga_multi.sharded_host_step(pinned, n_local, read_len, lo, k, F)
            if graph is not None:
                d2h = sum(a.nbytes for a in (graph.rowptr, graph.col, graph.indeg, graph.branching,
                                             graph.last_char, graph.keys_a))
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        e2e = {"value": occ_total / e2e_s, "unit": "k-mers/s", "h2d_bytes_per_step": int(n_reads * read_len),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3}
    if world == 1 and not os.environ.get("GA_BENCH_SKIP_E2E"):      # (A/B runs of kernel variants skip the host legs)
        ascii_dev = torch.empty(n_local * mates * read_len, dtype=torch.uint8, device=dev)
        gn.check(L.ga_unpack_reads(gn.ptr(words), n_local * mates, read_len, stride, 2,
                                   gn.ptr(reads.alphabet.inv_dev), gn.ptr(ascii_dev), None))
        pinned = torch.empty(ascii_dev.shape, dtype=torch.uint8, pin_memory=True)
        pinned.copy_(ascii_dev)
        torch.cuda.synchronize()
        del ascii_dev
        for _ in range(max(1, min(args.warmup, 2))):
            gd.host_step(pinned, n_local * mates, read_len, paired, k, F, 10 if args.sketch else 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d2h = 0
        graph = None
        for _ in range(args.steps):
            graph = None      # drop the previous result first: its pinned CSR buffers are then reused
            graph = gd.host_step(pinned, n_local * mates, read_len, paired, k, F, 10 if args.sketch else 0)
            d2h = sum(a.nbytes for a in (graph.rowptr, graph.col, graph.indeg, graph.branching,
                                         graph.last_char, graph.keys_a))
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / args.steps
        e2e = {"value": occ_total / e2e_s, "unit": "k-mers/s", "h2d_bytes_per_step": int(pinned.numel()),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
               "input": "ASCII reads, one byte per base, in pinned host memory"}
        if not paired and not args.sketch:
            # the same with the host buffer already in the 2-bit ingest format (reported next to e2e, not instead)
            del pinned
            packed = torch.empty(words.shape, dtype=torch.int64, pin_memory=True)
            packed.copy_(words)
            torch.cuda.synchronize()
            graph = None
            gd.host_step_packed(packed, n_local, read_len, k, F)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                graph = None
                graph = gd.host_step_packed(packed, n_local, read_len, k, F)
            torch.cuda.synchronize()
            packed_s = (time.perf_counter() - t0) / args.steps
            e2e["packed_input"] = {"value": occ_total / packed_s, "unit": "k-mers/s", "ms_per_step": packed_s * 1e3,
                                   "h2d_bytes_per_step": int(packed.numel() * 8),
                                   "input": "reads packed 2 bits per base on the host (ga_reads layout)"}
            del packed

    # The call a user of the reference makes: the reference's stdin text -> IOHandler.read_input (raw-byte parser)
    # -> DeBruijnGraph / PairedDeBruijnGraph(reads, k, F) -> enumerate_contigs(), everything on the clock.
    if world == 1 and e2e is not None and n_reads * mates * read_len <= (1 << 30) and not args.sketch:
        import io
        import assemble as cli
        import debruijn_graph as dg
        ascii_dev = torch.empty(n_local * mates * read_len, dtype=torch.uint8, device=dev)
        gn.check(L.ga_unpack_reads(gn.ptr(words), n_local * mates, read_len, stride, 2,
                                   gn.ptr(reads.alphabet.inv_dev), gn.ptr(ascii_dev), None))
        rows = ascii_dev.cpu().numpy().reshape(n_local, mates * read_len)
        del ascii_dev
        if paired:
            tail = np.frombuffer(b"|125\n", dtype=np.uint8)
            lines = np.concatenate([rows[:, :read_len], np.full((n_local, 1), ord("|"), dtype=np.uint8),
                                    rows[:, read_len:], np.broadcast_to(tail, (n_local, tail.size))], axis=1)
        else:
            lines = np.concatenate([rows, np.full((n_local, 1), ord("\n"), dtype=np.uint8)], axis=1)
        text = (b"%d\n" % n_local) + lines.tobytes()
        del rows, lines
        cls = dg.PairedDeBruijnGraph if paired else dg.DeBruijnGraph
        best = None
        for _ in range(1 + max(2, min(args.steps, 5))):           # first pass is warm-up
            t0 = time.perf_counter()
            parsed, is_paired, _, _ = cli.IOHandler.read_input(io.BytesIO(text))
            t1 = time.perf_counter()
            g = cls(parsed, k=k, hamming_dist=F)
            t2 = time.perf_counter()
            contigs = g.enumerate_contigs()
            t3 = time.perf_counter()
            cur = {"ms_ingest": (t1 - t0) * 1e3, "ms_count_filter_build": (t2 - t1) * 1e3,
                   "ms_contigs": (t3 - t2) * 1e3, "ms_total": (t3 - t0) * 1e3, "contigs": len(contigs)}
            if best is None or _ == 1 or cur["ms_total"] < best["ms_total"]:
                best = cur
            del g, parsed
        best.update({"value": occ_total / ((best["ms_ingest"] + best["ms_count_filter_build"]) * 1e-3), "unit": "k-mers/s",
                     "input_bytes": len(text),
                     "call": "assemble.IOHandler.read_input(stdin bytes) -> %s(reads, k, F) -> enumerate_contigs(); "
                             "value counts ingest + count/filter/build, best of the timed passes" % cls.__name__})
        e2e["user_api"] = best
        del text

    cpu_baseline = None
    if rank == 0:
        sample, sreads = cpu_arm_sample(args, 20.0, 1)
        rate, mean, occ, kind, wall = cpu_arm_measure(args, sreads, 1, 0)
        cpu_baseline = {"value": rate, "unit": "k-mers/s", "cores": 1, "kind": kind,
                        "sample": cpu_sample_text(kind, sample, n_reads, paired, mean, wall),
                        "host_cores": os.cpu_count()}
    if rank == 0:
        line = {"metric": "k-mers/sec counted+filtered+graph-built", "value": value, "unit": "k-mers/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic",
                "config": workload_config(args, world),
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
                "memory_gb": {"torch_reserved": round(torch.cuda.memory_reserved() / 1e9, 1),
                              "torch_peak_allocated": round(torch.cuda.max_memory_allocated() / 1e9, 1),
                              "device_free": round(torch.cuda.mem_get_info()[0] / 1e9, 1),
                              "after_timed_steps": mem_after_steps},
                "clocks": clocks,
                "graph": {"nodes": result.n_nodes, "edges": result.n_edges, "checksum": graph_sum}
                if result is not None else None}
        print(json.dumps(line))
    if world > 1:
        import ga_multi
        ga_multi.release_peers()           # collective: unmap and free the NVLink exchange buffers
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GA_BENCH_WORKLOAD", "c4"), choices=sorted(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="override the number of reads / pairs")
    ap.add_argument("--genome", type=int, default=0, help="override the genome size (profiling: scale reads and "
                                                         "genome together to keep the workload's coverage)")
    ap.add_argument("--sketch", action="store_true", help="the -c route: exact counts poured into the reference's "
                                                          "10-row CountMinSketch, filter on the sketch estimate")
    ap.add_argument("--k", type=int, default=0, help="override --kmer_length of the workload (C5: k sweep on C3's read pairs)")
    ap.add_argument("--sketch-rows", type=int, default=10, help="rows of the CountMinSketch with --sketch (reference: 10, "
                                                                "8 in its -m mode)")
    ap.add_argument("--sketch-widths", default="", choices=["", "6e7", "1e7", "5e6"],
                    help="prime table the sketch rows are taken from (reference: 1e7; countminsketch.py:9-24)")
    ap.add_argument("--sample-reads", type=int, default=0, help="cap on the reads (pairs) of the CPU sample "
                                                                "(default: what fits the time budget)")
    args = ap.parse_args()
    if args.k:                      # C5: k sweep on the read pairs of C3 (64-bit keys up to k = 32, 128-bit beyond)
        g, n, length, paired, _, F, desc = WORKLOADS[args.workload]
        WORKLOADS[args.workload] = (g, n, length, paired, args.k, F, desc.replace("k=%d" % _, "k=%d" % args.k) + " [k overridden]")
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
