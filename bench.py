#!/usr/bin/env python
"""bench.py — GNSS IQ synthesis Msamples/s + PCPS acquisition cells/s on N B200s (BASELINE.json's metric).

One STEP = one pass of the hot path over one batch of synthetic input on every rank:
  (1) synthesise the rank's 20 s time segment of e1c_8prn_20s_clean.yaml (8 Galileo E1C PRNs, noise on,
      1e8 samples, cf32) straight into HBM, then
  (2) run PCPS over consecutive 4 ms snapshots of that segment for the config's 8 PRNs on the
      20 000-lag x 41-Doppler grid (+-5 kHz / 250 Hz) and gather the peak table.
Weak scaling: with N ranks the scenario is N x 20 s long and rank r owns segment r (time sharding, no data-path
collective; one all-gather of the peak table).  `value` is synthesis throughput (the metric's first half) and
the `acq` object carries the cells/s half with its own roofline / e2e / cpu_baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --workload e1c_8prn_60s_cn34_orbital.yaml      # another BASELINE.json config as the segment (weak: one per GPU)
  python bench.py --workload e1c_8prn_600s_cn34_orbital.yaml --strong --acq-snapshots 5000   # config 5: 600 s split over the ranks
  python bench.py --impl reference [...]                         # the reference algorithm on the host cores
R4WB_BENCH_DEBUG=1 makes every rank print its own synthesis time, kernel times and host enqueue time on stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "e1c_8prn_20s_clean.yaml"
SEGMENT_S = 20.0
CODE_LENGTH = 20000            # samples per E1C primary-code period at 5 MHz
DOPPLER_MAX, DOPPLER_STEP = 5000.0, 250.0
E2E_MAX_SAMPLES = 300_000_000  # the host-buffer (e2e) legs cover at most this many samples of a rank's segment
FLOP_PER_CELL = 264.6          # reference-equivalent flop per (PRN, Doppler, lag) cell
FP32_PEAK_TFLOPS = 74.4        # 148 SM x 128 lanes x 2 x 1.965 GHz (nominal; MEASURED_PEAKS.json has no FP32 figure)
HBM_FALLBACK_GBS = 6650.0      # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture of this
    bench command (profiles/ncu_traffic.json, written by tools/ncu_summary.py --traffic); None when not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_workload(n_ranks: int, args=None):
    """The scenario all ranks share.  Default: the 20 s bench workload, weak-scaled (N x 20 s).  `--workload X.yaml` names
    another of BASELINE.json's configs: weak scaling repeats its duration per GPU, `--strong` keeps the config's own
    duration and splits it over the ranks (config 5: the 600 s file time-sharded across 2/4/8 GPUs)."""
    global WORKLOAD, SEGMENT_S
    from r4w_b200.config import load_config
    if args is not None and args.workload:
        WORKLOAD = os.path.basename(args.workload)
    cfg = load_config(os.path.join(ROOT, "configs", WORKLOAD), cli_elevation_mask_deg=5.0)   # the CLI's default mask
    if args is not None and args.workload:
        SEGMENT_S = cfg.output.duration_s / n_ranks if args.strong else cfg.output.duration_s
    cfg.output.duration_s = SEGMENT_S * n_ranks
    return cfg


# ------------------------------------------------------------------------------------------------- CPU legs
def cpu_synth(cfg, seconds: float, threads: int):
    """oracle port of GnssScenario::generate_block, one thread per satellite (the reference's rayon `parallel` feature)."""
    from oracle import oracle as O
    c = cfg.copy()
    c.output.duration_s = seconds
    sc = O.OracleScenario(c, noise=True, threads=threads)
    bs = sc.block_size()
    t = time.perf_counter()
    n = 0
    while not sc.is_done():
        n += sc.generate_block(bs).size
    dt = time.perf_counter() - t
    return n / dt / 1e6, n, dt


def cpu_acq(x: np.ndarray, codes: np.ndarray, prns, n_snap: int, threads: int):
    """oracle port of PcpsAcquisition::acquire; independent (snapshot, PRN) calls spread over host threads
    (ctypes drops the GIL; the reference itself has no parallel acquisition)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    acq = O.OraclePcps(CODE_LENGTH, 5e6).with_doppler_range(DOPPLER_MAX, DOPPLER_STEP)
    x64 = np.ascontiguousarray(x[: n_snap * CODE_LENGTH], np.complex128)
    jobs = [(s, c) for s in range(n_snap) for c in range(len(prns))]

    def one(j):
        s, c = j
        r = acq.acquire(x64[s * CODE_LENGTH:(s + 1) * CODE_LENGTH], codes[c], prns[c])
        return int(r.code_phase)

    t = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, jobs))
    dt = time.perf_counter() - t
    cells = len(jobs) * acq.num_bins() * CODE_LENGTH
    return cells / dt, cells, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Rust crate cannot be built in this image)
    on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    cfg = load_workload(max(1, args.gpus), args)
    n_sats = len(cfg.satellites)
    th_s = max(1, min(n_sats, cores))
    prns = [s.prn for s in cfg.satellites]
    codes = np.stack([O.e1c_replica(p, 5e6, CODE_LENGTH) for p in prns])
    seconds = 0.1                                     # 0.5 Msamples per synthesis step
    n_snap = max(1, min(2, cores // 8 + 1))
    x = None
    syn, acq = [], []
    for it in range(args.warmup + args.steps):
        v, n, dt = cpu_synth(cfg, seconds, th_s)
        if x is None:
            c = cfg.copy(); c.output.duration_s = n_snap * CODE_LENGTH / 5e6
            x = O.to_cf32(O.OracleScenario(c, noise=True).generate_range(0, n_snap * CODE_LENGTH))
        a, cells, adt = cpu_acq(x, codes, prns, n_snap, cores)
        if it >= args.warmup:
            syn.append((v, dt)); acq.append((a, adt))
    v = float(np.mean([s[0] for s in syn])); a = float(np.mean([s[0] for s in acq]))
    ms = float(np.mean([s[1] for s in syn])) * 1e3
    sample = f"synthesis: {seconds} s of the scenario per step ({int(seconds*5e6)} samples), {th_s} threads (one per satellite); " \
             f"acquisition: {n_snap} snapshot(s) x {len(prns)} PRNs x 41 bins per step, {cores} threads"
    line = {
        "impl": "reference", "metric": "gnss_iq_synth_msamples_per_s", "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: 8 Galileo E1C PRNs, 5 MS/s, noise on; bounded sample per step", "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": th_s, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "acq": {"metric": "pcps_acq_cells_per_s", "value": a, "unit": "cells/s", "ms_per_step": float(np.mean([s[1] for s in acq])) * 1e3,
                "cpu_baseline": {"value": a, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": a, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import r4w_b200 as R
    from r4w_b200 import _lib
    from r4w_b200.dist import all_gather_table, all_reduce_power, results_to_table, segment_for_rank

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    R.init(local)
    dev = torch.device("cuda", local)

    cfg = load_workload(world, args)
    prns = [s.prn for s in cfg.satellites]
    codes = np.stack([R.e1c_replica(p, 5e6, CODE_LENGTH) for p in prns])
    scen = R.GnssScenario(cfg, noise=True)
    scen.set_profiling(True)
    first, n = segment_for_rank(scen.total_samples(), CODE_LENGTH, rank, world)
    n_snap_total = n // CODE_LENGTH
    n_snap = n_snap_total if args.acq_snapshots <= 0 else min(args.acq_snapshots, n_snap_total)
    acq = R.PcpsAcquisition(CODE_LENGTH, 5e6).with_doppler_range(DOPPLER_MAX, DOPPLER_STEP)
    acq.set_profiling(True)
    bins = acq.num_doppler_bins()
    iq = torch.empty(n, dtype=torch.complex64, device=dev)       # 0.8 GB: larger than the 126 MB L2, no flush needed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(record):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        th = time.perf_counter()
        scen.generate_device(first, n, iq)
        th = time.perf_counter() - th
        e1.record()
        pods = acq.acquire_batch_raw(iq, n_snap, CODE_LENGTH, CODE_LENGTH, codes, prns)
        table = all_gather_table(results_to_table(pods, n_snap, len(prns)))
        e2.record()
        if record is not None:
            record.append((e0, e1, e2, acq.last_profile(), acq.guard_count(), scen.last_profile(), th))
        return table

    for _ in range(max(args.warmup, 0)):
        step(None)
    barrier()
    launches0 = R.kernel_launches()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    rec = []
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    table = None
    for _ in range(args.steps):
        table = step(rec)
    t1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = R.kernel_launches() - launches0
    ms_syn = float(np.mean([r[0].elapsed_time(r[1]) for r in rec]))
    ms_acq = float(np.mean([r[1].elapsed_time(r[2]) for r in rec]))
    ms_total = t0.elapsed_time(t1) / args.steps
    prof = rec[-1][3]
    sprof = {k: (float(np.mean([r[5][k][0] for r in rec])), rec[-1][5][k][1]) for k in rec[-1][5]}
    guards = int(np.sum([r[4] for r in rec]))
    if os.environ.get("R4WB_BENCH_DEBUG"):
        print(f"[rank {rank}] synth e0->e1 {ms_syn:.3f} ms, kernels {dict((k, round(v[0], 3)) for k, v in sprof.items())}, "
              f"host enqueue {np.mean([r[6] for r in rec]) * 1e3:.3f} ms, first {first}, n {n}", file=sys.stderr, flush=True)

    # the CLI's average-power line (main.rs:4494-4509) over all ranks' segments: one all-reduce of {sum |s|^2, count}
    scen.generate_device(first, n, iq)
    torch.cuda.synchronize()
    pw_sum, pw_cnt = all_reduce_power(scen.last_power_sum(), n)
    # the reference's default Doppler grid (+-5 kHz / 500 Hz = 21 bins, acquisition.rs:63-74) on a 296-snapshot sample
    acq21 = R.PcpsAcquisition(CODE_LENGTH, 5e6)
    ns21 = min(n_snap, 296)
    acq21.acquire_batch_raw(iq, ns21, CODE_LENGTH, CODE_LENGTH, codes, prns)
    torch.cuda.synchronize()
    ts = time.perf_counter()
    acq21.acquire_batch_raw(iq, ns21, CODE_LENGTH, CODE_LENGTH, codes, prns)
    torch.cuda.synchronize()
    ms_acq21 = (time.perf_counter() - ts) * 1e3
    cells21 = ns21 * len(prns) * acq21.num_doppler_bins() * CODE_LENGTH

    # ---- e2e: the same step through the C-ABI with HOST buffers (pinned), copies inside the timed region
    import ctypes as C
    host = C.c_void_p()
    n_full, n_snap_full = n, n_snap
    n = min(n, E2E_MAX_SAMPLES)                                  # host-buffer legs: at most 60 s of samples (2.4 GB pinned)
    n_snap = min(n_snap, n // CODE_LENGTH)
    _lib.check(_lib.lib().r4wb_host_alloc(C.byref(host), n * 8))
    host_np = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(2 * n,)).view(np.complex64)
    e2e_steps = max(1, min(args.steps, 3))
    _lib.set_stream(torch.cuda.current_stream().cuda_stream)
    scen.generate_range_into(first, n, host.value)              # warm-up (staging buffers, page faults)
    acq.acquire_batch_raw(host_np, n_snap, CODE_LENGTH, CODE_LENGTH, codes, prns)      # warm-up at full size: the library's device input
    #                                                                                   buffer (0.8 GB) and event pool are allocated here
    barrier()
    ts = time.perf_counter()
    for _ in range(e2e_steps):
        scen.generate_range_into(first, n, host.value)
    torch.cuda.synchronize()
    ms_syn_e2e = (time.perf_counter() - ts) * 1e3 / e2e_steps
    # the CLI's integer sink formats, converted in the store epilogue: fewer bytes over PCIe
    fmt_e2e = {}
    for name, code, bps in (("ci16", _lib.FMT_CI16, 4), ("ci8", _lib.FMT_CI8, 2)):
        scen.generate_range_into(first, n, host.value, fmt=code)
        torch.cuda.synchronize()
        ts = time.perf_counter()
        scen.generate_range_into(first, n, host.value, fmt=code)
        torch.cuda.synchronize()
        fmt_e2e[name] = {"msamples_per_s": n / (time.perf_counter() - ts) / 1e6, "d2h_bytes_per_step": n * bps}
    scen.generate_range_into(first, n, host.value)              # host buffer back to cf32 for the acquisition leg
    torch.cuda.synchronize()
    ts = time.perf_counter()
    for _ in range(e2e_steps):
        pods = acq.acquire_batch_raw(host_np, n_snap, CODE_LENGTH, CODE_LENGTH, codes, prns)
        tab2 = results_to_table(pods, n_snap, len(prns))
    torch.cuda.synchronize()
    ms_acq_e2e = (time.perf_counter() - ts) * 1e3 / e2e_steps
    n_e2e, n_snap_e2e, n, n_snap = n, n_snap, n_full, n_snap_full

    # ---- tracking channels, rank 0, outside the timed step: one E1C channel per satellite of the scenario over
    # the first second of the rendered stream (250 code periods of 20 000 samples), device-resident input
    track = None
    if rank == 0 and not args.no_track:
        # replica per half-chip (E1C chips x BOC(1,1) sub-carrier): 8 184 "chips" at 2.046 Mchip/s, so the reference's per-chip
        # tracker (gnss/tracking.rs) sees the sub-carrier; start values from the acquisition of the first snapshot
        e1c = np.stack([np.repeat(R.e1_code(1, p).astype(np.int8), 2) * np.tile(np.array([1, -1], np.int8), 4092) for p in prns])
        chans = [dict(prn=p, code_length=8184, sample_rate=5e6, chipping_rate=2.046e6,
                      initial_code_phase=float(((CODE_LENGTH - int(table[0, c, 2])) % CODE_LENGTH) * 2.046e6 / 5e6),
                      initial_doppler=float(table[0, c, 3])) for c, p in enumerate(prns)]
        n_per, n_p = CODE_LENGTH, 250
        R.TrackerBank(chans).process(iq, e1c, n_per, 8)            # warm-up
        bank = R.TrackerBank(chans)
        torch.cuda.synchronize()
        ts = time.perf_counter()
        st = bank.process(iq, e1c, n_per, n_p)
        dt = time.perf_counter() - ts
        track = {"metric": "tracking_channel_periods_per_s", "value": len(prns) * n_p / dt, "unit": "channel-periods/s",
                 "channels": len(prns), "periods": n_p, "samples_per_period": n_per, "ms": dt * 1e3,
                 "api": "r4wb_track_process(..., R4WB_MEM_DEVICE): states D2H inside the timed call",
                 "locked_channels": int(np.sum(st["code_lock"][-1]))}
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle as O
            O.build()
            xs = iq[: 25 * n_per].cpu().numpy()
            tc = time.perf_counter()
            O.OracleTrackingChannel(prns[0], 8184, 5e6, 2.046e6, chans[0]["initial_code_phase"], chans[0]["initial_doppler"]).run(xs, e1c[0], n_per, 25)
            dtc = time.perf_counter() - tc
            track["cpu_baseline"] = {"value": 25 / dtc, "unit": "channel-periods/s", "cores": 1, "kind": "port",
                                     "sample": f"1 channel x 25 periods ({dtc:.2f} s wall), oracle port, 1 thread"}

    # ---- max over ranks
    def rmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_syn, ms_acq, ms_total, ms_syn_e2e, ms_acq_e2e = (rmax(v) for v in (ms_syn, ms_acq, ms_total, ms_syn_e2e, ms_acq_e2e))
    inv_ms, inv_n = prof["inverse_fft_peak"]
    fwd_ms, fwd_n = prof["forward_fft"]
    inv_ms, fwd_ms = rmax(inv_ms), rmax(fwd_ms)
    # dominant synthesis kernel of the step and the samples it wrote (the period-resident kernel renders whole primary-code
    # periods, k_synth the partial periods at the ends; k_synth alone when the scenario is not static)
    syn_kernel = "k_synth_periodic" if sprof["k_synth_periodic"][1] else "k_synth"
    syn_kernel_ms = rmax(sprof[syn_kernel][0])
    if syn_kernel == "k_synth_periodic":
        k_lo, k_hi = max(1, -(-first // CODE_LENGTH)), (first + n) // CODE_LENGTH
        syn_kernel_samples = (k_hi - k_lo) * CODE_LENGTH
    else:
        syn_kernel_samples = n

    total_samples = n * world
    cells_rank = n_snap * len(prns) * bins * CODE_LENGTH
    cells = cells_rank * world
    peak, peak_src = hbm_peak()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        cores = os.cpu_count() or 1
        th_s = max(1, min(len(prns), cores))
        v_s, n_s, dt_s = cpu_synth(cfg, 3.0, th_s)                    # 3 s of the scenario: ~10-15 s of host work
        v_1, n_1, dt_1 = cpu_synth(cfg, 0.3, 1)                       # the reference's default build is single-threaded
        ns_c = max(1, min(n_snap, 8 * cores))                         # ~10 s of host work on all cores
        v_a, c_a, dt_a = cpu_acq(host_np, codes, prns, ns_c, cores)
        cpu = ({"value": v_s, "unit": "Msamples/s", "cores": th_s, "kind": "port",
                "sample": f"3 s of {WORKLOAD} ({n_s} samples, {dt_s:.1f} s wall), oracle port, one thread per satellite",
                "single_thread": {"value": v_1, "unit": "Msamples/s", "cores": 1,
                                  "sample": f"0.3 s of {WORKLOAD} ({n_1} samples, {dt_1:.1f} s wall), oracle port, 1 thread"}},
               {"value": v_a, "unit": "cells/s", "cores": cores, "kind": "port",
                "sample": f"{ns_c} snapshot(s) x {len(prns)} PRNs x {bins} bins ({c_a} cells, {dt_a:.1f} s wall), oracle port, {cores} threads"})
    _lib.check(_lib.lib().r4wb_host_free(host))

    if rank == 0:
        synth_gbs = syn_kernel_samples * 8 / (syn_kernel_ms * 1e-3) / 1e9
        acq_tflops = cells_rank * FLOP_PER_CELL / ((inv_ms + fwd_ms) * 1e-3) / 1e12 if inv_ms + fwd_ms > 0 else None
        line = {
            "metric": "gnss_iq_synth_msamples_per_s", "value": total_samples / (ms_syn * 1e-3) / 1e6, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_syn, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD} x {world} GPU(s): {SEGMENT_S:g} s segment per GPU (time-sharded), {len(prns)} Galileo E1C PRNs, "
                                   f"5 MS/s, noise on, cf32 into HBM; then PCPS over {n_snap} snapshots/GPU x {len(prns)} PRNs x {bins} Doppler bins x "
                                   f"{CODE_LENGTH} lags", "samples_per_gpu": n, "snapshots_per_gpu": n_snap, "prns": prns,
                       "l2": f"per-step output {n * 8 / 1e9:.1f} GB and spectra working set exceed the 126 MB L2 (no explicit flush)",
                       "step": "synth then acquire; ms_per_step/value cover the synthesis half, acq.* the acquisition half, ms_step_total both"},
            "ms_step_total": ms_total,
            "roofline": {"bound": "hbm", "kernel": syn_kernel, "achieved": synth_gbs, "peak": peak, "unit": "GB/s", "frac": synth_gbs / peak,
                         "traffic": ncu_traffic(syn_kernel), "peak_source": peak_src, "algorithmic_bytes_per_launch": syn_kernel_samples * 8,
                         "kernel_ms": syn_kernel_ms,
                         "note": "8 B per output sample (one cf32 store) x the samples this kernel wrote / its CUDA-event time on the launching "
                                 "stream; `value` is over the whole generate call (all kernels + launch gaps)"},
            "synth_kernel_ms": {k: v[0] for k, v in sprof.items()}, "synth_kernel_launches": {k: v[1] for k, v in sprof.items()},
            "e2e": {"value": n_e2e * world / (ms_syn_e2e * 1e-3) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": n_e2e * 8, "samples_per_gpu": n_e2e, "api": "r4wb_scenario_generate(..., R4WB_MEM_HOST, CF32) into pinned host memory",
                    "other_formats_rank0": fmt_e2e},
            "acq": {"metric": "pcps_acq_cells_per_s", "value": cells / (ms_acq * 1e-3), "unit": "cells/s", "ms_per_step": ms_acq,
                    "cells_per_step": cells, "f64_guard_reruns": guards,
                    "kernel_ms": {k: v[0] for k, v in prof.items()}, "kernel_launches": {k: v[1] for k, v in prof.items()},
                    "roofline": {"bound": "fp32", "kernel": "k_rf_inv_peak_tm (+k_rf_fwd)", "achieved": acq_tflops, "peak": FP32_PEAK_TFLOPS,
                                 "unit": "TFLOP/s", "frac": (acq_tflops / FP32_PEAK_TFLOPS) if acq_tflops else None, "traffic": ncu_traffic("k_rf_inv_peak_tm"),
                                 "peak_source": "nominal FP32 FMA peak (148 SM x 128 lanes x 2 x 1.965 GHz)",
                                 "note": f"{FLOP_PER_CELL} reference-equivalent flop per cell / summed CUDA-event time of the forward and inverse FFT kernels"},
                    "e2e": {"value": n_snap_e2e * len(prns) * bins * CODE_LENGTH * world / (ms_acq_e2e * 1e-3), "unit": "cells/s",
                            "h2d_bytes_per_step": n_snap_e2e * CODE_LENGTH * 8, "d2h_bytes_per_step": n_snap_e2e * len(prns) * 32,
                            "snapshots_per_gpu": n_snap_e2e,
                            "api": "r4wb_pcps_acquire_batch(..., R4WB_MEM_HOST) from pinned host memory"},
                    "first_snapshot": [[int(table[0, c, 2]), float(table[0, c, 3])] for c in range(len(prns))],
                    "default_grid_rank0": {"bins": acq21.num_doppler_bins(), "snapshots": ns21, "cells_per_s": cells21 / (ms_acq21 * 1e-3),
                                           "note": "reference default +-5 kHz / 500 Hz, device-resident input, wall clock of one call"}},
            "avg_power": {"value": pw_sum / max(pw_cnt, 1), "samples": pw_cnt, "note": "sum |s|^2 / count all-reduced over the ranks (the CLI's avg-power line)"},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        if track is not None:
            line["track"] = track
        if cpu is not None:
            line["cpu_baseline"] = cpu[0]
            line["acq"]["cpu_baseline"] = cpu[1]
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Everything a library prints on fd 1 from here on (NCCL's version banner at communicator creation, for one) goes to
    stderr; `emit` writes the ONE JSON line of the contract on the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--acq-snapshots", type=int, default=0, help="snapshots per GPU per step (0 = the whole segment)")
    ap.add_argument("--workload", default="", help="another config of BASELINE.json (file name under configs/); default " + WORKLOAD)
    ap.add_argument("--strong", action="store_true", help="with --workload: split the config's own duration over the ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-track", action="store_true", help="skip the tracking-channel leg (SURVEY.md section 8 f2)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
