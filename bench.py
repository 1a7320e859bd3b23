#!/usr/bin/env python
"""bench.py — GNSS IQ synthesis Msamples/s + PCPS acquisition cells/s on N B200s (BASELINE.json's metric).

Default workload (what the driver runs): BASELINE config 5, `e1c_8prn_600s_cn34_orbital.yaml` — 8 Galileo E1C PRNs with
orbital Doppler / delay dynamics, C/N0 34 dB-Hz AWGN, 600 s = 3e9 samples = 24 GB of cf32 — STRONG-scaled: the one file is
time-sharded over the N ranks (N = 1: the whole 24 GB in one HBM; rank r of N renders [r, r+1) x 600/N s).  One STEP =
  (1) the rank's time segment synthesised straight into HBM (`value`, Msamples/s, whole job = 3e9 samples / max-over-ranks time),
  (2) PCPS over the rank's share of a fixed set of 4 ms snapshots of that stream for the config's 8 PRNs on the
      20 000-lag x 41-Doppler grid (+-5 kHz / 250 Hz), and one all-gather of the peak table (`acq`).
No data-path collective exists.  The one-time table / phase prologue of a cold handle is measured
separately (`prologue_ms`, `value_cold`).  `per_config` repeats the measurement (bounded repetitions) for all five BASELINE
configs at the rank's share, config 4 with the full PRN 1-50 x 41-bin grid.  Random (snapshot, PRN) results of every
acquisition leg are compared with the oracle (`parity_checked` / `parity_mismatches`; the oracle is the checker only).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --workload e1c_8prn_20s_clean.yaml --weak      # another config; --weak = one copy of its duration per GPU
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

CONFIGS = ["e1c_prn3_20s_withdoppler.yaml", "e1c_8prn_20s_clean.yaml", "e1c_8prn_60s_cn34_orbital.yaml", "e1c_60s_all_prns.yaml",
           "e1c_8prn_600s_cn34_orbital.yaml"]          # BASELINE.json `configs`, in order
WORKLOAD = CONFIGS[4]
SEGMENT_S = 600.0
CODE_LENGTH = 20000            # samples per E1C primary-code period at 5 MHz
DOPPLER_MAX, DOPPLER_STEP = 5000.0, 250.0
ACQ_SNAPSHOTS_TOTAL = 23680    # snapshots of the main workload acquired per step over all ranks (160 waves of 148 rows at N = 1)
ACQ_SNAPSHOTS_CONFIG = 592     # per_config legs: snapshots per rank (4 x 148)
E2E_MAX_SAMPLES = 300_000_000  # the host-buffer (e2e) legs cover at most this many samples of a rank's segment
FLOP_PER_CELL = 264.6          # reference-equivalent flop per (PRN, Doppler, lag) cell
FP32_PEAK_TFLOPS = 74.4        # 148 SM x 128 lanes x 2 x 1.965 GHz (nominal; the measured figure replaces it when the probe runs)
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


def load_workload(n_ranks: int, args=None, name: str = None):
    """The scenario all ranks share.  Strong scaling (default) keeps the config's own duration and splits it over the
    ranks; `--weak` repeats the config's duration once per GPU (rank r owns copy r)."""
    global WORKLOAD, SEGMENT_S
    from r4w_b200.config import load_config
    if name is None:
        if args is not None and args.workload:
            WORKLOAD = os.path.basename(args.workload)
        name = WORKLOAD
    cfg = load_config(os.path.join(ROOT, "configs", name), cli_elevation_mask_deg=5.0)   # the CLI's default mask
    weak = bool(args is not None and args.weak)
    seg = cfg.output.duration_s if weak else cfg.output.duration_s / n_ranks
    if name == WORKLOAD:
        SEGMENT_S = seg
    if weak:
        cfg.output.duration_s = cfg.output.duration_s * n_ranks
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
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak" if args.weak else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: {n_sats} Galileo E1C PRN(s), 5 MS/s, noise on; bounded sample per step", "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": th_s, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "acq": {"metric": "pcps_acq_cells_per_s", "value": a, "unit": "cells/s", "ms_per_step": float(np.mean([s[1] for s in acq])) * 1e3,
                "cpu_baseline": {"value": a, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": a, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------- GPU arm
def numa_bind(local: int):
    """Pin this rank's host threads and its future host allocations (the pinned e2e buffers) to the NUMA node of its GPU:
    eight ranks copying device -> host through one socket's memory controllers is what capped the aggregate D2H rate in
    round 1.  Best effort: a container may not allow the CPUs / memory nodes of the other socket; what happened is reported."""
    info = {"gpu": local}
    try:
        import ctypes
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read().strip())
        cpus = open(base + "/local_cpulist").read().strip()
        info.update({"pci": bdf, "numa_node": node, "local_cpulist": cpus})
        want = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-"); want.update(range(int(a), int(b) + 1))
            elif part:
                want.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = want & allowed
        if use:
            os.sched_setaffinity(0, use)
            info["cpu_affinity"] = f"{len(use)} local CPUs"
        else:
            info["cpu_affinity"] = f"unchanged ({len(allowed)} allowed CPUs, none local to the GPU)"
        if node >= 0:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            MPOL_PREFERRED = 1
            rc = libc.syscall(238, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))      # set_mempolicy (x86-64)
            info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy failed (errno %d)" % ctypes.get_errno()
    except Exception as e:                                                                   # noqa: BLE001
        info["error"] = repr(e)[:120]
    return info


def fp32_peak():
    """Measured FP32 FMA rate of this box (tools/ubench/fp32_peak, built by __graft_entry__.build()) or the nominal figure."""
    exe = os.path.join(ROOT, "tools", "ubench", "fp32_peak")
    try:
        out = subprocess.run([exe, "--json"], capture_output=True, text=True, timeout=60).stdout
        d = json.loads(out.strip().splitlines()[-1])
        return float(d["ffma_tflops"]), f"measured on this GPU (tools/ubench/fp32_peak: FFMA {d.get('ffma_scalar_tflops')} / FFMA2 {d.get('ffma2_tflops')} TFLOP/s; {d.get('detail', '')})"
    except Exception:
        return FP32_PEAK_TFLOPS, "nominal FP32 FMA peak (148 SM x 128 lanes x 2 x 1.965 GHz); tools/ubench/fp32_peak did not run"


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import r4w_b200 as R
    from r4w_b200 import _lib
    from r4w_b200.dist import all_gather_table, all_reduce_power, results_to_table, segment_for_rank

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = numa_bind(local) if not args.no_numa_bind else {"gpu": local, "skipped": True}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R.init(local)
    dev = torch.device("cuda", local)
    debug = bool(os.environ.get("R4WB_BENCH_DEBUG"))
    rng = np.random.default_rng(1234 + rank)
    peak, peak_src = hbm_peak()
    fpeak, fpeak_src = fp32_peak() if rank == 0 else (FP32_PEAK_TFLOPS, "")
    oracle_mod = None
    if not args.no_parity:
        from oracle import oracle as oracle_mod           # the CHECKER of the parity samples below; never on the measured path
        if rank == 0:
            oracle_mod.build()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def rmax(vals):
        """max over ranks of a list of floats"""
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def rsum(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    barrier()
    # every kernel image loaded and every one-time attribute set before anything is timed
    warm_cfg = load_workload(1, None, CONFIGS[2]); warm_cfg.output.duration_s = 0.2
    wsc = R.GnssScenario(warm_cfg, noise=True)
    wbuf = torch.empty(1_000_000, dtype=torch.complex64, device=dev)
    wsc.generate_device(0, 1_000_000, wbuf)
    wsc.close(); del wbuf
    torch.cuda.synchronize()

    main_cfg = load_workload(world, args)
    # one device buffer for every leg: the main workload's share is the largest
    def share(cfg):
        probe = R.GnssScenario(cfg, noise=True)
        tot = probe.total_samples()
        probe.close()
        return segment_for_rank(tot, CODE_LENGTH, rank, world), tot
    (first_main, n_main), total_main = share(main_cfg)
    iq = torch.empty(max(n_main, 1), dtype=torch.complex64, device=dev)    # 24 GB / N: far beyond the 126 MB L2, no flush needed
    host = C.c_void_p()
    n_host = min(n_main, E2E_MAX_SAMPLES)
    _lib.check(_lib.lib().r4wb_host_alloc(C.byref(host), max(n_host, 1) * 8))
    host_np = np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_float)), shape=(2 * max(n_host, 1),)).view(np.complex64)

    parity = {"checked": 0, "mismatches": 0, "detail": []}

    def check_parity(cfg_name, table, n_snap, codes, prns, acq, how_many):
        """`how_many` random (snapshot, PRN) results of `table` against the oracle's acquire on the same device samples"""
        if oracle_mod is None or n_snap == 0:
            return
        oacq = oracle_mod.OraclePcps(CODE_LENGTH, 5e6).with_doppler_range(acq.doppler_max_hz, acq.doppler_step_hz)
        for _ in range(how_many):
            s_i, c_i = int(rng.integers(0, n_snap)), int(rng.integers(0, len(prns)))
            x = iq[s_i * CODE_LENGTH:(s_i + 1) * CODE_LENGTH].cpu().numpy().astype(np.complex128)
            o = oacq.acquire(x, codes[c_i], prns[c_i])
            ok = (float(table[s_i, c_i, 2]), float(table[s_i, c_i, 3]), bool(table[s_i, c_i, 1])) == (o.code_phase, o.doppler_hz, bool(o.detected))
            parity["checked"] += 1
            if not ok:
                parity["mismatches"] += 1
                parity["detail"].append({"config": cfg_name, "rank": rank, "snapshot": s_i, "prn": int(prns[c_i]),
                                         "gpu": [float(table[s_i, c_i, 2]), float(table[s_i, c_i, 3])], "oracle": [o.code_phase, o.doppler_hz]})

    def synth_kernel_of(sprof):
        """dominant synthesis kernel of a generate call: name, its CUDA-event ms"""
        name = max(sprof, key=lambda k: sprof[k][0])
        return name, sprof[name][0]

    def measure_config(name, cfg, first, n, reps, warm, n_snap, prns, bins_step=DOPPLER_STEP, n_parity=4, main=False):
        """cold prologue, device-timed synthesis (CUDA events), acquisition over n_snap snapshots, e2e legs; this rank's numbers"""
        out = {}
        codes = np.stack([R.e1c_replica(p, 5e6, CODE_LENGTH) for p in prns])
        scen = R.GnssScenario(cfg, noise=True)
        scen.set_profiling(True)
        acq = R.PcpsAcquisition(CODE_LENGTH, 5e6).with_doppler_range(DOPPLER_MAX, bins_step)
        acq.set_profiling(True)
        bins = acq.num_doppler_bins()
        n_snap = min(n_snap, n // CODE_LENGTH)
        # ---- cold: a fresh handle pays for its block table / exact-phase prologue once (part of the path)
        barrier()
        tc = time.perf_counter()
        scen.generate_device(first, n, iq)
        torch.cuda.synchronize()
        cold_ms = (time.perf_counter() - tc) * 1e3

        def step(record):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            th = time.perf_counter()
            scen.generate_device(first, n, iq)
            th = time.perf_counter() - th
            e1.record()
            pods = acq.acquire_batch_raw(iq, n_snap, CODE_LENGTH, CODE_LENGTH, codes, prns)
            table = all_gather_table(results_to_table(pods, n_snap, len(prns))) if main else results_to_table(pods, n_snap, len(prns))
            e2.record()
            if record is not None:
                record.append((e0, e1, e2, acq.last_profile(), acq.guard_count(), scen.last_profile(), th))
            return table

        for _ in range(max(warm, 0)):
            step(None)
        barrier()
        launches0 = R.kernel_launches()
        clocks = ClockSampler(local) if (main and rank == 0) else None
        if clocks:
            clocks.start()
        rec = []
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record()
        table = None
        for _ in range(reps):
            table = step(rec)
        t1.record()
        barrier()
        out["clocks"] = clocks.stop() if clocks else None
        out["launches"] = R.kernel_launches() - launches0
        out["ms_syn"] = float(np.mean([r[0].elapsed_time(r[1]) for r in rec]))
        out["ms_acq"] = float(np.mean([r[1].elapsed_time(r[2]) for r in rec]))
        out["ms_total"] = t0.elapsed_time(t1) / reps
        out["cold_ms"] = cold_ms
        prof = rec[-1][3]
        sprof = {k: (float(np.mean([r[5][k][0] for r in rec])), rec[-1][5][k][1]) for k in rec[-1][5]}
        out["prof"], out["sprof"] = prof, sprof
        out["guards"] = int(np.sum([r[4] for r in rec]))
        out["host_enqueue_ms"] = float(np.mean([r[6] for r in rec]) * 1e3)
        kname, kms = synth_kernel_of(sprof)
        out["kernel"], out["kernel_ms"] = kname, kms
        if kname == "k_synth_periodic":      # it renders the whole primary-code periods of the range, k_synth the ragged ends
            k_lo, k_hi = max(1, -(-first // CODE_LENGTH)), (first + n) // CODE_LENGTH
            out["kernel_samples"] = (k_hi - k_lo) * CODE_LENGTH
        else:
            out["kernel_samples"] = n
        out["bins"], out["n_snap"], out["n"], out["first"] = bins, n_snap, n, first
        out["table"], out["codes"] = table, codes
        my_table = table[rank * n_snap:(rank + 1) * n_snap] if main else table      # every rank acquires the same number of snapshots
        check_parity(name, my_table, n_snap, codes, prns, acq, n_parity)
        if debug:
            print(f"[rank {rank}] {name}: synth {out['ms_syn']:.3f} ms (cold {cold_ms:.1f}), kernels {dict((k, round(v[0], 3)) for k, v in sprof.items())}, "
                  f"host enqueue {out['host_enqueue_ms']:.3f} ms, first {first}, n {n}", file=sys.stderr, flush=True)

        # ---- avg power line of the CLI (main.rs:4494-4509)
        scen.generate_device(first, n, iq)
        torch.cuda.synchronize()
        out["power"] = (scen.last_power_sum(), n)

        # ---- e2e: the same calls through the C-ABI with HOST buffers (pinned), copies inside the timed region
        ne = min(n, n_host)
        nse = min(n_snap, ne // CODE_LENGTH)
        scen.generate_range_into(first, ne, host.value)              # warm-up (staging buffers, page faults)
        acq.acquire_batch_raw(host_np[:ne], nse, CODE_LENGTH, CODE_LENGTH, codes, prns)
        barrier()
        e2e_reps = max(1, min(reps, 3))
        ts = time.perf_counter()
        for _ in range(e2e_reps):
            scen.generate_range_into(first, ne, host.value)
        torch.cuda.synchronize()
        out["ms_syn_e2e"] = (time.perf_counter() - ts) * 1e3 / e2e_reps
        ts = time.perf_counter()
        for _ in range(e2e_reps):
            pods = acq.acquire_batch_raw(host_np[:ne], nse, CODE_LENGTH, CODE_LENGTH, codes, prns)
            results_to_table(pods, nse, len(prns))
        torch.cuda.synchronize()
        out["ms_acq_e2e"] = (time.perf_counter() - ts) * 1e3 / e2e_reps
        out["n_e2e"], out["n_snap_e2e"] = ne, nse
        if main:
            fmt_e2e = {}
            for fname, code, bps in (("ci16", _lib.FMT_CI16, 4), ("ci8", _lib.FMT_CI8, 2)):
                scen.generate_range_into(first, ne, host.value, fmt=code)
                torch.cuda.synchronize()
                ts = time.perf_counter()
                scen.generate_range_into(first, ne, host.value, fmt=code)
                torch.cuda.synchronize()
                fmt_e2e[fname] = {"msamples_per_s": ne / (time.perf_counter() - ts) / 1e6, "d2h_bytes_per_step": ne * bps}
            out["fmt_e2e"] = fmt_e2e
            scen.generate_range_into(first, ne, host.value)          # host buffer back to cf32 (the CPU legs read it)
            torch.cuda.synchronize()
            # the reference's default Doppler grid (+-5 kHz / 500 Hz = 21 bins, acquisition.rs:63-74) on a 296-snapshot sample
            acq21 = R.PcpsAcquisition(CODE_LENGTH, 5e6)
            ns21 = min(n_snap, 296)
            acq21.acquire_batch_raw(iq, ns21, CODE_LENGTH, CODE_LENGTH, codes, prns)
            torch.cuda.synchronize()
            ts = time.perf_counter()
            acq21.acquire_batch_raw(iq, ns21, CODE_LENGTH, CODE_LENGTH, codes, prns)
            torch.cuda.synchronize()
            out["default_grid"] = {"bins": acq21.num_doppler_bins(), "snapshots": ns21,
                                   "cells_per_s": ns21 * len(prns) * acq21.num_doppler_bins() * CODE_LENGTH / (time.perf_counter() - ts),
                                   "note": "reference default +-5 kHz / 500 Hz, device-resident input, wall clock of one call"}
            acq21.close()
            # one reference-style call at a time (acquisition.rs:104: one input, one PRN; host Complex64 in, result out)
            x1 = iq[:CODE_LENGTH].cpu().numpy().astype(np.complex128)
            acq.acquire(x1, codes[0], prns[0])
            ts = time.perf_counter()
            for k in range(20):
                acq.acquire(x1, codes[k % len(prns)], prns[k % len(prns)])
            out["acquire_call_ms"] = (time.perf_counter() - ts) * 1e3 / 20
        scen.close(); acq.close()
        return out

    # ---- the main workload: K timed steps after W warm-up steps
    acq_total = args.acq_snapshots * world if args.acq_snapshots > 0 else ACQ_SNAPSHOTS_TOTAL
    main_prns = [s.prn for s in main_cfg.satellites]
    M = measure_config(WORKLOAD, main_cfg, first_main, n_main, args.steps, args.warmup, max(1, acq_total // world), main_prns, n_parity=16, main=True)

    # ---- the reference's own call pattern: `while !is_done { generate_block(block_size) }` (main.rs:4488-4500) through the C-ABI
    blk = None
    if rank == 0 and not args.no_block_api:
        bc = load_workload(1, None, CONFIGS[2]); bc.output.duration_s = args.block_api_seconds
        bsc = R.GnssScenario(bc, noise=True)
        nb = bsc.block_size()
        bsc.generate_block(nb); bsc.reset()                         # table built, ring allocated: the loop below is the steady state
        torch.cuda.synchronize()
        # the loop a Rust caller runs, with one preallocated block buffer, as a compiled caller of the C-ABI
        # (tools/ubench/block_loop.c; a ctypes call costs ~1 us of interpreter time, a third of the 40 KB block's copy)
        L_ = _lib.lib()
        h_ = bsc._h
        blkbuf = np.empty(nb, np.complex64)
        loop_so = os.path.join(ROOT, "tools", "ubench", "libblock_loop.so")
        BL = None
        try:
            if not os.path.exists(loop_so):          # normally built by __graft_entry__.build(); a measurement aid, not the product
                subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", loop_so, os.path.join(ROOT, "tools", "ubench", "block_loop.c")])
            BL = C.CDLL(loop_so).r4wb_block_loop
            BL.restype = C.c_int
            BL.argtypes = [C.c_void_p] * 4 + [C.c_uint64, C.c_void_p, C.c_int, C.c_uint64] + [C.POINTER(C.c_uint64)] * 2 + [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        except Exception as exc:                      # no compiled driver: the ctypes loop below still gives the line
            print(f"bench: compiled block-loop driver unavailable ({exc}); e2e_block_api falls back to the ctypes loop", file=sys.stderr)
        fp = lambda f: C.cast(f, C.c_void_p)

        def block_loop(mode):
            bsc.reset()
            got, calls, sec, fold = C.c_uint64(0), C.c_uint64(0), C.c_double(0), C.c_uint64(0)
            rc = BL(fp(L_.r4wb_scenario_generate_block), fp(L_.r4wb_scenario_generate_block_view), fp(L_.r4wb_scenario_is_done), h_,
                    nb, blkbuf.ctypes.data, mode, 8, C.byref(got), C.byref(calls), C.byref(sec), C.byref(fold))
            _lib.check(rc)
            return {"value": got.value / sec.value / 1e6, "unit": "Msamples/s", "samples": got.value, "ms": sec.value * 1e3,
                    "us_per_call": sec.value / max(1, calls.value) * 1e6}

        # ... and through ctypes, as round 1 and the first round-2 records measured it
        gb, done_f = L_.r4wb_scenario_generate_block, L_.r4wb_scenario_is_done
        pbuf, wr = C.c_void_p(blkbuf.ctypes.data), C.c_uint64(0)
        pwr = C.byref(wr)
        bsc.reset()
        ts = time.perf_counter()
        got = 0
        while not done_f(h_):
            gb(h_, nb, pbuf, _lib.MEM_HOST, _lib.FMT_CF32, pwr)
            got += wr.value
        dtb = time.perf_counter() - ts
        ctypes_leg = {"value": got / dtb / 1e6, "unit": "Msamples/s", "us_per_call": dtb / max(1, got // nb) * 1e6,
                      "api": "the same copying loop driven from Python (one ctypes call per block)"}
        if BL is not None:
            blk = block_loop(0)
            blk.update({"block_size": nb,
                        "api": f"r4wb_scenario_generate_block({nb}) until is_done from a compiled loop (tools/ubench/block_loop.c), {CONFIGS[2]} truncated to "
                               f"{args.block_api_seconds} s, host cf32 out; canonical blocks are served from the library's render-ahead ring (pinned host chunks), "
                               f"one host copy per block",
                        "view": dict(block_loop(1), api="r4wb_scenario_generate_block_view: the ring's pinned block is handed out, no host copy"),
                        "view_read": dict(block_loop(2), api="r4wb_scenario_generate_block_view and the consumer reads every byte of the block"),
                        "ctypes": ctypes_leg})
        else:
            blk = dict(ctypes_leg, samples=got, block_size=nb, ms=dtb * 1e3,
                       api=f"r4wb_scenario_generate_block({nb}) until is_done through ctypes, {CONFIGS[2]} truncated to {args.block_api_seconds} s, host cf32 out")
        bsc.close()

    # ---- all five BASELINE configs at the rank's share (bounded repetitions: 1 warm-up + 2 timed)
    per_config = {}
    if not args.no_per_config:
        for name in CONFIGS:
            if name == WORKLOAD:
                P = M
            else:
                cfg = load_workload(world, args, name)
                (f_c, n_c), _tot = share(cfg)
                prns_c = [s.prn for s in cfg.satellites]
                if name == CONFIGS[3]:
                    prns_c = list(range(1, 51))               # the full PRN x Doppler grid: 50 x 41 rows per snapshot
                P = measure_config(name, cfg, f_c, n_c, 2, 1, ACQ_SNAPSHOTS_CONFIG, prns_c)
            v = rmax([P["ms_syn"], P["ms_acq"], P["kernel_ms"], P["cold_ms"], P["ms_syn_e2e"], P["ms_acq_e2e"]])
            tot = rsum([P["n"], P["n_snap"], P["kernel_samples"], P["n_e2e"], P["n_snap_e2e"]])
            n_prn = P["codes"].shape[0]
            cells = tot[1] * n_prn * P["bins"] * CODE_LENGTH
            inv_fwd = rmax([P["prof"]["inverse_fft_peak"][0] + P["prof"]["forward_fft"][0]])[0]
            gbs = P["kernel_samples"] * 8 / (v[2] * 1e-3) / 1e9 if v[2] > 0 else None
            per_config[name] = {
                "samples": int(tot[0]), "synth_msamples_per_s": tot[0] / (v[0] * 1e-3) / 1e6, "ms_synth": v[0],
                "synth_kernel": P["kernel"], "roofline_frac": gbs / peak if gbs else None,
                "prologue_ms": max(0.0, v[3] - v[0]), "synth_msamples_per_s_cold": tot[0] / (v[3] * 1e-3) / 1e6,
                "e2e_msamples_per_s": tot[3] / (v[4] * 1e-3) / 1e6,
                "acq_prns": n_prn, "acq_bins": P["bins"], "acq_snapshots": int(tot[1]), "acq_cells_per_s": cells / (v[1] * 1e-3),
                "acq_roofline_frac": (P["n_snap"] * n_prn * P["bins"] * CODE_LENGTH * FLOP_PER_CELL / (inv_fwd * 1e-3) / 1e12 / fpeak) if inv_fwd > 0 else None,
                "acq_e2e_cells_per_s": tot[4] * n_prn * P["bins"] * CODE_LENGTH / (v[5] * 1e-3),
                "note": "rank's share of the config (time-sharded); 1 warm-up + 2 timed repetitions" if name != WORKLOAD else "the main workload (see top level)",
            }

    # ---- tracking channels, rank 0, outside the timed step
    track = None
    if rank == 0 and args.track:
        table = M["table"]
        e1c = np.stack([np.repeat(R.e1_code(1, p).astype(np.int8), 2) * np.tile(np.array([1, -1], np.int8), 4092) for p in main_prns])
        chans = [dict(prn=p, code_length=8184, sample_rate=5e6, chipping_rate=2.046e6,
                      initial_code_phase=float(((CODE_LENGTH - int(table[0, c, 2])) % CODE_LENGTH) * 2.046e6 / 5e6),
                      initial_doppler=float(table[0, c, 3])) for c, p in enumerate(main_prns)]
        n_per, n_p = CODE_LENGTH, 250
        R.TrackerBank(chans).process(iq, e1c, n_per, 8)            # warm-up
        bank = R.TrackerBank(chans)
        torch.cuda.synchronize()
        ts = time.perf_counter()
        st = bank.process(iq, e1c, n_per, n_p)
        dt = time.perf_counter() - ts
        track = {"metric": "tracking_channel_periods_per_s", "value": len(main_prns) * n_p / dt, "unit": "channel-periods/s",
                 "channels": len(main_prns), "periods": n_p, "samples_per_period": n_per, "ms": dt * 1e3,
                 "api": "r4wb_track_process(..., R4WB_MEM_DEVICE): states D2H inside the timed call",
                 "locked_channels": int(np.sum(st["code_lock"][-1]))}

    # ---- max over ranks
    ms_syn, ms_acq, ms_total, ms_syn_e2e, ms_acq_e2e, cold_ms, syn_kernel_ms = rmax(
        [M["ms_syn"], M["ms_acq"], M["ms_total"], M["ms_syn_e2e"], M["ms_acq_e2e"], M["cold_ms"], M["kernel_ms"]])
    inv_ms, fwd_ms = rmax([M["prof"]["inverse_fft_peak"][0], M["prof"]["forward_fft"][0]])
    total_samples, snaps_all, n_e2e_all, n_snap_e2e_all = rsum([M["n"], M["n_snap"], M["n_e2e"], M["n_snap_e2e"]])
    pw_sum, pw_cnt = all_reduce_power(*M["power"])
    par = rsum([parity["checked"], parity["mismatches"]])
    n, n_snap, bins, prns = M["n"], M["n_snap"], M["bins"], main_prns
    cells_rank = n_snap * len(prns) * bins * CODE_LENGTH
    cells = snaps_all * len(prns) * bins * CODE_LENGTH

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        cores = os.cpu_count() or 1
        th_s = max(1, min(len(prns), cores))
        v_s, n_s, dt_s = cpu_synth(main_cfg, 3.0, th_s)                    # 3 s of the scenario: ~10-15 s of host work
        v_1, n_1, dt_1 = cpu_synth(main_cfg, 0.3, 1)                       # the reference's default build is single-threaded
        ns_c = max(1, min(n_snap, 8 * cores))                              # ~10 s of host work on all cores
        v_a, c_a, dt_a = cpu_acq(host_np, M["codes"], prns, ns_c, cores)
        cpu = ({"value": v_s, "unit": "Msamples/s", "cores": th_s, "kind": "port",
                "sample": f"first 3 s of {WORKLOAD} ({n_s} samples, {dt_s:.1f} s wall), oracle port, one thread per satellite",
                "single_thread": {"value": v_1, "unit": "Msamples/s", "cores": 1,
                                  "sample": f"0.3 s of {WORKLOAD} ({n_1} samples, {dt_1:.1f} s wall), oracle port, 1 thread"}},
               {"value": v_a, "unit": "cells/s", "cores": cores, "kind": "port",
                "sample": f"{ns_c} snapshot(s) x {len(prns)} PRNs x {bins} bins ({c_a} cells, {dt_a:.1f} s wall), oracle port, {cores} threads"})
    _lib.check(_lib.lib().r4wb_host_free(host))

    if rank == 0:
        synth_gbs = M["kernel_samples"] * 8 / (syn_kernel_ms * 1e-3) / 1e9
        acq_tflops = cells_rank * FLOP_PER_CELL / ((inv_ms + fwd_ms) * 1e-3) / 1e12 if inv_ms + fwd_ms > 0 else None
        table = M["table"]
        line = {
            "metric": "gnss_iq_synth_msamples_per_s", "value": total_samples / (ms_syn * 1e-3) / 1e6, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_syn, "higher_is_better": True,
            "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD} ({main_cfg.output.duration_s:g} s, {int(total_samples)} samples) time-sharded over {world} GPU(s): "
                                   f"{SEGMENT_S:g} s segment per GPU, {len(prns)} Galileo E1C PRNs, 5 MS/s, noise on, cf32 into HBM; then PCPS over "
                                   f"{n_snap} snapshots/GPU x {len(prns)} PRNs x {bins} Doppler bins x {CODE_LENGTH} lags",
                       "samples_per_gpu": n, "snapshots_per_gpu": n_snap, "prns": prns,
                       "l2": f"per-step output {n * 8 / 1e9:.1f} GB and spectra working set exceed the 126 MB L2 (no explicit flush)",
                       "step": "synth then acquire; ms_per_step/value cover the synthesis half, acq.* the acquisition half, ms_step_total both"},
            "ms_step_total": ms_total,
            "prologue_ms": max(0.0, cold_ms - ms_syn), "value_cold": total_samples / (cold_ms * 1e-3) / 1e6,
            "prologue_note": "first generate call of a fresh handle (block table, exact-phase scan, tile records + render), wall clock, max "
                             "over ranks, minus the warm render time; value_cold = samples / that first call",
            "roofline": {"bound": "hbm", "kernel": M["kernel"], "achieved": synth_gbs, "peak": peak, "unit": "GB/s", "frac": synth_gbs / peak,
                         "traffic": ncu_traffic(M["kernel"]), "peak_source": peak_src, "algorithmic_bytes_per_launch": M["kernel_samples"] * 8,
                         "kernel_ms": syn_kernel_ms,
                         "note": "8 B per output sample (one cf32 store) x the samples this kernel wrote / its CUDA-event time on the launching "
                                 "stream; `value` is over the whole generate call (all kernels + launch gaps)"},
            "synth_kernel_ms": {k: v[0] for k, v in M["sprof"].items()}, "synth_kernel_launches": {k: v[1] for k, v in M["sprof"].items()},
            "e2e": {"value": n_e2e_all / (ms_syn_e2e * 1e-3) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": M["n_e2e"] * 8, "samples_per_gpu": M["n_e2e"],
                    "api": "r4wb_scenario_generate(..., R4WB_MEM_HOST, CF32) into pinned host memory",
                    "other_formats_rank0": M.get("fmt_e2e")},
            "acq": {"metric": "pcps_acq_cells_per_s", "value": cells / (ms_acq * 1e-3), "unit": "cells/s", "ms_per_step": ms_acq,
                    "cells_per_step": cells, "f64_guard_reruns": M["guards"],
                    "kernel_ms": {k: v[0] for k, v in M["prof"].items()}, "kernel_launches": {k: v[1] for k, v in M["prof"].items()},
                    "roofline": {"bound": "fp32", "kernel": "k_rf_inv_peak_tm (+k_rf_fwd)", "achieved": acq_tflops, "peak": fpeak,
                                 "unit": "TFLOP/s", "frac": (acq_tflops / fpeak) if acq_tflops else None, "traffic": ncu_traffic("k_rf_inv_peak_tm"),
                                 "peak_source": fpeak_src, "nominal_peak": FP32_PEAK_TFLOPS,
                                 "note": f"{FLOP_PER_CELL} reference-equivalent flop per cell / summed CUDA-event time of the forward and inverse FFT kernels"},
                    "e2e": {"value": n_snap_e2e_all * len(prns) * bins * CODE_LENGTH / (ms_acq_e2e * 1e-3), "unit": "cells/s",
                            "h2d_bytes_per_step": M["n_snap_e2e"] * CODE_LENGTH * 8, "d2h_bytes_per_step": M["n_snap_e2e"] * len(prns) * 32,
                            "snapshots_per_gpu": M["n_snap_e2e"],
                            "api": "r4wb_pcps_acquire_batch(..., R4WB_MEM_HOST) from pinned host memory"},
                    "acquire_call_ms": M.get("acquire_call_ms"),
                    "first_snapshot": [[int(table[0, c, 2]), float(table[0, c, 3])] for c in range(len(prns))],
                    "default_grid_rank0": M.get("default_grid")},
            "parity_checked": int(par[0]), "parity_mismatches": int(par[1]), "parity_detail_rank0": parity["detail"][:8],
            "parity_note": "random (snapshot, PRN) results of every acquisition leg vs the oracle's acquire on the same device samples: "
                           "(lag, Doppler, detected) must be identical",
            "avg_power": {"value": pw_sum / max(pw_cnt, 1), "samples": pw_cnt, "note": "sum |s|^2 / count all-reduced over the ranks (the CLI's avg-power line)"},
            "gpu_launches": int(M["launches"]),
            "clocks": M["clocks"],
            "numa_rank0": numa,
        }
        if per_config:
            line["per_config"] = per_config
        if blk is not None:
            line["e2e_block_api"] = blk
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
    ap.add_argument("--acq-snapshots", type=int, default=0, help=f"snapshots per GPU per step (0 = {ACQ_SNAPSHOTS_TOTAL} / N)")
    ap.add_argument("--workload", default="", help="another config of BASELINE.json (file name under configs/); default " + WORKLOAD)
    ap.add_argument("--weak", action="store_true", help="one copy of the config's duration per GPU instead of splitting it")
    ap.add_argument("--strong", action="store_true", help="(default) split the config's own duration over the ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="skip the per_config table (the other four BASELINE configs)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of random acquisition results")
    ap.add_argument("--no-block-api", action="store_true", help="skip the generate_block loop leg")
    ap.add_argument("--block-api-seconds", type=float, default=10.0)
    ap.add_argument("--no-numa-bind", action="store_true", help="leave CPU affinity / memory policy alone")
    ap.add_argument("--track", action="store_true", help="add the tracking-channel leg (SURVEY.md section 8 f2)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
