"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed for the plumbing).

Both halves of the path shard into independent units: synthesis by time segment — every output
sample is a pure function of (config, global sample index) — and acquisition by snapshot.  No data-path
collective exists; the only exchange is ONE all-gather of the tiny per-(snapshot, PRN) peak table at the end
of a batch (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def segment_for_rank(total_samples: int, align: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous time segment [first, first+n) of rank `rank`: whole multiples of `align` samples (a 1 ms
    synthesis block or a 4 ms acquisition snapshot), remainder spread over the first ranks, tail to the last."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    units = total_samples // align
    base, extra = divmod(units, world)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    first, end = u0 * align, u1 * align
    if rank == world - 1:
        end = total_samples
    return first, end - first


def snapshots_for_rank(n_snapshots: int, rank: int, world: int) -> Tuple[int, int]:
    """(first_snapshot, count) of rank `rank`."""
    first, n = segment_for_rank(n_snapshots, 1, rank, world)
    return first, n


def all_gather_table(local: np.ndarray, counts: List[int] = None) -> np.ndarray:
    """All-gather rows of a float64 table whose leading dimension differs per rank.  Returns the concatenation
    in rank order on every rank.  Uses the default process group (nccl -> staged through the rank's GPU,
    gloo -> CPU)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.ascontiguousarray(local, np.float64)
    world = dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    loc = torch.from_numpy(np.ascontiguousarray(local, np.float64)).to(dev)
    n_local = torch.tensor([loc.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(ns, n_local)
    ns = [int(x.item()) for x in ns]
    n_max = max(ns) if ns else 0
    row_shape = tuple(loc.shape[1:])
    pad = torch.zeros((n_max,) + row_shape, dtype=torch.float64, device=dev)
    pad[: loc.shape[0]] = loc
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.cat([p[:n] for p, n in zip(parts, ns)], dim=0)
    return out.cpu().numpy()


def all_reduce_power(power_sum: float, n_samples: int) -> Tuple[float, int]:
    """The CLI's average-power line over a time-sharded render (crates/r4w-cli/src/main.rs:4494-4509: sum |s|^2 / count):
    one all-reduce of {sum |s|^2, count} over the ranks.  Returns (total sum, total count)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(power_sum), int(n_samples)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(power_sum), float(n_samples)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0].item()), int(round(float(t[1].item())))


ACQ_DTYPE = np.dtype([("prn", "u1"), ("detected", "u1"), ("has_cn0", "u1"), ("_pad", "u1", (5,)), ("code_phase", "f8"),
                      ("doppler_hz", "f8"), ("peak_metric", "f8"), ("threshold", "f8"), ("cn0_estimate", "f8")])


def results_to_table(pods, n_snapshots: int, n_codes: int) -> np.ndarray:
    """r4wb_acq_result array -> [n_snapshots][n_codes][6] float64: prn, detected, code_phase, doppler_hz,
    peak_metric, cn0_estimate (nan = None)."""
    a = np.frombuffer(pods, dtype=ACQ_DTYPE, count=n_snapshots * n_codes)
    t = np.empty((n_snapshots * n_codes, 6), np.float64)
    t[:, 0] = a["prn"]; t[:, 1] = a["detected"]; t[:, 2] = a["code_phase"]; t[:, 3] = a["doppler_hz"]
    t[:, 4] = a["peak_metric"]
    t[:, 5] = np.where(a["has_cn0"] != 0, a["cn0_estimate"], np.nan)
    return t.reshape(n_snapshots, n_codes, 6)
