"""Multi-GPU inside the library (r4wb_init_devices): the host-buffer batch calls shard over the devices of one process —
no torch.distributed, no NCCL — and return exactly what one device returns.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import numpy as np
import pytest

from tests.conftest import config_path
from r4w_b200.dist import results_to_table

pytestmark = pytest.mark.gpu


def _cfg(name):
    from r4w_b200.config import load_config
    return load_config(config_path(name), cli_elevation_mask_deg=5.0)


def test_sharded_batch_calls_equal_single_device(gpu):
    if gpu.device_count() < 2:
        pytest.skip("one visible GPU")
    cfg = _cfg("e1c_8prn_60s_cn34_orbital")
    first, n, n_snap = 3_000_000, 2_000_000, 100
    prns = [s.prn for s in cfg.satellites]
    codes = np.stack([gpu.e1c_replica(p, 5e6, 20000) for p in prns])
    try:
        assert gpu.init_devices(1) == 1
        sc = gpu.GnssScenario(cfg, noise=True)
        one = sc.generate_range(first, n)
        p_one = sc.last_power_sum()
        acq = gpu.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
        t_one = results_to_table(acq.acquire_batch_raw(one, n_snap, 20000, 20000, codes, prns), n_snap, len(prns))
        nd = gpu.init_devices(0)
        assert nd >= 2
        sc2 = gpu.GnssScenario(cfg, noise=True)
        many = sc2.generate_range(first, n)
        assert np.array_equal(one, many)                              # every sample is a function of (config, global index)
        assert sc2.last_power_sum() == pytest.approx(p_one, rel=1e-6)
        acq2 = gpu.PcpsAcquisition(20000, 5e6).with_doppler_range(5000.0, 250.0)
        t_many = results_to_table(acq2.acquire_batch_raw(one, n_snap, 20000, 20000, codes, prns), n_snap, len(prns))
        assert np.array_equal(t_one, t_many, equal_nan=True)
        # device-resident input on device 0: the other devices fetch their share over NVLink
        import torch
        d_in = torch.from_numpy(one).to("cuda:0")
        t_dev = results_to_table(acq2.acquire_batch_raw(d_in, n_snap, 20000, 20000, codes, prns), n_snap, len(prns))
        assert np.array_equal(t_one, t_dev, equal_nan=True)
    finally:
        gpu.init_devices(1)
