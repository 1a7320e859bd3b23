"""The C-ABI boundary: libr4w_b200.so loads, exports every symbol include/r4w_b200.h declares, mirrors r4w-ffi's
error conventions, and refuses to compute without a CUDA device (no CPU fallback).  No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from r4w_b200 import _lib
    _lib.build()
    return _lib.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "r4w_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(r4wb_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(L):
    from r4w_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/r4w_b200.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype in r4w_b200/_lib.py"


def test_error_enum_matches_r4w_ffi():
    """values 0-7 are r4w.h's R4wError (crates/r4w-ffi/include/r4w.h:8-25)"""
    text = open(os.path.join(ROOT, "include", "r4w_b200.h")).read()
    for name, val in [("OK", 0), ("ERR_NULL_POINTER", 1), ("ERR_INVALID_SIZE", 2), ("ERR_BUFFER_FULL", 3), ("ERR_BUFFER_EMPTY", 4),
                      ("ERR_INVALID_PARAMETER", 5), ("ERR_ALLOCATION_FAILED", 6), ("ERR_NOT_SUPPORTED", 7), ("ERR_CUDA", 100)]:
        assert re.search(rf"R4WB_{name}\s*=\s*{val}\b", text)


def test_pod_layouts_match_header():
    from r4w_b200 import config as K
    assert C.sizeof(K.SatCfgPod) == 96 and C.sizeof(K.LlaPod) == 24
    assert C.sizeof(K.OutputCfgPod) == 48 and C.sizeof(K.EnvironmentCfgPod) == 16 + 12 * 8
    assert C.sizeof(K.AcqResultPod) == 48 and C.sizeof(K.SatStatusPod) == 88
    assert C.sizeof(K.ReceiverCfgPod) == 24 + 8 + 5 * 8 + 48 + 8 + 8


def test_version_and_null_safety(L):
    assert L.r4wb_version().decode().startswith("r4w_b200")
    L.r4wb_scenario_destroy(None)           # NULL-safe, like r4w_*_free (r4w-ffi/src/lib.rs:272-277)
    L.r4wb_pcps_destroy(None)
    assert L.r4wb_scenario_total_samples(None) == 0 and L.r4wb_scenario_is_done(None) == 1
    assert L.r4wb_pcps_fft_size(None) == 0
    assert L.r4wb_scenario_reset(None) == 1   # NullPointer


def test_host_side_code_tables(L, oracle):
    """r4wb_e1_code / r4wb_e1c_secondary / r4wb_e1c_replica are table unpacks on the host: bit-exact vs the oracle"""
    import r4w_b200 as R
    for prn in (1, 3, 25, 50):
        for ch in (0, 1):
            assert np.array_equal(R.e1_code(ch, prn), oracle.e1_code(ch, prn))
        assert np.array_equal(R.e1c_replica(prn, 5e6, 20000), oracle.e1c_replica(prn, 5e6, 20000))
    assert np.array_equal(R.e1c_secondary(), oracle.e1c_secondary())
    assert np.array_equal(R.e1c_replica(7, 4.092e6, 9000), oracle.e1c_replica(7, 4.092e6, 9000))
    with pytest.raises(R.R4wB200Error):
        R.e1_code(1, 0)
    with pytest.raises(R.R4wB200Error):
        R.e1_code(2, 1)


def test_no_cpu_fallback(L):
    """without a CUDA device every compute entry point fails loudly with R4WB_ERR_CUDA"""
    import r4w_b200 as R
    if R.device_count() > 0:
        pytest.skip("a GPU is present")
    from tests.conftest import config_path
    with pytest.raises(R.R4wB200Error) as e:
        R.GnssScenario.from_yaml(config_path("e1c_prn3_20s_withdoppler"))
    assert e.value.code == 100 and "no CPU fallback" in str(e.value)
    with pytest.raises(R.R4wB200Error) as e:
        R.PcpsAcquisition(1023, 1023.0)
    assert e.value.code == 100


def test_product_does_not_import_oracle():
    """nothing under r4w_b200/ may reference oracle/ or tests/emu/"""
    pkg = os.path.join(ROOT, "r4w_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dp, f)).read()
                assert "from oracle" not in text and "import oracle" not in text and "libr4w_oracle" not in text, f
                assert "r4w_oracle" not in text and "libr4w_emu" not in text, f


def test_gps_ca_codes_match_oracle(L, oracle):
    """the product's own Gold-code generator (synth_model.cpp) against the oracle's restatement of gnss/prn.rs:34-162"""
    import r4w_b200 as R
    for prn in range(1, 33):
        assert np.array_equal(R.gps_ca_code(prn), oracle.gps_ca_code(prn))
    with pytest.raises(R.R4wB200Error):
        R.gps_ca_code(33)


def test_gps_l5_and_glonass_codes_match_oracle(L, oracle):
    """the product's own shift-register generators against the oracle's restatement of GpsL5CodeGenerator / GlonassCodeGenerator
    (gnss/prn.rs:170-216, 345-397); reference KATs: test_glonass_code_length (511), test_gps_l5_code_length (10230),
    test_gps_l5_iq_different; plus the m-sequence property of the GLONASS code (autocorrelation -1 off the peak)"""
    import r4w_b200 as R
    g = R.glonass_code()
    assert g.size == 511 and np.array_equal(g, oracle.glonass_code(0)) and np.array_equal(g, oracle.glonass_code(6))
    gi = g.astype(np.int64)
    assert int(gi @ gi) == 511 and all(int(gi @ np.roll(gi, k)) == -1 for k in (1, 7, 255, 510))
    for prn in (1, 7, 32):
        c = R.gps_l5_code(prn)
        assert c.size == 10230 and np.array_equal(c, oracle.gps_l5_code(prn))
    assert not np.array_equal(oracle.gps_l5_code(1), oracle.gps_l5_code(1, q_channel=True))
    assert not np.array_equal(R.gps_l5_code(1), R.gps_l5_code(2))
    with pytest.raises(R.R4wB200Error):
        R.gps_l5_code(33)
    with pytest.raises(ValueError):
        oracle.glonass_code(7)
