"""r4w_b200 — B200 (sm_100a) drop-in for r4w's GNSS hot path: multi-satellite IQ scenario synthesis and
FFT-based PCPS acquisition.  Host-side mirror of the reference's `GnssScenarioConfig` / `GnssScenario` /
`PcpsAcquisition` surface over the C-ABI library libr4w_b200.so (include/r4w_b200.h)."""
from .config import (GnssScenarioConfig, SatelliteConfig, ReceiverConfig, EnvironmentConfig, OutputConfig,  # noqa: F401
                     LlaPosition, AntennaPattern, ReceiverTrajectory, load_config, loads_config, preset_config, PRESETS,
                     gps_time_from_utc)
from ._lib import R4wB200Error, init, init_devices, kernel_launches, device_count, version, build  # noqa: F401
from .scenario import GnssScenario, SatelliteStatus  # noqa: F401
from .tracking import TrackingChannel, TrackerBank, TrackingState  # noqa: F401
from .acquisition import (PcpsAcquisition, AcquisitionResult, AcquisitionGrid, e1_code, e1c_secondary,  # noqa: F401
                          e1c_replica, gps_ca_code, gps_l5_code, glonass_code)
