"""Regenerates tests/golden/track_case.npz from the oracle (run in the build container): 20 ms of a noisy GPS L1 C/A signal
(cf32), the C/A code, the channel's start values and the TrackingState the oracle's TrackingChannel::process returns for each
of the 20 code periods.  The reference has no test of `process`, so these are oracle outputs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from tests.test_tracking import _gps_signal, _code_phase_of

x, _ = _gps_signal(O, 20, prn=11, doppler=-1530.0, cn0=48.0, nav=True, seed=7)
cp0, dop0 = _code_phase_of(O, x, 11)
code = O.gps_ca_code(11).astype(np.int8)
states = O.OracleTrackingChannel(11, 1023, 5e6, 1.023e6, cp0, dop0).with_dll_bandwidth(2.0).run(x, code, 5000, 20)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "track_case.npz"), x=x, code=code, start=np.array([cp0, dop0]),
                    states=states)
print("golden written", x.shape, states.shape, states["carrier_freq_hz"][-1])
