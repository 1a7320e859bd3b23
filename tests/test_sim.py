"""r4w-sim's generic scenario engine (crates/r4w-sim/src/scenario/): the reference's own tests restated
against the oracle and the Python host mirror (CPU), and parity of the GPU composer with the oracle (pytest -m gpu)."""
import math

import numpy as np
import pytest


def _mirror():
    from r4w_b200 import sim
    return sim


def test_config_kats(oracle):
    """config.rs test_default_config / test_noise_power"""
    sim = _mirror()
    cfg = sim.ScenarioConfig()
    assert cfg.total_samples() == 2046 and cfg.num_blocks() == 1
    n = cfg.noise_power_linear()
    assert 0.0 < n < 1e-10 and n == pytest.approx(oracle.sim_noise_power(-204.0, 2_046_000.0), rel=1e-15)


def test_trajectory_kats_and_mirror(oracle):
    """trajectory.rs test_static_trajectory / test_linear_trajectory / test_circular_trajectory on the oracle, and the host
    mirror (r4w_b200/sim.py) equal to the oracle for all four variants"""
    sim = _mirror()
    s0, s1 = oracle.sim_trajectory("Static", (40.0, -75.0, 100.0), 0.0), oracle.sim_trajectory("Static", (40.0, -75.0, 100.0), 10.0)
    assert abs(s0[0] - s1[0]) < 1e-6 and np.linalg.norm(s0[3:]) < 1e-10
    l0, l1 = oracle.sim_trajectory("Linear", (0.0, 0.0, 0.0, 10.0, 0.0, 0.0), 0.0), oracle.sim_trajectory("Linear", (0.0, 0.0, 0.0, 10.0, 0.0, 0.0), 1.0)
    assert abs(np.linalg.norm(l1[:3] - l0[:3]) - 10.0) < 0.1
    c = oracle.sim_trajectory("Circular", (0.0, 0.0, 0.0, 100.0, 0.1, 0.0), 0.0)
    assert abs(np.linalg.norm(c[3:]) - 10.0) < 0.5
    pts = [(0.0, (41.0, -85.0, 200.0)), (10.0, (41.01, -85.02, 250.0)), (25.0, (41.05, -85.0, 300.0))]
    cases = [(sim.Trajectory("Static", position=(40.0, -75.0, 100.0)), "Static", (40.0, -75.0, 100.0)),
             (sim.Trajectory("Linear", position=(41.08, -85.14, 240.0), velocity_enu=(30.0, -12.0, 1.5)), "Linear", (41.08, -85.14, 240.0, 30.0, -12.0, 1.5)),
             (sim.Trajectory("Circular", position=(10.0, 20.0, 1000.0), radius_m=500.0, omega_rad_s=0.2, initial_bearing_deg=30.0), "Circular",
              (10.0, 20.0, 1000.0, 500.0, 0.2, 30.0)),
             (sim.Trajectory("Waypoints", points=pts), "Waypoints", [(t, *p) for t, p in pts])]
    for traj, kind, params in cases:
        for t in (-1.0, 0.0, 3.7, 10.0, 17.25, 25.0, 40.0):
            st, want = traj.state_at(t), oracle.sim_trajectory(kind, params, t)
            assert np.allclose(st.position, want[:3], rtol=0, atol=1e-8) and np.allclose(st.velocity, want[3:], rtol=0, atol=1e-10), (kind, t)


class ToneEmitter:
    """the reference's test emitter (engine.rs:217-244): a 1 kHz complex tone from a fixed position"""

    def __init__(self, position, freq_hz, power_dbm, tone_hz=1000.0, velocity=(0.0, 0.0, 0.0), name="tone", on_after=-1.0):
        self.position, self.freq_hz, self.power_dbm, self.tone_hz = np.asarray(position, float), freq_hz, power_dbm, tone_hz
        self.velocity, self.name, self.on_after = np.asarray(velocity, float), name, on_after

    def state_at(self, t):
        from r4w_b200.sim import EmitterState
        return EmitterState(self.position + self.velocity * t, self.velocity, self.power_dbm, t >= self.on_after)

    def generate_iq(self, t, num_samples, sample_rate):
        ts = t + np.arange(num_samples) / sample_rate
        ph = 2.0 * math.pi * self.tone_hz * ts
        return np.cos(ph) + 1j * np.sin(ph)

    def carrier_frequency_hz(self): return self.freq_hz
    def nominal_power_dbm(self): return self.power_dbm
    def id(self): return self.name


def _oracle_engine(oracle, sim, cfg, emitters, traj):
    """the reference's generate_all with the oracle's pieces (geometry, compose), noise-free"""
    comp = oracle.OracleComposer(len(emitters), cfg.sample_rate)
    out, cur = [], 0
    kind = traj.kind
    params = {"Static": traj.position, "Linear": (*traj.position, *traj.velocity_enu),
              "Circular": (*traj.position, traj.radius_m, traj.omega_rad_s, traj.initial_bearing_deg),
              "Waypoints": [(t, *p) for t, p in traj.points]}[kind]
    while cur < cfg.total_samples():
        n = min(cfg.total_samples() - cur, cfg.block_size)
        t_start = cur / cfg.sample_rate
        t_mid = t_start + (n / 2.0) / cfg.sample_rate
        rx = oracle.sim_trajectory(kind, params, t_mid)
        bb = np.zeros((len(emitters), n), np.complex128)
        dop, amp, act = np.zeros(len(emitters)), np.zeros(len(emitters)), np.zeros(len(emitters), bool)
        for k, em in enumerate(emitters):
            st = em.state_at(t_mid)
            if not st.active:
                continue
            _, dop[k], _, amp[k] = oracle.sim_link(rx, np.concatenate([st.position, st.velocity]), em.carrier_frequency_hz(), st.power_dbm)
            bb[k] = em.generate_iq(t_start, n, cfg.sample_rate)
            act[k] = True
        out.append(comp.block(bb, dop, amp, act))
        cur += n
    return np.concatenate(out), comp.phases


@pytest.mark.gpu
def test_engine_generates_samples(gpu):
    """engine.rs test_engine_generates_samples: 1 ms at 10 kHz in blocks of 100 -> 10 samples, done"""
    sim = _mirror()
    cfg = sim.ScenarioConfig(duration_s=0.001, sample_rate=10000.0, center_frequency_hz=1e9, block_size=100, noise_floor_dbw_hz=-250.0, seed=42)
    em = ToneEmitter(sim.lla_to_ecef(0.0, 0.0, 20_200_000.0), 1e9, 50.0)
    eng = sim.ScenarioEngine(cfg, [em], sim.Trajectory("Static", position=(0.0, 0.0, 0.0)))
    x = eng.generate_all()
    assert x.size == 10 and eng.is_done() and eng.progress() == 1.0
    st = eng.emitter_status(0.0)[0]
    assert st.id == "tone" and st.range_m == pytest.approx(20_200_000.0, rel=1e-9) and st.doppler_hz == 0.0
    assert np.allclose(np.abs(x), 10.0 ** ((st.received_power_dbm - 30.0) / 20.0), rtol=1e-4)


@pytest.mark.gpu
def test_engine_matches_oracle(gpu, oracle):
    """three emitters (one moving, one switching on mid-run) seen from a circling receiver, 0.2 s at 1 MHz in blocks of 4 000:
    noise-free composite within 1e-5 relative RMS of the oracle, carrier phases equal, reset reproduces the stream"""
    sim = _mirror()
    cfg = sim.ScenarioConfig(duration_s=0.2, sample_rate=1_000_000.0, center_frequency_hz=1.5e9, block_size=4000, noise_floor_dbw_hz=-204.0, seed=7)
    ems = [ToneEmitter(sim.lla_to_ecef(0.0, 0.0, 20_200_000.0), 1.5e9, 50.0, 1000.0, name="a"),
           ToneEmitter(sim.lla_to_ecef(5.0, 3.0, 400_000.0), 1.5e9, 30.0, -23_000.0, velocity=(0.0, 7500.0, 100.0), name="leo"),
           ToneEmitter(sim.lla_to_ecef(0.1, 0.1, 50.0), 1.5e9, 10.0, 7_000.0, name="late", on_after=0.1)]
    traj = sim.Trajectory("Circular", position=(0.0, 0.0, 10_000.0), radius_m=2000.0, omega_rad_s=0.3, initial_bearing_deg=10.0)
    eng = sim.ScenarioEngine(cfg, ems, traj, noise=False)
    got = eng.generate_all()
    want, phases = _oracle_engine(oracle, sim, cfg, ems, traj)
    assert got.size == want.size == 200_000
    assert np.sqrt(np.sum(np.abs(got - want) ** 2) / np.sum(np.abs(want) ** 2)) <= 1e-5
    d = np.abs(eng.carrier_phases() - phases)
    assert np.all(np.minimum(d, np.abs(d - 2 * np.pi)) < 1e-6)
    eng.reset()
    assert np.array_equal(eng.generate_all(), got)
    # receiver noise: N(0, noise_power / 2) per component, white
    noisy = sim.ScenarioEngine(cfg, ems, traj, noise=True).generate_all()
    w = noisy - got
    sd = math.sqrt(cfg.noise_power_linear() / 2.0)
    assert abs(w.real.std() / sd - 1.0) < 0.01 and abs(w.imag.std() / sd - 1.0) < 0.01
    assert abs(np.mean(w[1:] * np.conj(w[:-1]))) / (2 * sd * sd) < 0.01


@pytest.mark.gpu
def test_engine_phase_wrap(gpu, oracle):
    """the `%= 2 pi` wrap of the accumulated carrier phase once it exceeds 1e6 rad (engine.rs:111-113): 4 kHz of Doppler at a
    10 kHz sample rate gets there after 4e5 samples"""
    sim = _mirror()
    cfg = sim.ScenarioConfig(duration_s=60.0, sample_rate=10_000.0, center_frequency_hz=1.5e9, block_size=50_000, noise_floor_dbw_hz=-204.0, seed=1)
    v = 4000.0 * sim.SPEED_OF_LIGHT / 1.5e9                          # closing speed for +4 kHz
    em = ToneEmitter(sim.lla_to_ecef(0.0, 0.0, 30_000_000.0), 1.5e9, 60.0, 100.0, velocity=(-v, 0.0, 0.0))
    traj = sim.Trajectory("Static", position=(0.0, 0.0, 0.0))
    eng = sim.ScenarioEngine(cfg, [em], traj, noise=False)
    got = eng.generate_all()
    want, phases = _oracle_engine(oracle, sim, cfg, [em], traj)
    assert abs(phases[0]) < 1.0e6 and np.sqrt(np.sum(np.abs(got - want) ** 2) / np.sum(np.abs(want) ** 2)) <= 1e-5
    # the reference adds 6e5 increments one by one at a magnitude of up to 1e6 rad (ulp 1.2e-10): its own rounding drift
    d = abs(eng.carrier_phases()[0] - phases[0]) % (2 * np.pi)
    assert min(d, 2 * np.pi - d) < 1e-4
