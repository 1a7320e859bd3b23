/*
 * r4w_oracle_track.c — CPU restatement (f64, plain C) of r4w's tracking channel.
 * TEST INFRASTRUCTURE ONLY (see r4w_oracle.h).  Follows crates/r4w-core/src/waveform/gnss/tracking.rs line by line:
 *   TrackingChannel::new / with_dll_bandwidth / with_pll_bandwidth   :107-167
 *   TrackingChannel::process                                         :177-313
 *   estimate_cn0 (Beaulieu moment estimator over the last 20 prompts) :316-337
 *   LoopFilter2nd :365-397, LoopFilter3rd :399-437, dll_s_curve :441-456
 * Rust semantics kept: f64::rem_euclid (r = fmod; r < 0 -> r + |rhs|), `as usize` truncation (negative -> 0),
 * sums in sample order, sin_cos of -2 pi (f t + phase).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "r4w_oracle.h"

#define ORC_PI 3.14159265358979323846

typedef struct { double k1, k2, integrator; } loop2;
typedef struct { double k1, k2, k3, i1, i2; } loop3;

static loop2 loop2_new(double bw, double T)
{
    const double omega_n = bw * 8.0 / 3.0, zeta = 1.0 / sqrt(2.0);
    loop2 f = {2.0 * zeta * omega_n * T, (omega_n * omega_n) * (T * T), 0.0};
    return f;
}
static double loop2_update(loop2* f, double d)
{
    f->integrator += f->k2 * d;
    return f->k1 * d + f->integrator;
}
static loop3 loop3_new(double bw, double T)
{
    const double omega_n = bw * 2.4, a3 = 1.1, b3 = 2.4;
    loop3 f = {b3 * omega_n * T, a3 * (omega_n * omega_n) * (T * T), (omega_n * omega_n * omega_n) * (T * T * T), 0.0, 0.0};
    return f;
}
static double loop3_update(loop3* f, double d)
{
    f->i1 += f->k2 * d;
    f->i2 += f->k3 * d;
    return f->k1 * d + f->i1 + f->i2;
}

struct orc_tracker {
    uint8_t prn;
    size_t code_length;
    double sample_rate, chipping_rate;
    double code_phase, code_freq, el_spacing;
    loop2 dll;
    double carrier_phase, carrier_freq, fll_bandwidth;
    loop3 pll;
    int fll_assist;
    double e_i, e_q, p_i, p_q, l_i, l_q;
    double cn0_buf[20];
    size_t cn0_n;
    int carrier_lock, code_lock, bit_sync;
    uint64_t ms_count;
    double nav_acc;
    uint32_t nav_count;
    int8_t* nav_bits;
    size_t nav_n, nav_cap;
    int prev_sign;
};

orc_tracker* orc_track_new(uint8_t prn, size_t code_length, double sample_rate, double chipping_rate, double initial_code_phase,
                           double initial_doppler)
{
    orc_tracker* t = (orc_tracker*)calloc(1, sizeof *t);
    if (!t) return NULL;
    const double code_doppler = initial_doppler * chipping_rate / 1575420000.0;       /* tracking.rs:122 */
    t->prn = prn; t->code_length = code_length; t->sample_rate = sample_rate; t->chipping_rate = chipping_rate;
    t->code_phase = initial_code_phase; t->code_freq = chipping_rate + code_doppler;
    t->el_spacing = 0.5; t->dll = loop2_new(1.0, 0.001);
    t->carrier_phase = 0.0; t->carrier_freq = initial_doppler; t->fll_bandwidth = 50.0;
    t->pll = loop3_new(15.0, 0.001); t->fll_assist = 1;
    return t;
}
void orc_track_free(orc_tracker* t) { if (t) { free(t->nav_bits); free(t); } }
void orc_track_set_dll_bandwidth(orc_tracker* t, double bw) { t->dll = loop2_new(bw, 0.001); }
void orc_track_set_pll_bandwidth(orc_tracker* t, double bw) { t->pll = loop3_new(bw, 0.001); }

static double rem_euclid(double a, double b)
{
    const double r = fmod(a, b);
    return r < 0.0 ? r + fabs(b) : r;
}
static size_t as_usize(double x) { return x > 0.0 ? (size_t)x : 0; }

static double estimate_cn0(const orc_tracker* t)        /* tracking.rs:316-337 */
{
    if (t->cn0_n < 2) return 0.0;
    const double n = (double)t->cn0_n;
    double sum = 0.0;
    for (size_t i = 0; i < t->cn0_n; ++i) sum += t->cn0_buf[i];
    const double mean = sum / n;
    double vs = 0.0;
    for (size_t i = 0; i < t->cn0_n; ++i) { const double d = t->cn0_buf[i] - mean; vs += d * d; }
    const double variance = vs / (n - 1.0);
    if (variance <= 0.0 || mean <= 0.0) return 0.0;
    const double snr = (mean * mean) / variance;
    const double x = snr - 1.0;
    const double cn0_linear = (1.0 / 0.001) * (x > 0.01 ? x : 0.01);
    return 10.0 * log10(cn0_linear);
}

static void fill_state(const orc_tracker* t, double cn0, orc_track_state* o)
{
    o->prn = t->prn; o->code_phase = t->code_phase; o->carrier_freq_hz = t->carrier_freq;
    o->carrier_phase_rad = t->carrier_phase * 2.0 * ORC_PI; o->prompt_i = t->p_i; o->prompt_q = t->p_q; o->cn0_dbhz = cn0;
    o->carrier_lock = (uint8_t)t->carrier_lock; o->code_lock = (uint8_t)t->code_lock; o->bit_sync = (uint8_t)t->bit_sync;
    o->ms_count = t->ms_count;
}

void orc_track_process(orc_tracker* t, const orc_c64* samples, size_t n, const int8_t* code, orc_track_state* out)
{
    const double samples_per_chip = t->sample_rate / t->code_freq;          /* :178 */
    const double cl = (double)t->code_length;
    t->e_i = t->e_q = t->p_i = t->p_q = t->l_i = t->l_q = 0.0;
    for (size_t i = 0; i < n; ++i) {                                          /* :186-217 */
        const double tt = (double)i / t->sample_rate;
        const double arg = -2.0 * ORC_PI * (t->carrier_freq * tt + t->carrier_phase);
        const double sn = sin(arg), cs = cos(arg);
        const double sre = samples[i].re * cs - samples[i].im * sn;
        const double sim = samples[i].re * sn + samples[i].im * cs;
        const double chip = t->code_phase + (double)i / samples_per_chip;
        const size_t ei = as_usize(rem_euclid(chip - t->el_spacing / 2.0, cl));
        const size_t pi = as_usize(rem_euclid(chip, cl));
        const size_t li = as_usize(rem_euclid(chip + t->el_spacing / 2.0, cl));
        const double ec = (double)code[ei % t->code_length], pc = (double)code[pi % t->code_length], lc = (double)code[li % t->code_length];
        t->e_i += sre * ec; t->e_q += sim * ec;
        t->p_i += sre * pc; t->p_q += sim * pc;
        t->l_i += sre * lc; t->l_q += sim * lc;
    }
    const double early_power = sqrt(t->e_i * t->e_i + t->e_q * t->e_q);
    const double late_power = sqrt(t->l_i * t->l_i + t->l_q * t->l_q);
    const double dll_disc = early_power + late_power > 0.0 ? (early_power - late_power) / (early_power + late_power) : 0.0;
    const double pll_disc = fabs(t->p_i) > 1e-10 ? atan2(t->p_q, t->p_i) / (2.0 * ORC_PI) : 0.0;
    const double fll_disc = pll_disc;
    const double code_correction = loop2_update(&t->dll, dll_disc);
    double carrier_correction;
    if (t->fll_assist && t->ms_count < 100) carrier_correction = loop3_update(&t->pll, pll_disc) * 0.5 + fll_disc * t->fll_bandwidth * 0.5;
    else carrier_correction = loop3_update(&t->pll, pll_disc);
    t->code_phase += code_correction * t->el_spacing;
    t->code_phase = rem_euclid(t->code_phase, cl);
    t->carrier_freq += carrier_correction;
    t->carrier_phase += t->carrier_freq / t->sample_rate;
    t->carrier_phase = rem_euclid(t->carrier_phase, 1.0);
    const double code_doppler = t->carrier_freq * t->chipping_rate / 1575420000.0;
    t->code_freq = t->chipping_rate + code_doppler;
    const double prompt_power = t->p_i * t->p_i + t->p_q * t->p_q;
    if (t->cn0_n == 20) { memmove(t->cn0_buf, t->cn0_buf + 1, 19 * sizeof(double)); t->cn0_n = 19; }   /* push, then drop the oldest */
    t->cn0_buf[t->cn0_n++] = prompt_power;
    const double cn0 = estimate_cn0(t);
    t->carrier_lock = cn0 > 25.0 && t->ms_count > 10;
    t->code_lock = cn0 > 20.0;
    t->nav_acc += t->p_i;
    t->nav_count += 1;
    const int sign = t->p_i >= 0.0 ? 1 : -1;
    if (t->prev_sign != 0 && sign != t->prev_sign) {
        if (!t->bit_sync && t->ms_count > 20) t->bit_sync = 1;
    }
    t->prev_sign = sign;
    if (t->nav_count >= 20) {
        if (t->nav_n == t->nav_cap) {
            t->nav_cap = t->nav_cap ? 2 * t->nav_cap : 64;
            t->nav_bits = (int8_t*)realloc(t->nav_bits, t->nav_cap);
        }
        t->nav_bits[t->nav_n++] = t->nav_acc >= 0.0 ? 1 : -1;
        t->nav_acc = 0.0;
        t->nav_count = 0;
    }
    t->ms_count += 1;
    if (t->ms_count > 200) t->fll_assist = 0;
    if (out) fill_state(t, cn0, out);
}

void orc_track_state_get(const orc_tracker* t, orc_track_state* out) { fill_state(t, estimate_cn0(t), out); }

size_t orc_track_nav_bits(const orc_tracker* t, int8_t* out, size_t cap)
{
    const size_t n = t->nav_n < cap ? t->nav_n : cap;
    if (out && n) memcpy(out, t->nav_bits, n);
    return t->nav_n;
}

/* KAT helpers: output of the n-th update with a constant discriminator (tests test_loop_filter_{2nd,3rd}_converges) */
double orc_loop_filter_2nd_run(double bw, double T, double disc, size_t n_updates)
{
    loop2 f = loop2_new(bw, T);
    double o = 0.0;
    for (size_t i = 0; i < n_updates; ++i) o = loop2_update(&f, disc);
    return o;
}
double orc_loop_filter_3rd_run(double bw, double T, double disc, size_t n_updates)
{
    loop3 f = loop3_new(bw, T);
    double o = 0.0;
    for (size_t i = 0; i < n_updates; ++i) o = loop3_update(&f, disc);
    return o;
}
/* dll_s_curve, tracking.rs:441-456 */
void orc_dll_s_curve(double el_spacing, size_t num_points, double* err, double* disc)
{
    for (size_t i = 0; i < num_points; ++i) {
        const double e = -1.5 + 3.0 * (double)i / (double)(num_points - 1);
        const double a = 1.0 - fabs(e - el_spacing / 2.0), b = 1.0 - fabs(e + el_spacing / 2.0);
        const double ec = a > 0.0 ? a : 0.0, lc = b > 0.0 ? b : 0.0;
        err[i] = e;
        disc[i] = ec + lc > 0.0 ? (ec - lc) / (ec + lc) : 0.0;
    }
}
