// api.cu — the extern "C" boundary of libr4w_b200.so (include/r4w_b200.h).
// Every entry point converts internal failures into r4wb_error + a thread-local message; nothing unwinds.
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "acq.cuh"
#include "compose.cuh"
#include "synth.cuh"
#include "track.cuh"

namespace r4wb {

std::atomic<uint64_t> g_kernel_launches{0};
static std::atomic<int> g_n_devices{1};            // devices the host-buffer batch calls shard over (r4wb_init_devices)
static thread_local cudaStream_t t_stream = nullptr;
static thread_local std::string t_error;

cudaStream_t current_stream() { return t_stream; }
void e1_code_chips(uint32_t channel, uint32_t prn, int8_t* out);
void gps_ca_code_chips(uint32_t prn, int8_t* out);
void gps_l5_i5_code_chips(uint32_t prn, int8_t* out);
void glonass_code_chips(int8_t* out);

template <typename F>
static r4wb_error guard(F&& body)
{
    try {
        body();
        return R4WB_OK;
    } catch (const Failure& f) {
        t_error = f.what;
        return f.code;
    } catch (const std::bad_alloc&) {
        t_error = "host allocation failed";
        return R4WB_ERR_ALLOCATION_FAILED;
    } catch (const std::exception& e) {
        t_error = e.what();
        return R4WB_ERR_INVALID_PARAMETER;
    } catch (...) {
        t_error = "unknown failure";
        return R4WB_ERR_INVALID_PARAMETER;
    }
}

template <typename H, typename F>
static r4wb_error guard_on(const H* h, F&& body)
{
    return guard([&] { R4WB_CUDA(cudaSetDevice(h->device)); body(); });
}

static void require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        fail(R4WB_ERR_CUDA, "no CUDA device available (%s); libr4w_b200 has no CPU fallback",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
}

}  // namespace r4wb

using namespace r4wb;

// A handle lives on the device that was current when it was created; every call that touches the device re-selects it, so a
// handle can be used from a host thread whose current device is another one (new threads start on device 0).
static int device_now() { int d = 0; cudaGetDevice(&d); return d; }
// `peers[d]` (d != device): the same object on another device, created on first use by a sharded batch call
struct r4wb_scenario {
    int device = device_now();
    Scenario impl;
    std::vector<std::unique_ptr<Scenario>> peers;
    double multi_power = -1.0;                     // sum |s|^2 of the last sharded generate (-1: last call was single-device)
    explicit r4wb_scenario(const r4wb_scenario_cfg& c) : impl(c) {}
    ~r4wb_scenario()
    {
        for (size_t d = 0; d < peers.size(); ++d) if (peers[d]) { cudaSetDevice((int)d); peers[d].reset(); }
        cudaSetDevice(device);
    }
};
// `PcpsAcquisition::acquire(&self)` is re-entrant in the reference; here a handle owns device scratch, so concurrent calls on
// one handle take turns (`mu`)
struct r4wb_pcps {
    int device = device_now();
    Pcps impl;
    std::mutex mu;
    uint64_t code_length; double fs;
    double dmax = 5000.0, dstep = 500.0, threshold = 2.5; uint64_t coherent = 1;
    std::vector<std::unique_ptr<Pcps>> peers;
    r4wb_pcps(uint64_t n, double f) : impl(n, f), code_length(n), fs(f) {}
    ~r4wb_pcps()
    {
        for (size_t d = 0; d < peers.size(); ++d) if (peers[d]) { cudaSetDevice((int)d); peers[d].reset(); }
        cudaSetDevice(device);
    }
    Pcps& on(int d)                                // the engine of device d (current device must be d)
    {
        if (d == device) return impl;
        if (peers.size() <= (size_t)d) peers.resize(d + 1);
        if (!peers[d]) peers[d].reset(new Pcps(code_length, fs));
        Pcps& p = *peers[d];
        p.set_doppler_range(dmax, dstep); p.set_threshold(threshold); p.set_coherent_periods(coherent);
        return p;
    }
};

// units [0, n) split into `parts` contiguous shares; runs body(part, begin, end) on one host thread per part (part 0 on the
// calling thread), each with its own current device set by the body.  The first failure is rethrown on the caller.
template <typename F>
static void run_sharded(int parts, uint64_t n, F&& body)
{
    std::vector<std::thread> th;
    std::vector<Failure> err(parts, Failure{R4WB_OK, ""});
    auto one = [&](int p) {
        const uint64_t b = n * (uint64_t)p / (uint64_t)parts, e = n * (uint64_t)(p + 1) / (uint64_t)parts;
        try { if (e > b) body(p, b, e); }
        catch (const Failure& f) { err[p] = f; }
        catch (const std::exception& x) { err[p] = Failure{R4WB_ERR_INVALID_PARAMETER, x.what()}; }
        catch (...) { err[p] = Failure{R4WB_ERR_INVALID_PARAMETER, "unknown failure"}; }
    };
    for (int p = 1; p < parts; ++p) th.emplace_back(one, p);
    one(0);
    for (auto& t : th) t.join();
    for (const Failure& f : err) if (f.code != R4WB_OK) throw f;
}
struct r4wb_composer { int device = device_now(); Composer impl; r4wb_composer(uint32_t n, double fs, double sd, uint64_t seed) : impl(n, fs, sd, seed) {} };
struct r4wb_tracker { int device = device_now(); TrackerBank impl; r4wb_tracker(const r4wb_track_cfg* c, uint32_t n) : impl(c, n) {} };

extern "C" {

const char* r4wb_version(void) { return "r4w_b200 0.2.0 (sm_100a)"; }
const char* r4wb_last_error(void) { return t_error.c_str(); }

r4wb_error r4wb_init(int device)
{
    return guard([&] {
        require_device();
        if (device >= 0) R4WB_CUDA(cudaSetDevice(device));
        R4WB_CUDA(cudaFree(nullptr));
    });
}

r4wb_error r4wb_init_devices(int n_gpus)
{
    return guard([&] {
        require_device();
        int have = 0, cur = 0;
        R4WB_CUDA(cudaGetDeviceCount(&have));
        R4WB_CUDA(cudaGetDevice(&cur));
        const int n = n_gpus <= 0 ? have : n_gpus;
        if (n > have) fail(R4WB_ERR_INVALID_PARAMETER, "%d devices requested, %d visible", n, have);
        for (int d = 0; d < n; ++d) {
            R4WB_CUDA(cudaSetDevice(d));
            R4WB_CUDA(cudaFree(nullptr));
        }
        R4WB_CUDA(cudaSetDevice(cur));
        g_n_devices.store(n);
    });
}

int r4wb_devices_initialised(void) { return g_n_devices.load(); }

r4wb_error r4wb_device_count(int* n)
{
    if (!n) { t_error = "n is NULL"; return R4WB_ERR_NULL_POINTER; }
    *n = 0;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { *n = 0; t_error = cudaGetErrorString(e); cudaGetLastError(); return R4WB_ERR_CUDA; }
    return R4WB_OK;
}

r4wb_error r4wb_set_stream(void* cuda_stream)
{
    t_stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return R4WB_OK;
}

r4wb_error r4wb_host_alloc(void** p, size_t bytes)
{
    if (!p) { t_error = "p is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard([&] {
        require_device();
        cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault);
        if (e != cudaSuccess) fail(R4WB_ERR_ALLOCATION_FAILED, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    });
}

r4wb_error r4wb_host_free(void* p)
{
    if (!p) return R4WB_OK;
    return guard([&] { R4WB_CUDA(cudaFreeHost(p)); });
}

uint64_t r4wb_kernel_launches(void) { return g_kernel_launches.load(); }

/* ---------------------------------------------------------------- scenario */
r4wb_error r4wb_scenario_create(const r4wb_scenario_cfg* cfg, r4wb_scenario** out)
{
    if (!cfg || !out) { t_error = "cfg/out is NULL"; return R4WB_ERR_NULL_POINTER; }
    *out = nullptr;
    return guard([&] {
        require_device();
        *out = new r4wb_scenario(*cfg);
    });
}

void r4wb_scenario_destroy(r4wb_scenario* h) { delete h; }
uint64_t r4wb_scenario_total_samples(const r4wb_scenario* h) { return h ? h->impl.total_samples() : 0; }
uint64_t r4wb_scenario_block_size(const r4wb_scenario* h) { return h ? h->impl.block_size() : 0; }
int r4wb_scenario_is_done(const r4wb_scenario* h) { return h ? (h->impl.is_done() ? 1 : 0) : 1; }
double r4wb_scenario_progress(const r4wb_scenario* h) { return h ? h->impl.progress() : 1.0; }
uint64_t r4wb_scenario_current_sample(const r4wb_scenario* h) { return h ? h->impl.current_sample() : 0; }

r4wb_error r4wb_scenario_reset(r4wb_scenario* h)
{
    if (!h) { t_error = "handle is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] { h->impl.reset(); });
}

r4wb_error r4wb_scenario_generate_block(r4wb_scenario* h, uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt,
                                        uint64_t* written)
{
    if (!h || !written) { t_error = "handle/written is NULL"; return R4WB_ERR_NULL_POINTER; }
    *written = 0;
    return guard_on(h, [&] { *written = h->impl.generate_block(n, dst, where, fmt); });
}

r4wb_error r4wb_scenario_generate_block_view(r4wb_scenario* h, uint64_t n, r4wb_fmt fmt, const void** block, uint64_t* written)
{
    if (!h || !block || !written) { t_error = "handle/block/written is NULL"; return R4WB_ERR_NULL_POINTER; }
    *written = 0;
    *block = nullptr;
    return guard_on(h, [&] { *block = h->impl.generate_block_view(n, fmt, written); });
}

r4wb_error r4wb_scenario_generate(r4wb_scenario* h, uint64_t first, uint64_t n, void* dst, r4wb_mem where, r4wb_fmt fmt)
{
    if (!h) { t_error = "handle is NULL"; return R4WB_ERR_NULL_POINTER; }
    const int nd = g_n_devices.load();
    const uint64_t B = h->impl.block_size();
    if (nd > 1 && current_stream() == nullptr && where == R4WB_MEM_HOST && dst && B && n >= (uint64_t)nd * 64 * B && first <= h->impl.total_samples() &&
        n <= h->impl.total_samples() - first && (unsigned)fmt <= (unsigned)R4WB_FMT_CU8) {
        // time segments on 1 ms block boundaries, one per device, each straight into its slice of the caller's buffer
        const r4wb_error rc = guard([&] {
            const uint64_t b0 = (first + B - 1) / B, b1 = (first + n) / B;       // whole blocks inside the range
            std::vector<double> pw(nd, 0.0);
            const size_t bps = fmt_bytes(fmt);
            if (h->peers.size() < (size_t)nd) h->peers.resize(nd);
            run_sharded(nd, b1 - b0, [&](int p, uint64_t lo, uint64_t hi) {
                R4WB_CUDA(cudaSetDevice(p));
                const uint64_t s0 = p == 0 ? first : (b0 + lo) * B, s1 = p == nd - 1 ? first + n : (b0 + hi) * B;
                Scenario* sc = &h->impl;
                if (p != h->device) {
                    if (!h->peers[p]) h->peers[p].reset(new Scenario(h->impl.config()));
                    sc = h->peers[p].get();
                }
                sc->generate(s0, s1 - s0, static_cast<unsigned char*>(dst) + (size_t)(s0 - first) * bps, R4WB_MEM_HOST, fmt);
                pw[p] = sc->last_power_sum();
            });
            double t = 0.0;
            for (double v : pw) t += v;
            h->multi_power = t;
        });
        cudaSetDevice(h->device);
        return rc;
    }
    h->multi_power = -1.0;
    return guard_on(h, [&] { h->impl.generate(first, n, dst, where, fmt); });
}

r4wb_error r4wb_scenario_generate_rest(r4wb_scenario* h, void* dst, uint64_t cap, r4wb_mem where, r4wb_fmt fmt, uint64_t* written)
{
    if (!h || !written) { t_error = "handle/written is NULL"; return R4WB_ERR_NULL_POINTER; }
    *written = 0;
    return guard_on(h, [&] { *written = h->impl.generate_rest(dst, cap, where, fmt); });
}

r4wb_error r4wb_scenario_write_file(r4wb_scenario* h, const char* path, r4wb_fmt fmt, uint64_t* samples, uint64_t* bytes,
                                    double* power_sum)
{
    if (!h || !path) { t_error = "handle/path is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] {
        const double p = h->impl.write_file(path, fmt, samples, bytes);
        if (power_sum) *power_sum = p;
    });
}

r4wb_error r4wb_scenario_last_power_sum(const r4wb_scenario* h, double* power_sum)
{
    if (!h || !power_sum) { t_error = "handle/power_sum is NULL"; return R4WB_ERR_NULL_POINTER; }
    if (h->multi_power >= 0.0) { *power_sum = h->multi_power; return R4WB_OK; }
    return guard_on(h, [&] { *power_sum = const_cast<r4wb_scenario*>(h)->impl.last_power_sum(); });
}

r4wb_error r4wb_scenario_set_profiling(r4wb_scenario* h, int enabled)
{
    if (!h) { t_error = "handle is NULL"; return R4WB_ERR_NULL_POINTER; }
    h->impl.set_profiling(enabled != 0);
    return R4WB_OK;
}

r4wb_error r4wb_scenario_last_profile(r4wb_scenario* h, double* ms, uint64_t* launches)
{
    if (!h || !ms || !launches) { t_error = "handle/ms/launches is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] { h->impl.last_profile(ms, launches); });
}

uint32_t r4wb_scenario_last_path(const r4wb_scenario* h) { return h ? h->impl.last_path() : 0u; }

r4wb_error r4wb_scenario_status(const r4wb_scenario* h, r4wb_sat_status* out, uint32_t cap, uint32_t* n)
{
    if (!h || !out) { t_error = "handle/out is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] { h->impl.status(out, cap, n); });
}

/* test hook (not part of the drop-in surface): prologue entry of a canonical block */
r4wb_error r4wb_debug_block_params(r4wb_scenario* h, uint64_t block, uint32_t sat, double* out12)
{
    if (!h || !out12) { t_error = "handle/out is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] { h->impl.debug_block(block, sat, out12); });
}

/* ---------------------------------------------------------------- codes */
r4wb_error r4wb_e1_code(uint32_t channel, uint8_t prn, int8_t* out, uint64_t cap)
{
    if (!out) { t_error = "out is NULL"; return R4WB_ERR_NULL_POINTER; }
    if (channel > 1 || prn < 1 || prn > 50) { t_error = "Galileo PRN must be 1-50, channel 0/1"; return R4WB_ERR_INVALID_PARAMETER; }
    if (cap < 4092) { t_error = "code buffer needs 4092 entries"; return R4WB_ERR_INVALID_SIZE; }
    e1_code_chips(channel, prn, out);
    return R4WB_OK;
}

r4wb_error r4wb_gps_ca_code(uint8_t prn, int8_t* out, uint64_t cap)
{
    if (!out) { t_error = "out is NULL"; return R4WB_ERR_NULL_POINTER; }
    if (cap < 1023) { t_error = "code buffer needs 1023 entries"; return R4WB_ERR_INVALID_SIZE; }
    return guard([&] { gps_ca_code_chips(prn, out); });
}

r4wb_error r4wb_gps_l5_code(uint8_t prn, int8_t* out, uint64_t cap)
{
    if (!out) { t_error = "out is NULL"; return R4WB_ERR_NULL_POINTER; }
    if (cap < 10230) { t_error = "code buffer needs 10230 entries"; return R4WB_ERR_INVALID_SIZE; }
    return guard([&] { gps_l5_i5_code_chips(prn, out); });
}

r4wb_error r4wb_glonass_code(int8_t* out, uint64_t cap)
{
    if (!out) { t_error = "out is NULL"; return R4WB_ERR_NULL_POINTER; }
    if (cap < 511) { t_error = "code buffer needs 511 entries"; return R4WB_ERR_INVALID_SIZE; }
    glonass_code_chips(out);
    return R4WB_OK;
}

r4wb_error r4wb_e1c_secondary(int8_t* out, uint64_t cap)
{
    static const int8_t sec[25] = {1, 1, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1, -1, 1, -1, 1, -1, -1, -1, -1, 1, 1, 1, -1};
    if (!out) { t_error = "out is NULL"; return R4WB_ERR_NULL_POINTER; }
    if (cap < 25) { t_error = "secondary code needs 25 entries"; return R4WB_ERR_INVALID_SIZE; }
    std::memcpy(out, sec, 25);
    return R4WB_OK;
}

r4wb_error r4wb_e1c_replica(uint8_t prn, double sample_rate, int8_t* out, uint64_t n)
{
    if (!out) { t_error = "out is NULL"; return R4WB_ERR_NULL_POINTER; }
    if (prn < 1 || prn > 50 || !(sample_rate > 0.0)) { t_error = "bad PRN / sample rate"; return R4WB_ERR_INVALID_PARAMETER; }
    int8_t code[4092];
    e1_code_chips(1, prn, code);
    // same index expression as the emitter with zero delay (gnss/satellite_emitter.rs:245, 268-270, 303-305)
    const double spc = sample_rate / 1023000.0;
    for (uint64_t i = 0; i < n; ++i) {
        const double cf = (double)i / spc;
        const double cm = std::fmod(cf, 4092.0);
        uint32_t c = cm > 0.0 ? (uint32_t)cm : 0u;
        if (c > 4091u) c = 4091u;
        const double cp = cf - std::floor(cf);
        out[i] = (int8_t)(std::fmod(cp * 2.0, 2.0) < 1.0 ? code[c] : -code[c]);
    }
    return R4WB_OK;
}

/* ---------------------------------------------------------------- PCPS */
r4wb_error r4wb_pcps_create(uint64_t code_length, double sample_rate, r4wb_pcps** out)
{
    if (!out) { t_error = "out is NULL"; return R4WB_ERR_NULL_POINTER; }
    *out = nullptr;
    return guard([&] {
        require_device();
        *out = new r4wb_pcps(code_length, sample_rate);
    });
}

void r4wb_pcps_destroy(r4wb_pcps* h) { delete h; }

r4wb_error r4wb_pcps_set_doppler_range(r4wb_pcps* h, double max_hz, double step_hz)
{
    if (!h) { t_error = "handle is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] { h->impl.set_doppler_range(max_hz, step_hz); h->dmax = max_hz; h->dstep = step_hz; });
}

r4wb_error r4wb_pcps_set_threshold(r4wb_pcps* h, double threshold)
{
    if (!h) { t_error = "handle is NULL"; return R4WB_ERR_NULL_POINTER; }
    h->impl.set_threshold(threshold);
    h->threshold = threshold;
    return R4WB_OK;
}

r4wb_error r4wb_pcps_set_coherent_periods(r4wb_pcps* h, uint64_t periods)
{
    if (!h) { t_error = "handle is NULL"; return R4WB_ERR_NULL_POINTER; }
    h->impl.set_coherent_periods(periods);
    h->coherent = periods;
    return R4WB_OK;
}

uint64_t r4wb_pcps_fft_size(const r4wb_pcps* h) { return h ? h->impl.fft_size() : 0; }
uint32_t r4wb_pcps_num_doppler_bins(const r4wb_pcps* h) { return h ? h->impl.num_bins() : 0; }
uint64_t r4wb_pcps_guard_count(const r4wb_pcps* h)
{
    if (!h) return 0;
    uint64_t n = h->impl.guard_count();
    for (const auto& p : h->peers) if (p) n += p->guard_count();
    return n;
}

r4wb_error r4wb_pcps_set_profiling(r4wb_pcps* h, int enabled)
{
    if (!h) { t_error = "handle is NULL"; return R4WB_ERR_NULL_POINTER; }
    h->impl.set_profiling(enabled != 0);
    return R4WB_OK;
}

r4wb_error r4wb_pcps_last_profile(const r4wb_pcps* h, double* ms, uint64_t* launches)
{
    if (!h || !ms || !launches) { t_error = "NULL argument"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] { h->impl.last_profile(ms, launches); });
}

r4wb_error r4wb_pcps_acquire(r4wb_pcps* h, const void* input, r4wb_fmt fmt, uint64_t n_input, const int8_t* code,
                             uint64_t code_len, uint8_t prn, r4wb_acq_result* out)
{
    if (!h || !input || !code || !out) { t_error = "NULL argument"; return R4WB_ERR_NULL_POINTER; }
    std::lock_guard<std::mutex> lk(h->mu);
    return guard_on(h, [&] { h->impl.acquire_batch(input, fmt, R4WB_MEM_HOST, 1, 0, n_input, code, code_len, &prn, 1, out); });
}

r4wb_error r4wb_pcps_acquire_batch(r4wb_pcps* h, const void* input, r4wb_fmt fmt, r4wb_mem where, uint64_t n_snapshots,
                                   uint64_t snapshot_stride, uint64_t n_input, const int8_t* codes, uint64_t code_len,
                                   const uint8_t* prns, uint32_t n_codes, r4wb_acq_result* out)
{
    if (!h || !input || !codes || !out) { t_error = "NULL argument"; return R4WB_ERR_NULL_POINTER; }
    std::lock_guard<std::mutex> lk(h->mu);
    const int nd = g_n_devices.load();
    if (nd > 1 && current_stream() == nullptr && n_snapshots >= (uint64_t)(2 * nd) && (fmt == R4WB_FMT_CF32 || fmt == R4WB_FMT_CF64)) {
        // snapshots are independent: device p takes a contiguous share and writes its rows of `out` itself
        const r4wb_error rc = guard([&] {
            const size_t bps = fmt_bytes(fmt);
            if (h->peers.size() < (size_t)nd) h->peers.resize(nd);
            run_sharded(nd, n_snapshots, [&](int p, uint64_t lo, uint64_t hi) {
                R4WB_CUDA(cudaSetDevice(p));
                Pcps& eng = h->on(p);
                const unsigned char* src = static_cast<const unsigned char*>(input) + (size_t)lo * snapshot_stride * bps;
                r4wb_acq_result* dst = out + (size_t)lo * n_codes;
                if (where == R4WB_MEM_HOST || p == h->device) {
                    eng.acquire_batch(src, fmt, where, hi - lo, snapshot_stride, n_input, codes, code_len, prns, n_codes, dst);
                } else {
                    // device-resident input lives on the handle's device: the share is copied over NVLink first
                    const size_t bytes = ((size_t)(hi - lo - 1) * snapshot_stride + n_input) * bps;
                    DevBuf<unsigned char> tmp;
                    tmp.reserve(bytes);
                    R4WB_CUDA(cudaMemcpyPeer(tmp.p, p, src, h->device, bytes));
                    eng.acquire_batch(tmp.p, fmt, R4WB_MEM_DEVICE, hi - lo, snapshot_stride, n_input, codes, code_len, prns, n_codes, dst);
                }
            });
        });
        cudaSetDevice(h->device);
        return rc;
    }
    return guard_on(h, [&] { h->impl.acquire_batch(input, fmt, where, n_snapshots, snapshot_stride, n_input, codes, code_len, prns, n_codes, out); });
}

r4wb_error r4wb_pcps_acquire_grid(r4wb_pcps* h, const void* input, r4wb_fmt fmt, uint64_t n_input, const int8_t* code,
                                  uint64_t code_len, double* power_out, uint64_t cap)
{
    if (!h || !input || !code || !power_out) { t_error = "NULL argument"; return R4WB_ERR_NULL_POINTER; }
    std::lock_guard<std::mutex> lk(h->mu);
    return guard_on(h, [&] { h->impl.acquire_grid(input, fmt, n_input, code, code_len, power_out, cap); });
}


// ---- tracking channels
r4wb_error r4wb_track_create(const r4wb_track_cfg* cfgs, uint32_t n_channels, r4wb_tracker** out)
{
    if (!cfgs || !out) { t_error = "cfgs/out is NULL"; return R4WB_ERR_NULL_POINTER; }
    *out = nullptr;
    return guard([&] {
        require_device();
        *out = new r4wb_tracker(cfgs, n_channels);
    });
}

void r4wb_track_destroy(r4wb_tracker* h) { delete h; }

r4wb_error r4wb_track_process(r4wb_tracker* h, const void* samples, r4wb_fmt fmt, r4wb_mem where, uint64_t n_per_period, uint64_t n_periods,
                              uint64_t channel_stride, const int8_t* codes, uint64_t code_stride, r4wb_track_state* out)
{
    if (!h) { t_error = "handle is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] { h->impl.process(samples, fmt, where, n_per_period, n_periods, channel_stride, codes, code_stride, out); });
}

r4wb_error r4wb_track_state_get(const r4wb_tracker* h, r4wb_track_state* out, uint32_t cap)
{
    if (!h || !out) { t_error = "handle/out is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] { h->impl.state(out, cap); });
}

r4wb_error r4wb_track_nav_bits(const r4wb_tracker* h, uint32_t channel, int8_t* out, uint64_t cap, uint64_t* n)
{
    if (!h || !n) { t_error = "handle/n is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] { *n = h->impl.nav_bits(channel, out, cap); });
}


// ---- r4w-sim composer
r4wb_error r4wb_composer_create(uint32_t n_emitters, double sample_rate, double noise_std, uint64_t seed, r4wb_composer** out)
{
    if (!out) { t_error = "out is NULL"; return R4WB_ERR_NULL_POINTER; }
    *out = nullptr;
    return guard([&] {
        require_device();
        *out = new r4wb_composer(n_emitters, sample_rate, noise_std, seed);
    });
}

void r4wb_composer_destroy(r4wb_composer* h) { delete h; }

r4wb_error r4wb_composer_reset(r4wb_composer* h)
{
    if (!h) { t_error = "handle is NULL"; return R4WB_ERR_NULL_POINTER; }
    h->impl.reset();
    return R4WB_OK;
}

r4wb_error r4wb_composer_block(r4wb_composer* h, const void* baseband, r4wb_fmt in_fmt, r4wb_mem in_where, uint64_t n, const double* doppler_hz,
                               const double* amplitude, const uint8_t* active, void* out, r4wb_fmt out_fmt, r4wb_mem out_where)
{
    if (!h) { t_error = "handle is NULL"; return R4WB_ERR_NULL_POINTER; }
    return guard_on(h, [&] { h->impl.block(baseband, in_fmt, in_where, n, doppler_hz, amplitude, active, out, out_fmt, out_where); });
}

r4wb_error r4wb_composer_phases(const r4wb_composer* h, double* out, uint32_t cap)
{
    if (!h || !out) { t_error = "handle/out is NULL"; return R4WB_ERR_NULL_POINTER; }
    const std::vector<double>& p = h->impl.phases();
    for (uint32_t k = 0; k < cap && k < p.size(); ++k) out[k] = p[k];
    return R4WB_OK;
}

}  // extern "C"
