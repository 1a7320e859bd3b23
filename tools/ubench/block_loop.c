/* block_loop.c — the reference's sink loop `while !scenario.is_done() { let block = scenario.generate_block(block_size); ... }`
 * (/root/reference/crates/r4w-cli/src/main.rs:4488-4500) as a COMPILED caller of the C-ABI, so that bench.py's e2e_block_api
 * times the library and not the Python interpreter's ~1 us per ctypes call.  bench.py hands in the handle and the entry points.
 *   gcc -O2 -shared -fPIC -o libblock_loop.so block_loop.c
 * mode 0: r4wb_scenario_generate_block into `buf` (the owned-Vec form: one host copy per block)
 * mode 1: r4wb_scenario_generate_block_view, pointer only (the cost of the call itself)
 * mode 2: r4wb_scenario_generate_block_view and the consumer reads every byte of the block (a 64-bit fold; stands for a
 *         sink's write()/checksum) */
#include <stdint.h>
#include <time.h>

typedef int (*gen_block_fn)(void*, uint64_t, void*, int, int, uint64_t*);
typedef int (*gen_view_fn)(void*, uint64_t, int, const void**, uint64_t*);
typedef int (*is_done_fn)(const void*);

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int r4wb_block_loop(void* gen, void* view, void* done, void* h, uint64_t nb, void* buf, int mode, uint64_t bytes_per_sample,
                    uint64_t* samples, uint64_t* calls, double* seconds, uint64_t* fold)
{
    gen_block_fn g = (gen_block_fn)gen;
    gen_view_fn v = (gen_view_fn)view;
    is_done_fn d = (is_done_fn)done;
    uint64_t got = 0, n_calls = 0, acc = 0;
    int rc = 0;
    const double t0 = now_s();
    while (!d(h)) {
        uint64_t w = 0;
        if (mode == 0) {
            rc = g(h, nb, buf, /*R4WB_MEM_HOST*/ 0, /*R4WB_FMT_CF32*/ 0, &w);
        } else {
            const void* p = 0;
            rc = v(h, nb, /*R4WB_FMT_CF32*/ 0, &p, &w);
            if (rc == 0 && mode == 2) {
                const uint64_t* q = (const uint64_t*)p;
                const uint64_t nq = w * bytes_per_sample / 8;
                uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                uint64_t i = 0;
                for (; i + 4 <= nq; i += 4) { a0 ^= q[i]; a1 += q[i + 1]; a2 ^= q[i + 2]; a3 += q[i + 3]; }
                for (; i < nq; ++i) a0 ^= q[i];
                acc += a0 ^ a1 ^ a2 ^ a3;
            }
        }
        if (rc != 0) break;
        got += w;
        ++n_calls;
    }
    *seconds = now_s() - t0;
    *samples = got;
    *calls = n_calls;
    *fold = acc;
    return rc;
}
