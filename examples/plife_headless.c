/*
 * plife_headless.c -- a plain C client of the C ABI (include/plife.h): no Python, no torch, no CUDA headers.
 *
 * What a non-Python host (the Java shim's native side, a C++ tool) does with the library: create a handle, set the
 * reference's physics settings, produce a state (here the library's headless generators, which reproduce the reference's
 * default setters on a seeded stream), step, and read the result back.
 *
 *   plife_headless [n] [types] [rmax] [steps] [precision: 32|64] [seed]
 *
 * Prints one line:  n=<n> steps=<k> nx=<nx> pair_evals=<p> checksum=<hex> ms_per_step=<t>
 * The checksum is an order-sensitive hash of the downloaded fp64 positions, velocities, types and ids, so the same
 * arguments give the same line as the Python host (tests/test_c_client.py).
 * Exit status: 0 ok; 2 = the library reported an error (message on stderr).  There is no CPU fallback: without a usable
 * CUDA device plife_create() fails and this program exits 2.
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "plife.h"

static int check(plife_handle *h, int rc, const char *what)
{
    if (rc == PLIFE_OK) return 0;
    fprintf(stderr, "plife_headless: %s failed: %s (%s)\n", what, plife_status_string(rc), h ? plife_last_error(h) : "no handle");
    return 1;
}

static uint64_t mix(uint64_t hash, const void *data, size_t bytes)
{
    const unsigned char *p = (const unsigned char *)data;
    for (size_t i = 0; i < bytes; i++) hash = (hash ^ p[i]) * 0x100000001B3ull; /* FNV-1a */
    return hash;
}

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int main(int argc, char **argv)
{
    const int64_t n = argc > 1 ? atoll(argv[1]) : 10000;
    const int32_t m = argc > 2 ? atoi(argv[2]) : 6;
    const double rmax = argc > 3 ? atof(argv[3]) : 0.04;
    const int32_t steps = argc > 4 ? atoi(argv[4]) : 10;
    const int32_t precision = (argc > 5 && atoi(argv[5]) == 64) ? PLIFE_F64 : PLIFE_F32;
    const uint64_t seed = argc > 6 ? strtoull(argv[6], NULL, 0) : 0x5EED0001ull;

    plife_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.precision = precision;
    cfg.capacity = n;
    plife_handle *h = NULL;
    if (check(NULL, plife_create(&cfg, &h), "plife_create")) return 2;

    plife_settings st;
    memset(&st, 0, sizeof st);
    st.rmax = rmax;      /* B/PhysicsSettings.java defaults otherwise */
    st.friction = 0.85;
    st.force = 1.0;
    st.wrap = 1;
    int bad = check(h, plife_set_settings(h, &st), "plife_set_settings") ||
              check(h, plife_random_matrix(h, m, seed), "plife_random_matrix") ||
              check(h, plife_init_uniform(h, n, seed), "plife_init_uniform");
    double ms = 0.0;
    if (!bad) {
        bad = check(h, plife_step(h, 0.02, 1), "plife_step (warm-up)") || check(h, plife_sync(h), "plife_sync");
        const double t0 = now_ms();
        if (!bad && steps > 1) bad = check(h, plife_step(h, 0.02, steps - 1), "plife_step") || check(h, plife_sync(h), "plife_sync");
        ms = steps > 1 ? (now_ms() - t0) / (steps - 1) : 0.0;
    }
    plife_step_stats stats;
    memset(&stats, 0, sizeof stats);
    uint64_t hash = 0xCBF29CE484222325ull;
    if (!bad) bad = check(h, plife_get_step_stats(h, &stats), "plife_get_step_stats");
    if (!bad) {
        double *pos = (double *)malloc(sizeof(double) * 2 * (size_t)n), *vel = (double *)malloc(sizeof(double) * 2 * (size_t)n);
        int32_t *type = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
        uint32_t *id = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
        if (!pos || !vel || !type || !id) {
            fprintf(stderr, "plife_headless: out of host memory\n");
            bad = 1;
        } else if (!(bad = check(h, plife_download(h, pos, vel, type, id), "plife_download"))) {
            hash = mix(hash, pos, sizeof(double) * 2 * (size_t)n);
            hash = mix(hash, vel, sizeof(double) * 2 * (size_t)n);
            hash = mix(hash, type, sizeof(int32_t) * (size_t)n);
            hash = mix(hash, id, sizeof(uint32_t) * (size_t)n);
        }
        free(pos); free(vel); free(type); free(id);
    }
    if (!bad)
        printf("n=%" PRId64 " steps=%d nx=%d pair_evals=%" PRId64 " checksum=%016" PRIx64 " ms_per_step=%.4f\n", plife_count(h), steps,
               stats.nx, stats.pair_evals, hash, ms);
    plife_destroy(h);
    return bad ? 2 : 0;
}
