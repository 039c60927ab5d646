/*
 * plife_oracle.c -- CPU fp64 restatement of the Particle Life physics step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (libplife.so, the
 * `plife` host package) may link, import or call this file.  It is used by
 * tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs, always as the checker or as the reported CPU
 * baseline, never as the thing shipped.
 *
 * PARITY UNPINNED: the reference (tom-mohr/particle-life-app) ships no tests,
 * golden vectors or fixtures for this path and is Java 21 + JOML 1.10.1, which
 * cannot be executed in the build environment (no JVM).  This file is pinned
 * only by (1) line-by-line review against the citations below, (2) the
 * hand-derived known-answer tests in tests/test_oracle_kat.py, and (3) an
 * independent O(N^2) numpy restatement (oracle/bruteforce.py).
 *
 * Citations are relative to /root/reference/src/main/java/com/particle_life/
 *   B/ = backend/, A/ = app/.
 *
 * Third-party arithmetic restated here (not vendored in the reference):
 *   org.joml:joml:1.10.1 Vector3d (build.gradle.kts:201)
 *     div(s)        -> multiply by (1.0 / s)
 *     length()      -> sqrt(x*x + (y*y + z*z)), non-fused (joml.useMathFma off)
 *     lengthSquared -> x*x + (y*y + z*z), non-fused
 *     mulAdd(a,b,d) -> d = this*a + b, non-fused
 *   Compile with -ffp-contract=off so gcc does not fuse either.
 *
 * Data layout: AoS of {pos[3], vel[3], type, id}, sorted BY VALUE each step.
 * The reference sorts object references (B/Physics.java:343-353), so memory
 * locality here is better than the JVM's: as a timing baseline this is an
 * optimistic proxy for the Java path.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_ERR_INVALID (-1)
#define ORACLE_ERR_OOM (-2)

/* Accelerator kinds.  Kind 0 is the only accelerator in the reference snapshot
 * (A/Main.java:275-280).  Kinds 1..5 are builder-defined extensions with no
 * reference definition; they are specified in DESIGN.md. */
enum {
    ACC_PARTICLE_LIFE = 0,
    ACC_PARTICLE_LIFE_R = 1,
    ACC_PARTICLE_LIFE_R2 = 2,
    ACC_ROTATOR_90 = 3,
    ACC_ROTATOR_ATTR = 4,
    ACC_PLANETS = 5,
    ACC_KIND_COUNT = 6
};

typedef struct {
    double rmax;      /* B/PhysicsSettings.java:13 */
    double friction;  /* B/PhysicsSettings.java:27 */
    double force;     /* B/PhysicsSettings.java:32 */
    double dt;        /* B/PhysicsSettings.java:37 */
    int32_t wrap;     /* B/PhysicsSettings.java:8  */
    int32_t accel_kind;
    double accel_params[4]; /* [0] = beta for kinds 0..2 */
    int32_t m;              /* matrix size, B/DefaultMatrix.java:5-15 */
    int32_t pad_;
    const double *matrix;   /* row-major, [own type][other type], B/DefaultMatrix.java:39-41 */
} oracle_settings;

/* B/Particle.java:5-9 (Vector3d position, velocity; int type) + a persistent id
 * that the reference does not have (harness bookkeeping only). */
typedef struct {
    double p[3];
    double v[3];
    int32_t type;
    uint32_t id;
} particle;

typedef struct oracle_handle oracle_handle;

typedef struct {
    oracle_handle *h;
    int index;
    pthread_t thread;
} worker;

struct oracle_handle {
    int64_t n;
    particle *particles;       /* B/Physics.java:15 */
    particle *particles_buffer;/* B/Physics.java:20 */
    int32_t *containers;       /* B/Physics.java:18 */
    int64_t containers_len;
    int nx, ny;                /* B/Physics.java:23-24 */
    double container_size;     /* B/Physics.java:25 */
    oracle_settings s;
    double *matrix_copy;
    /* diagnostics */
    int diag;
    int32_t *nbr_count;
    uint64_t *nbr_hash;
    uint8_t *borderline;
    int64_t diag_cap;
    int64_t pair_evals, pair_hits;
    /* thread pool: stands in for B/LoadDistributor.java's cached pool */
    int nworkers;
    worker *workers;
    pthread_mutex_t mu;
    pthread_cond_t cv_go, cv_done;
    uint64_t generation;
    int pending;
    int shutdown;
    int pass;           /* 0 = velocity, 1 = position */
    int64_t chunk_len;  /* ceil(N/T), B/LoadDistributor.java:42 */
    int nchunks;
    int64_t *chunk_evals, *chunk_hits;
};

/* ------------------------------------------------------------------ */
/* B/Range.java                                                        */
/* ------------------------------------------------------------------ */

/* B/Range.java:46-57 */
static double range_wrap(double value)
{
    if (value < 0) {
        if (!(value > -1e15)) return value - floor(value); /* reference would spin ~forever */
        do {
            value += 1;
        } while (value < 0);
        return value;
    }
    if (!(value < 1e15)) return value - floor(value);
    while (value >= 1) {
        value -= 1;
    }
    return value;
}

/* B/Range.java:74-81 */
static double range_wrap_connection(double value)
{
    if (value < -0.5) {
        return value + 1;
    } else if (value >= 0.5) {
        return value - 1;
    }
    return value;
}

/* B/Range.java:89-96 */
static double range_clamp(double val)
{
    if (val < 0) {
        return 0;
    } else if (val > 1) {
        return 1;
    }
    return val;
}

/* ------------------------------------------------------------------ */
/* Accelerators: pos is the neighbour offset / rmax (B/Accelerator.java:5-17);
 * the result may alias pos.                                            */
/* ------------------------------------------------------------------ */

static double vec_length(const double *p) /* JOML Vector3d.length(), non-fused */
{
    return sqrt(p[0] * p[0] + (p[1] * p[1] + p[2] * p[2]));
}

static double particle_life_force(double a, double dist, double beta) /* A/Main.java:276-278 */
{
    return dist < beta ? (dist / beta - 1) : a * (1 - fabs(1 + beta - 2 * dist) / (1 - beta));
}

static void accelerate(const oracle_settings *s, double a, double *pos)
{
    switch (s->accel_kind) {
    default:
    case ACC_PARTICLE_LIFE: { /* A/Main.java:275-280 */
        double beta = s->accel_params[0];
        double dist = vec_length(pos);
        double force = particle_life_force(a, dist, beta);
        double k = force / dist;
        pos[0] *= k; pos[1] *= k; pos[2] *= k;
        break;
    }
    case ACC_PARTICLE_LIFE_R: { /* builder-defined */
        double beta = s->accel_params[0];
        double dist = vec_length(pos);
        double force = particle_life_force(a, dist, beta);
        double k = force / (dist * dist);
        pos[0] *= k; pos[1] *= k; pos[2] *= k;
        break;
    }
    case ACC_PARTICLE_LIFE_R2: { /* builder-defined */
        double beta = s->accel_params[0];
        double dist = vec_length(pos);
        double force = particle_life_force(a, dist, beta);
        double k = force / (dist * dist * dist);
        pos[0] *= k; pos[1] *= k; pos[2] *= k;
        break;
    }
    case ACC_ROTATOR_90: { /* builder-defined */
        double dist = vec_length(pos);
        double force = a * (1 - dist);
        double k = force / dist;
        double dx = -pos[1], dy = pos[0];
        pos[0] = dx * k; pos[1] = dy * k; pos[2] = 0;
        break;
    }
    case ACC_ROTATOR_ATTR: { /* builder-defined */
        double dist = vec_length(pos);
        double force = 1 - dist;
        double angle = -a * 3.14159265358979323846;
        double c = cos(angle), sn = sin(angle);
        double k = force / dist;
        double dx = c * pos[0] + sn * pos[1];
        double dy = -sn * pos[0] + c * pos[1];
        pos[0] = dx * k; pos[1] = dy * k; pos[2] = 0;
        break;
    }
    case ACC_PLANETS: { /* builder-defined */
        double r = vec_length(pos);
        r = r > 0.01 ? r : 0.01;
        double k = 0.01 / (r * r * r);
        pos[0] *= k; pos[1] *= k; pos[2] *= k;
        break;
    }
    }
}

/* ------------------------------------------------------------------ */
/* B/Physics.java                                                      */
/* ------------------------------------------------------------------ */

/* B/Physics.java:82-85 */
static void calc_nx_ny(oracle_handle *h)
{
    h->nx = (int)floor(1 / h->container_size);
    h->ny = (int)floor(1 / h->container_size);
}

/* B/Physics.java:362-375 */
static int get_container_index(const oracle_handle *h, const double *position)
{
    int cx = (int)(position[0] / h->container_size);
    int cy = (int)(position[1] / h->container_size);
    if (cx == h->nx) {
        cx = h->nx - 1;
    }
    if (cy == h->ny) {
        cy = h->ny - 1;
    }
    return cx + cy * h->nx;
}

/* B/Physics.java:377-385 */
static int wrap_container_x(const oracle_handle *h, int cx)
{
    if (cx < 0) {
        return cx + h->nx;
    } else if (cx >= h->nx) {
        return cx - h->nx;
    } else {
        return cx;
    }
}

/* B/Physics.java:387-395 */
static int wrap_container_y(const oracle_handle *h, int cy)
{
    if (cy < 0) {
        return cy + h->ny;
    } else if (cy >= h->ny) {
        return cy - h->ny;
    } else {
        return cy;
    }
}

/* B/Physics.java:309-354 -- serial stable counting sort */
static int make_containers(oracle_handle *h)
{
    h->container_size = h->s.rmax; /* :312 */
    calc_nx_ny(h);                 /* :313 */
    int64_t ncell = (int64_t)h->nx * h->ny;
    if (h->nx <= 0 || ncell > (int64_t)1 << 31) return ORACLE_ERR_INVALID;

    if (h->containers == NULL || h->containers_len != ncell) { /* :320-322 */
        free(h->containers);
        h->containers = (int32_t *)malloc(sizeof(int32_t) * (size_t)ncell);
        if (!h->containers) return ORACLE_ERR_OOM;
        h->containers_len = ncell;
    }
    memset(h->containers, 0, sizeof(int32_t) * (size_t)ncell); /* :323 */

    for (int64_t i = 0; i < h->n; i++) { /* :329-332 */
        int ci = get_container_index(h, h->particles[i].p);
        h->containers[ci]++;
    }

    int32_t offset = 0; /* :335-340 */
    for (int64_t i = 0; i < ncell; i++) {
        int32_t cap = h->containers[i];
        h->containers[i] = offset;
        offset += cap;
    }

    for (int64_t k = 0; k < h->n; k++) { /* :343-348 */
        int ci = get_container_index(h, h->particles[k].p);
        int32_t i = h->containers[ci];
        h->particles_buffer[i] = h->particles[k];
        h->containers[ci]++; /* afterwards containers[ci] is the END offset of cell ci */
    }

    particle *t = h->particles; /* :351-353 */
    h->particles = h->particles_buffer;
    h->particles_buffer = t;
    return ORACLE_OK;
}

static uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* B/Physics.java:397-441 */
static void update_velocity(oracle_handle *h, int64_t i, int64_t *evals, int64_t *hits)
{
    static const int neighborhood[9][2] = { /* B/Physics.java:88-98 */
        {-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {0, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
    const oracle_settings *s = &h->s;
    particle *p = &h->particles[i];

    double friction_factor = pow(s->friction, 60 * s->dt); /* :401 */
    p->v[0] *= friction_factor; /* :402 */
    p->v[1] *= friction_factor;
    p->v[2] *= friction_factor;

    int cx0 = (int)floor(p->p[0] / h->container_size); /* :404 (no ==nx clamp) */
    int cy0 = (int)floor(p->p[1] / h->container_size); /* :405 */

    int32_t cnt = 0;
    uint64_t hash = 0;
    uint8_t border = 0;
    const double r2 = s->rmax * s->rmax;

    for (int k = 0; k < 9; k++) { /* :407 */
        int cx = wrap_container_x(h, cx0 + neighborhood[k][0]); /* :408 */
        int cy = wrap_container_y(h, cy0 + neighborhood[k][1]); /* :409 */
        if (s->wrap) {
            cx = wrap_container_x(h, cx); /* :411 */
            cy = wrap_container_y(h, cy); /* :412 */
        } else {
            if (cx < 0 || cx >= h->nx || cy < 0 || cy >= h->ny) { /* :414 */
                continue;
            }
        }
        int ci = cx + cy * h->nx; /* :418 */

        int32_t start = ci == 0 ? 0 : h->containers[ci - 1]; /* :420 */
        int32_t stop = h->containers[ci];                    /* :421 */

        for (int32_t j = start; j < stop; j++) { /* :423 */
            if (i == j) continue;                /* :424 */
            const particle *q = &h->particles[j];
            (*evals)++;

            /* connection(): B/Physics.java:460-470 */
            double rel[3];
            rel[0] = q->p[0] - p->p[0];
            rel[1] = q->p[1] - p->p[1];
            rel[2] = q->p[2] - p->p[2];
            if (s->wrap) { /* B/Range.java:68-72 */
                rel[0] = range_wrap_connection(rel[0]);
                rel[1] = range_wrap_connection(rel[1]);
                rel[2] = 0;
            }

            double d2 = rel[0] * rel[0] + (rel[1] * rel[1] + rel[2] * rel[2]); /* :430 */
            if (h->diag && d2 != 0 && fabs(d2 - r2) <= 2e-6 * r2) border = 1;
            if (d2 != 0 && d2 <= s->rmax * s->rmax) { /* :432 */
                (*hits)++;
                if (h->diag) {
                    cnt++;
                    hash += mix64(q->id);
                }
                double inv = 1.0 / s->rmax; /* :434, JOML div = mul by reciprocal */
                rel[0] *= inv; rel[1] *= inv; rel[2] *= inv;
                accelerate(s, s->matrix[(size_t)p->type * s->m + q->type], rel); /* :435 */
                double k2 = s->rmax * s->force * s->dt; /* :437 */
                p->v[0] += rel[0] * k2;
                p->v[1] += rel[1] * k2;
                p->v[2] += rel[2] * k2;
            }
        }
    }
    if (h->diag) {
        h->nbr_count[i] = cnt;
        h->nbr_hash[i] = hash;
        h->borderline[i] = border;
    }
}

/* B/Physics.java:443-450, :499-505 */
static void update_position(oracle_handle *h, int64_t i)
{
    const oracle_settings *s = &h->s;
    particle *p = &h->particles[i];
    p->p[0] = p->v[0] * s->dt + p->p[0]; /* :447 mulAdd, non-fused */
    p->p[1] = p->v[1] * s->dt + p->p[1];
    p->p[2] = p->v[2] * s->dt + p->p[2];
    if (s->wrap) { /* B/Range.java:40-44 */
        p->p[0] = range_wrap(p->p[0]);
        p->p[1] = range_wrap(p->p[1]);
        p->p[2] = 0;
    } else { /* B/Range.java:83-87 */
        p->p[0] = range_clamp(p->p[0]);
        p->p[1] = range_clamp(p->p[1]);
        p->p[2] = 0;
    }
}

/* B/LoadDistributor.java:19-29 BatchProcessor.run */
static void run_chunk(oracle_handle *h, int c)
{
    int64_t start = (int64_t)c * h->chunk_len;
    int64_t stop = start + h->chunk_len;
    if (stop > h->n) stop = h->n;
    if (h->pass == 0) {
        int64_t ev = 0, hi = 0;
        for (int64_t i = start; i < stop; i++) update_velocity(h, i, &ev, &hi);
        h->chunk_evals[c] = ev;
        h->chunk_hits[c] = hi;
    } else {
        for (int64_t i = start; i < stop; i++) update_position(h, i);
    }
}

static void *worker_main(void *arg)
{
    worker *w = (worker *)arg;
    oracle_handle *h = w->h;
    uint64_t seen = 0;
    for (;;) {
        pthread_mutex_lock(&h->mu);
        while (!h->shutdown && h->generation == seen) pthread_cond_wait(&h->cv_go, &h->mu);
        if (h->shutdown) {
            pthread_mutex_unlock(&h->mu);
            return NULL;
        }
        seen = h->generation;
        pthread_mutex_unlock(&h->mu);
        if (w->index < h->nchunks) run_chunk(h, w->index);
        pthread_mutex_lock(&h->mu);
        if (--h->pending == 0) pthread_cond_signal(&h->cv_done);
        pthread_mutex_unlock(&h->mu);
    }
}

static void pool_stop(oracle_handle *h)
{
    if (h->nworkers == 0) return;
    pthread_mutex_lock(&h->mu);
    h->shutdown = 1;
    pthread_cond_broadcast(&h->cv_go);
    pthread_mutex_unlock(&h->mu);
    for (int i = 0; i < h->nworkers; i++) pthread_join(h->workers[i].thread, NULL);
    free(h->workers);
    free(h->chunk_evals);
    free(h->chunk_hits);
    h->workers = NULL;
    h->chunk_evals = h->chunk_hits = NULL;
    h->nworkers = 0;
    h->shutdown = 0;
}

static int pool_ensure(oracle_handle *h, int t)
{
    if (h->nworkers == t) return ORACLE_OK;
    pool_stop(h);
    free(h->chunk_evals);
    free(h->chunk_hits);
    h->workers = (worker *)calloc((size_t)t, sizeof(worker));
    h->chunk_evals = (int64_t *)calloc((size_t)t + 1, sizeof(int64_t));
    h->chunk_hits = (int64_t *)calloc((size_t)t + 1, sizeof(int64_t));
    if (!h->workers || !h->chunk_evals || !h->chunk_hits) return ORACLE_ERR_OOM;
    h->nworkers = t;
    for (int i = 0; i < t; i++) {
        h->workers[i].h = h;
        h->workers[i].index = i;
        pthread_create(&h->workers[i].thread, NULL, worker_main, &h->workers[i]);
    }
    return ORACLE_OK;
}

/* B/LoadDistributor.java:37-66: ceil(N/T)-sized contiguous chunks, join all */
static int distribute_load_evenly(oracle_handle *h, int threads, int pass)
{
    if (h->n <= 0) return ORACLE_OK; /* :39 */
    int64_t length = (h->n + threads - 1) / threads; /* :42 */
    int nchunks = (int)((h->n + length - 1) / length);
    h->pass = pass;
    h->chunk_len = length;
    h->nchunks = nchunks;
    if (threads == 1) {
        if (!h->chunk_evals) {
            h->chunk_evals = (int64_t *)calloc(2, sizeof(int64_t));
            h->chunk_hits = (int64_t *)calloc(2, sizeof(int64_t));
        }
        run_chunk(h, 0);
        return ORACLE_OK;
    }
    int rc = pool_ensure(h, threads);
    if (rc) return rc;
    pthread_mutex_lock(&h->mu);
    h->pending = h->nworkers;
    h->generation++;
    pthread_cond_broadcast(&h->cv_go);
    while (h->pending > 0) pthread_cond_wait(&h->cv_done, &h->mu); /* future.get() barrier, :59-65 */
    pthread_mutex_unlock(&h->mu);
    return ORACLE_OK;
}

/* ------------------------------------------------------------------ */
/* exported API                                                        */
/* ------------------------------------------------------------------ */

oracle_handle *oracle_create(void)
{
    oracle_handle *h = (oracle_handle *)calloc(1, sizeof(oracle_handle));
    if (!h) return NULL;
    pthread_mutex_init(&h->mu, NULL);
    pthread_cond_init(&h->cv_go, NULL);
    pthread_cond_init(&h->cv_done, NULL);
    h->container_size = 0.065; /* B/Physics.java:25 */
    return h;
}

void oracle_destroy(oracle_handle *h)
{
    if (!h) return;
    pool_stop(h);
    free(h->chunk_evals);
    free(h->chunk_hits);
    free(h->particles);
    free(h->particles_buffer);
    free(h->containers);
    free(h->matrix_copy);
    free(h->nbr_count);
    free(h->nbr_hash);
    free(h->borderline);
    pthread_mutex_destroy(&h->mu);
    pthread_cond_destroy(&h->cv_go);
    pthread_cond_destroy(&h->cv_done);
    free(h);
}

void oracle_set_diag(oracle_handle *h, int on) { h->diag = on; }

/* pos_xy / vel_xy are interleaved (x0,y0,x1,y1,...); z is identically 0
 * (B/Range.java:43,71,86).  id may be NULL (then id[i] = i). */
int oracle_set_particles(oracle_handle *h, int64_t n, const double *pos_xy, const double *vel_xy,
                         const int32_t *type, const uint32_t *id)
{
    if (n < 0) return ORACLE_ERR_INVALID;
    if (n != h->n || !h->particles) {
        free(h->particles);
        free(h->particles_buffer);
        h->particles = (particle *)malloc(sizeof(particle) * (size_t)(n ? n : 1));
        h->particles_buffer = (particle *)malloc(sizeof(particle) * (size_t)(n ? n : 1));
        if (!h->particles || !h->particles_buffer) return ORACLE_ERR_OOM;
        h->n = n;
    }
    for (int64_t i = 0; i < n; i++) {
        particle *p = &h->particles[i];
        p->p[0] = pos_xy[2 * i];
        p->p[1] = pos_xy[2 * i + 1];
        p->p[2] = 0;
        p->v[0] = vel_xy ? vel_xy[2 * i] : 0;
        p->v[1] = vel_xy ? vel_xy[2 * i + 1] : 0;
        p->v[2] = 0;
        p->type = type ? type[i] : 0;
        p->id = id ? id[i] : (uint32_t)i;
    }
    return ORACLE_OK;
}

int oracle_get_particles(const oracle_handle *h, double *pos_xy, double *vel_xy, int32_t *type,
                         uint32_t *id)
{
    for (int64_t i = 0; i < h->n; i++) {
        const particle *p = &h->particles[i];
        if (pos_xy) {
            pos_xy[2 * i] = p->p[0];
            pos_xy[2 * i + 1] = p->p[1];
        }
        if (vel_xy) {
            vel_xy[2 * i] = p->v[0];
            vel_xy[2 * i + 1] = p->v[1];
        }
        if (type) type[i] = p->type;
        if (id) id[i] = p->id;
    }
    return ORACLE_OK;
}

int64_t oracle_count(const oracle_handle *h) { return h->n; }

/* B/Physics.java:112-134: one update() with `threads` = preferredNumberOfThreads (:37) */
int oracle_update(oracle_handle *h, const oracle_settings *s, int threads)
{
    if (!s || !(s->rmax > 0) || s->rmax > 1 || s->m < 1 || !s->matrix || threads < 1 ||
        s->accel_kind < 0 || s->accel_kind >= ACC_KIND_COUNT)
        return ORACLE_ERR_INVALID;
    for (int64_t i = 0; i < h->n; i++)
        if (h->particles[i].type < 0 || h->particles[i].type >= s->m) return ORACLE_ERR_INVALID;

    h->s = *s;
    free(h->matrix_copy);
    h->matrix_copy = (double *)malloc(sizeof(double) * (size_t)s->m * s->m);
    if (!h->matrix_copy) return ORACLE_ERR_OOM;
    memcpy(h->matrix_copy, s->matrix, sizeof(double) * (size_t)s->m * s->m);
    h->s.matrix = h->matrix_copy;

    if (h->diag && h->diag_cap < h->n) {
        free(h->nbr_count);
        free(h->nbr_hash);
        free(h->borderline);
        h->nbr_count = (int32_t *)malloc(sizeof(int32_t) * (size_t)h->n);
        h->nbr_hash = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)h->n);
        h->borderline = (uint8_t *)malloc((size_t)h->n);
        if (!h->nbr_count || !h->nbr_hash || !h->borderline) return ORACLE_ERR_OOM;
        h->diag_cap = h->n;
    }

    int rc = make_containers(h); /* :120 */
    if (rc) return rc;
    rc = distribute_load_evenly(h, threads, 0); /* :122-126 */
    if (rc) return rc;
    h->pair_evals = h->pair_hits = 0;
    for (int c = 0; c < h->nchunks && h->n > 0; c++) {
        h->pair_evals += h->chunk_evals[c];
        h->pair_hits += h->chunk_hits[c];
    }
    rc = distribute_load_evenly(h, threads, 1); /* :127-131 */
    return rc;
}

/* diagnostics of the LAST update, in the post-sort order of that update */
int oracle_grid(const oracle_handle *h, int32_t *nx, int32_t *ny)
{
    *nx = h->nx;
    *ny = h->ny;
    return ORACLE_OK;
}

int oracle_get_containers(const oracle_handle *h, int32_t *out) /* END offsets, nx*ny */
{
    memcpy(out, h->containers, sizeof(int32_t) * (size_t)h->containers_len);
    return ORACLE_OK;
}

int oracle_get_neighbor_diag(const oracle_handle *h, int32_t *count, uint64_t *hash, uint8_t *borderline)
{
    if (!h->diag || h->diag_cap < h->n) return ORACLE_ERR_INVALID;
    if (count) memcpy(count, h->nbr_count, sizeof(int32_t) * (size_t)h->n);
    if (hash) memcpy(hash, h->nbr_hash, sizeof(uint64_t) * (size_t)h->n);
    if (borderline) memcpy(borderline, h->borderline, (size_t)h->n);
    return ORACLE_OK;
}

int oracle_pair_stats(const oracle_handle *h, int64_t *evals, int64_t *hits)
{
    *evals = h->pair_evals;
    *hits = h->pair_hits;
    return ORACLE_OK;
}

/* pure helpers exported for the known-answer tests */
double oracle_range_wrap(double v) { return range_wrap(v); }
double oracle_range_wrap_connection(double v) { return range_wrap_connection(v); }
double oracle_range_clamp(double v) { return range_clamp(v); }
double oracle_particle_life_force(double a, double dist, double beta) { return particle_life_force(a, dist, beta); }

void oracle_accelerate(int kind, const double *params, double a, double *pos_xyz)
{
    oracle_settings s;
    memset(&s, 0, sizeof s);
    s.accel_kind = kind;
    for (int i = 0; i < 4; i++) s.accel_params[i] = params ? params[i] : 0;
    accelerate(&s, a, pos_xyz);
}

int oracle_container_index(double rmax, double x, double y, int32_t *nx_out)
{
    oracle_handle h;
    memset(&h, 0, sizeof h);
    h.container_size = rmax;
    calc_nx_ny(&h);
    double p[3] = {x, y, 0};
    if (nx_out) *nx_out = h.nx;
    return get_container_index(&h, p);
}
