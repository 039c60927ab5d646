/*
 * plife.h -- C ABI of the B200-native Particle Life physics step.
 *
 * Drop-in boundary for the reference's `Physics` object
 * (src/main/java/com/particle_life/backend/Physics.java, "B/" below; "A/" is
 * .../particle_life/app/).  Each entry point names the reference interface it
 * replaces.  A JNI / Java-FFM shim binds exactly these symbols (see
 * INTEGRATION.md); nothing here exposes torch or C++ types.
 *
 * Conventions
 *   - every function returns an int status (PLIFE_OK == 0, negative = error)
 *     unless documented otherwise; nothing throws across the ABI;
 *   - the caller owns host buffers, the library owns device buffers; no host
 *     pointer is retained after a call returns;
 *   - a handle is single-owner: one calling thread at a time (the reference
 *     drives Physics from its single Loop thread, B/Loop.java:102-121).
 *     plife_request_stop() and plife_last_error() may be called from any thread;
 *   - particle order is the reference's: after every step the particle array is
 *     in stable cell-sorted order (B/Physics.java:343-353).
 *   - there is NO CPU fallback: without a usable CUDA device plife_create()
 *     fails with PLIFE_ERR_CUDA.
 */
#ifndef PLIFE_H
#define PLIFE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLIFE_VERSION 200 /* 0.2.0: plife_config.bins, plife_step_stats.graph_steps, plife_rebuild, host-free slab steps */

/* status codes */
#define PLIFE_OK 0
#define PLIFE_ERR_INVALID (-1) /* bad argument (IllegalArgumentException in the reference) */
#define PLIFE_ERR_OOM (-2)     /* device or host allocation failed */
#define PLIFE_ERR_CUDA (-3)    /* CUDA runtime error; sticky errors poison the handle */
#define PLIFE_ERR_NCCL (-4)    /* reserved for the multi-GPU exchange */
#define PLIFE_ERR_STATE (-5)   /* call not valid in the handle's current state */
#define PLIFE_ERR_STOPPED (-6) /* plife_request_stop() interrupted a multi-step call */

/* arithmetic / storage precision of a handle */
#define PLIFE_F32 0 /* fp32 storage and arithmetic; cell assignment still in fp64 */
#define PLIFE_F64 1 /* fp64 throughout; bit-exact with the reference's operation order */

/* plife_config.flags */
#define PLIFE_FLAG_UNSTABLE_SORT 1 /* skip the in-cell stable ordering (faster; order in a cell arbitrary) */
#define PLIFE_FLAG_NO_GRAPH 2      /* never replay the step as a CUDA graph (default: captured below 262144 particles, once the
                                      settings have been stable for two steps; any change of dt, settings or particles re-captures) */
/* 16 was PLIFE_FLAG_PAIRS (round 1's experimental two-targets-per-lane kernel, removed); the bit is ignored */
#define PLIFE_FLAG_SCAN3 32         /* exclusive scan over the bins as three launches (tile sums, scan of sums, apply) instead of one */
#define PLIFE_FLAG_NO_CELLS 64      /* fp32, at most 65536 particles: use the staged force kernel instead of the warp-per-cell one */
#define PLIFE_FLAG_NO_FUSED_BIN 8  /* do not fuse the next step's binning into the force pass */
#define PLIFE_FLAG_FORCE_V1 4      /* fp32: use the global-memory force kernel instead of the shared-memory staged one */

/* Accelerator kinds (B/Accelerator.java:5-17).  Kind 0 is the only accelerator
 * in the reference snapshot (A/Main.java:275-280); kinds 1..5 are
 * builder-defined (no reference definition). */
#define PLIFE_ACC_PARTICLE_LIFE 0    /* params = {beta (default 0.3)} */
#define PLIFE_ACC_PARTICLE_LIFE_R 1  /* kind 0 force, divided by r once more */
#define PLIFE_ACC_PARTICLE_LIFE_R2 2 /* kind 0 force, divided by r^2 more */
#define PLIFE_ACC_ROTATOR_90 3       /* a*(1-d), rotated by 90 degrees */
#define PLIFE_ACC_ROTATOR_ATTR 4     /* (1-d), rotated by -a*pi */
#define PLIFE_ACC_PLANETS 5          /* 0.01 / max(d, 0.01)^2 attraction */
#define PLIFE_ACC_KIND_COUNT 6

typedef struct plife_handle plife_handle;

typedef struct plife_config {
    int32_t device;    /* CUDA device ordinal */
    int32_t precision; /* PLIFE_F32 | PLIFE_F64 */
    int64_t capacity;  /* particle capacity hint (buffers grow on upload) */
    int32_t flags;     /* PLIFE_FLAG_* */
    int32_t bins;      /* fp32: fine bins per cell along x for the internal cell list: 0 = chosen from the density, or 1, 2, 4, 8
                          (results and the particle order do not depend on it, only the speed) */
    void *stream;      /* cudaStream_t to run on, or NULL: the library creates its own */
} plife_config;

/* B/PhysicsSettings.java:8-37 minus `dt` (per step) and `matrix` (own setter) */
typedef struct plife_settings {
    double rmax;     /* :13, 0 < rmax <= 1 (rmax > 1 makes nx = 0: the reference throws, B/Physics.java:362-374) */
    double friction; /* :27 */
    double force;    /* :32 */
    int32_t wrap;    /* :8  */
    int32_t reserved;
} plife_settings;

/* counters of the most recent step */
typedef struct plife_step_stats {
    int64_t n;          /* particles */
    int32_t nx, ny;     /* grid, B/Physics.java:82-85 */
    int64_t pair_evals; /* candidate pairs (i,j), j != i, in the 3x3 cells: B/Physics.java:423-439 */
    int64_t steps;      /* steps executed by this handle so far */
    int64_t graph_steps; /* ... of which replayed from a captured CUDA graph (launch-bound regime, see PLIFE_FLAG_NO_GRAPH) */
} plife_step_stats;

/* slots of plife_kernel_times() */
#define PLIFE_K_BIN 0     /* cell id + histogram */
#define PLIFE_K_SCAN 1    /* exclusive scan over cells */
#define PLIFE_K_SCATTER 2 /* cursor scatter of source indices */
#define PLIFE_K_GATHER 3  /* stable in-cell rank + reorder */
#define PLIFE_K_FORCE 4   /* 3x3 force + friction + integrate + wrap */
#define PLIFE_K_COUNT 5

int plife_version(void);
const char *plife_status_string(int status);

/* new ExtendedPhysics(...) / Physics ctor (A/Main.java:281-285, B/Physics.java:65-80).
 * Defaults after create: PhysicsSettings defaults (rmax 0.02, friction 0.85,
 * force 1.0, wrap on), accelerator kind 0 with beta 0.3, a 1x1 zero matrix and
 * zero particles; the host-side setters produce the initial state and upload it. */
int plife_create(const plife_config *cfg, plife_handle **out);
/* Physics.kill() + GC (B/Physics.java:156-158) */
int plife_destroy(plife_handle *h);

/* writes to physics.settings.{rmax,friction,force,wrap} (A/Main.java:811,818,827,834) */
int plife_set_settings(plife_handle *h, const plife_settings *s);
int plife_get_settings(const plife_handle *h, plife_settings *out);

/* settings.matrix = ... (A/Main.java:714,1330), Physics.setMatrixSize (B/Physics.java:238-258).
 * row_major[i*m+j] = matrix.get(i,j): i = own type, j = other type (B/Physics.java:435).
 * 1 <= m <= 256 (A/Main.java:745).  PLIFE_ERR_STATE if a resident particle has
 * type >= m (the reference's ensureTypes, B/Physics.java:266-272, is host-side:
 * download, retype, upload). */
int plife_set_matrix(plife_handle *h, int32_t m, const double *row_major);
/* settings.matrix.set(i,j,v) (A/Main.java:703) */
int plife_set_matrix_entry(plife_handle *h, int32_t i, int32_t j, double v);
int plife_get_matrix(const plife_handle *h, int32_t *m_out, double *row_major_out, int32_t capacity_m);

/* physics.accelerator = ... (B/Physics.java:27,70).  A Java lambda cannot run on
 * the device, so accelerators are device functors selected by kind. */
int plife_set_accelerator(plife_handle *h, int32_t kind, const double *params, int32_t nparams);

/* physics.particles = ... / setParticleCount / setPositions / setTypes
 * (B/Physics.java:15,166,190,509).  Host SoA in fp64: pos_xy and vel_xy are
 * interleaved (x0,y0,x1,y1,...); z is identically 0 (B/Range.java:43,71,86).
 * vel_xy may be NULL (zero velocities, B/Physics.java:300-302).  id may be
 * NULL (ids 0..n-1).  Validates 0 <= x,y <= 1 and 0 <= type < m. */
int plife_upload(plife_handle *h, int64_t n, const double *pos_xy, const double *vel_xy,
                 const int32_t *type, const uint32_t *id);
/* PhysicsSnapshot.take (A/PhysicsSnapshot.java:24-62).  Any pointer may be NULL.
 * Order = current particle order (cell-sorted after a step). Synchronises. */
int plife_download(plife_handle *h, double *pos_xy, double *vel_xy, int32_t *type, uint32_t *id);
/* float snapshot for rendering: xy + vxy as fp32 (display-time handoff) */
int plife_download_f32(plife_handle *h, float *pos_xy, float *vel_xy, int32_t *type);

/* The same snapshot without stalling the physics (A/Main.java:600-603 requests the next snapshot while it draws):
 * a snapshot kernel on the compute stream fills one of two staging buffers, the device->host copies run on a
 * separate copy stream and overlap the following steps.  The caller's buffers (pinned for full PCIe speed) must
 * stay valid until plife_snapshot_wait() returns. */
int plife_snapshot_async(plife_handle *h, float *pos_xy, float *vel_xy, int32_t *type);
/* The same with the types as one byte each (the app allows at most 256 types, A/Main.java:745): 17 instead of
 * 20 bytes per particle over PCIe; a GL renderer binds it as an unsigned-byte integer attribute. */
int plife_snapshot_async_u8(plife_handle *h, float *pos_xy, float *vel_xy, uint8_t *type_u8);
/* Waits for the OLDEST snapshot requested and not yet waited for.  Up to two requests may be in flight: request snapshot
 * k + 1 before waiting for snapshot k and the copy engine never idles (a third request first waits for the oldest). */
int plife_snapshot_wait(plife_handle *h);

/* Headless generators with the distributions of the reference's default setters
 * (B/DefaultPositionSetter.java, B/DefaultTypeSetter.java, B/DefaultMatrix.java:25-31)
 * on a seeded SplitMix64 stream; bit-identical to plife/synth.py. */
int plife_init_uniform(plife_handle *h, int64_t n, uint64_t seed);
int plife_random_matrix(plife_handle *h, int32_t m, uint64_t seed);

/* physics.settings.dt = dt; physics.update() (A/Main.java:292-295, B/Physics.java:112),
 * nsteps times.  Asynchronous on the handle's stream. */
int plife_step(plife_handle *h, double dt, int32_t nsteps);
int plife_sync(plife_handle *h);

/* physics.particles.length */
int64_t plife_count(const plife_handle *h);
/* ExtendedPhysics.getTypeCount (A/ExtendedPhysics.java:19-26); out has m entries */
int plife_type_histogram(plife_handle *h, int64_t *out_m);

/* Physics.forceUpdateStop (B/Physics.java:149-151): honoured between steps of a
 * multi-step plife_step(); thread-safe. */
int plife_request_stop(plife_handle *h);
/* message of the last failing call on this handle; owned by the handle */
const char *plife_last_error(const plife_handle *h);

/* ---- parity / measurement instrumentation ---- */

/* `containers` of the most recent step (B/Physics.java:18): END offset of every
 * cell, nx*ny int32.  Synchronises. */
int plife_get_containers(plife_handle *h, int32_t *out, int64_t capacity);
int plife_get_step_stats(plife_handle *h, plife_step_stats *out);
/* In-range neighbour set of every particle for the CURRENT state and settings
 * (re-bins first; does not advance time): count and an order-independent hash
 * (sum of mix64(id_j)) per particle, in the sorted order.  Synchronises. */
int plife_debug_neighbors(plife_handle *h, int32_t *count, uint64_t *hash);
/* Per-kernel device time: enable, run steps, read accumulated milliseconds and
 * launch counts (PLIFE_K_* slots).  Profiling inserts events between kernels.
 * In slab mode the PLIFE_K_GATHER slot also covers the halo pack / wait / unpack
 * kernels that run between the gather and the force pass; PLIFE_K_FORCE is the
 * force kernel alone. */
int plife_set_profiling(plife_handle *h, int32_t enabled);
int plife_kernel_times(plife_handle *h, double *ms_out, int64_t *launches_out);
/* Measured FP32 peak of a device: an FFMA loop (8 independent chains per thread, 2 FLOP per FFMA), best of 6
 * launches.  The denominator of bench.py's FP32 fraction (MEASURED_PEAKS.json carries no FP32 figure). */
int plife_measure_fp32_peak(int32_t device, double *tflops_out);
/* device pointers of the current state (for CUDA-GL interop or torch views):
 * F32: pos = float4{x,y,type bits,id bits}[n], vel = float2[n]
 * F64: pos = double2[n], vel = double2[n], type = int32[n], id = uint32[n]
 * The pointers are valid until the next call that steps, edits or uploads: query again afterwards (the
 * velocity buffers of an F32 handle alternate from step to step). */
int plife_device_ptrs(plife_handle *h, void **pos, void **vel, void **type, void **id);

/* ---- particle-set editing on the device (what the GUI does to physics.particles directly) ----
 * Cursor.isInside (A/cursors/Cursor.java:16-35): delta = p - (x,y); if wrap: delta -= floor(delta + 0.5);
 * delta /= size; circle |delta| <= 0.5, square |dx|,|dy| <= 0.5, infinity selects everything; size 0 nothing. */
#define PLIFE_CURSOR_CIRCLE 0
#define PLIFE_CURSOR_SQUARE 1
#define PLIFE_CURSOR_INFINITY 2
typedef struct plife_cursor {
    double x, y, size;
    int32_t shape; /* PLIFE_CURSOR_* */
    int32_t wrap;  /* physics.settings.wrap at the time of the call */
} plife_cursor;
/* Cursor.countSelection (A/cursors/Cursor.java:45-51; every frame at A/Main.java:497) */
int plife_cursor_count(plife_handle *h, const plife_cursor *c, int64_t *out);
/* cursor action MOVE (A/Main.java:540-548): position += (dx,dy), then ensurePosition with the handle's wrap setting */
int plife_cursor_move(plife_handle *h, const plife_cursor *c, double dx, double dy);
/* cursor action DELETE (A/Main.java:568-580): removes the selected particles, keeps the order of the others */
int plife_cursor_delete(plife_handle *h, const plife_cursor *c, int64_t *removed);
/* cursor action BRUSH / growing setParticleCount (A/Main.java:550-566, B/Physics.java:212-220): appends k particles
 * sampled by the caller's setters; ids continue after the largest id seen; vel_xy may be NULL */
int plife_append(plife_handle *h, int64_t k, const double *pos_xy, const double *vel_xy, const int32_t *type);

/* Physics.setParticleCount incl. its shuffle-before-shrink (B/Physics.java:190-223, :278-280), ensureTypes (:266-272), setTypes
 * (:509-511), ExtendedPhysics.setTypeCount / setTypeCountEqual (A/ExtendedPhysics.java:28-130) all rearrange, retype, drop or
 * create particles by rules that depend on the TYPES only (plus the host-side setter plugins for new positions).  The host
 * plans the new array from the types (plife_download_f32 with only `type`: 4 bytes per particle) and the device applies it:
 *   new particle k = old particle src[k] (src[k] >= 0; its id, position and velocity are kept) or a new one (src[k] < 0);
 *   type[k] is its type; if place[k] >= 0 (place may be NULL: nobody is placed) its position becomes placed_xy[place[k]] and
 *   its velocity zero, as Physics.setPosition does (B/Physics.java:297-303).  New particles must be placed; they get the ids
 *   next_id + place[k].  n_placed = number of (x, y) pairs in placed_xy. */
int plife_rebuild(plife_handle *h, int64_t n_new, const int32_t *src, const int32_t *type, const int32_t *place, int64_t n_placed,
                  const double *placed_xy);

/* ---- multi-GPU slab decomposition (one process per GPU) ----
 * Rank g owns grid rows [g*ny/G, (g+1)*ny/G).  The library packs / unpacks the halo and migration
 * messages; the host exchanges them between the phases (NCCL send/recv, e.g. torch.distributed):
 *   halo_send[0] -> down neighbour's halo_recv[1],  halo_send[1] -> up neighbour's halo_recv[0]
 *   mig_send[0]  -> down neighbour's mig_recv[1],   mig_send[1]  -> up neighbour's mig_recv[0]
 * (down = rank-1, up = rank+1, periodic when wrap is on; no exchange across a closed boundary).
 * Displacement bound (the reference has none): the migration exchange of a step runs while the interior rows are
 * still being computed, so a particle may enter a neighbour slab only from the slab's first or last row - i.e. it must move
 * less than one grid row (rmax) per step towards the neighbour - and never further than the neighbour: PLIFE_ERR_STATE otherwise.
 * External exchange (bufs != NULL): buffers are device memory owned by the caller, in 16-byte records:
 * plife_slab_halo_records(nx, halo_cap) / plife_slab_migrate_records(mig_cap) records each.
 * fp32 handles only.  plife_upload accepts only particles of the rank's own rows (PLIFE_ERR_INVALID otherwise);
 * plife_init_uniform keeps the rank's share of the global stream. */
typedef struct plife_slab_buffers {
    void *halo_send[2], *halo_recv[2], *mig_send[2], *mig_recv[2];
} plife_slab_buffers;

/* A slab step never synchronises with the host: the particle counts of a rank (they change with migration) live in
 * device memory, kernels are sized by an upper bound, and the host reads the counts back a few steps late.  So an error
 * the device finds in a step (halo / migration overflow, a particle crossing more than one slab, a dead neighbour)
 * is returned by a LATER call: the next plife_slab_phase that sees it, or any synchronising call (plife_sync,
 * plife_count, plife_download, snapshots).  plife_count() drains the queued steps first. */
#define PLIFE_SLAB_SORT 0   /* cell-list build of the owned particles + pack halo rows */
#define PLIFE_SLAB_FORCE 1  /* force + integrate: interior rows, then place the ghost rows, then the two edge rows; pack leavers */
#define PLIFE_SLAB_FINISH 2 /* new counts on the device, append arrivals */

int64_t plife_slab_halo_records(int32_t nx, int64_t halo_cap);
int64_t plife_slab_migrate_records(int64_t mig_cap);
int plife_slab_configure(plife_handle *h, int32_t rank, int32_t world, int64_t halo_cap, int64_t mig_cap,
                         const plife_slab_buffers *bufs);
/* Peer exchange (bufs == NULL in plife_slab_configure): the library owns the buffers; every rank exports a
 * 64-byte CUDA IPC handle, the host distributes them (any transport), and each rank maps its neighbours'
 * receive slots.  The step then needs no host-side exchange: kernels push the messages over NVLink and
 * signal with flags.  NULL = no neighbour in that direction (closed boundary). */
int plife_slab_export(plife_handle *h, void *ipc_handle_64_bytes);
int plife_slab_connect_ipc(plife_handle *h, const void *down_handle, const void *up_handle);
int plife_slab_connect_local(plife_handle *h, plife_handle *down, plife_handle *up);
/* SORT, FORCE, FINISH back to back, nsteps times (peer exchange only) */
int plife_slab_step(plife_handle *h, double dt, int32_t nsteps);
/* rows [row_lo, row_hi) owned by this rank under the current settings, and nx */
int plife_slab_rows(plife_handle *h, int32_t *row_lo, int32_t *row_hi, int32_t *nx);
int plife_slab_phase(plife_handle *h, int32_t phase, double dt);

#ifdef __cplusplus
}
#endif
#endif /* PLIFE_H */
