// Particle-set editing on the device: the operations that the reference performs by poking
// `physics.particles` directly, so that a GUI can run against GPU-resident buffers without round-tripping
// all particles:
//   plife_cursor_count   Cursor.countSelection                     (A/cursors/Cursor.java:45-51, A/Main.java:497)
//   plife_cursor_move    cursor action MOVE                        (A/Main.java:540-548)
//   plife_cursor_delete  cursor action DELETE, order preserving    (A/Main.java:568-580)
//   plife_append         cursor action BRUSH / growing particle count (A/Main.java:550-566, B/Physics.java:212-220);
//                        the new particles are sampled on the host by the caller's PositionSetter/TypeSetter
// Cursor.isInside (A/cursors/Cursor.java:16-35) is evaluated in fp64 on the stored position in both precision
// modes: delta = p - cursor; if wrap: delta -= floor(delta + 0.5); delta /= size (JOML: * 1/size);
// circle: |delta| <= 0.5, square: |dx|,|dy| <= 0.5, infinity: everything; size == 0 selects nothing.
#include <math.h>

#include "plife_internal.h"

using namespace plife;

namespace plife {
int edit_fail(plife_handle *h, int code, const char *msg);
int edit_grow(plife_handle *h, int64_t cap);
}

namespace {

constexpr int kThreads = 256;

struct CursorArgs {
    double x, y, inv_size;
    int shape, wrap, empty;
};

__device__ __forceinline__ bool inside(double px, double py, const CursorArgs &c)
{
    if (c.empty) return false;
    double dx = px - c.x, dy = py - c.y;
    if (c.wrap) {
        dx -= floor(dx + 0.5);
        dy -= floor(dy + 0.5);
    }
    dx *= c.inv_size;
    dy *= c.inv_size;
    if (c.shape == PLIFE_CURSOR_CIRCLE) return sqrt(dx * dx + dy * dy) <= 0.5;
    if (c.shape == PLIFE_CURSOR_SQUARE) return fabs(dx) <= 0.5 && fabs(dy) <= 0.5;
    return true; // infinity
}

__device__ __forceinline__ void load_pos(const StateF32 &s, int i, double &x, double &y)
{
    float4 p = s.pt[i];
    x = p.x;
    y = p.y;
}
__device__ __forceinline__ void load_pos(const StateF64 &s, int i, double &x, double &y)
{
    double2 p = s.pos[i];
    x = p.x;
    y = p.y;
}

template <typename S>
__global__ void __launch_bounds__(kThreads) count_inside(S s, int n, CursorArgs c, unsigned long long *total)
{
    __shared__ unsigned int sh[kThreads / 32];
    unsigned int k = 0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        double x, y;
        load_pos(s, i, x, y);
        k += inside(x, y, c) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < kThreads / 32; w++) t += sh[w];
        if (t) atomicAdd(total, t);
    }
}

__device__ __forceinline__ double ensure(double v, int wrap)
{
    if (wrap) return (v < 0 || v >= 1) ? v - floor(v) : v; // B/Range.java:46-57
    return v < 0 ? 0 : (v > 1 ? 1 : v);                    // B/Range.java:89-96
}

__global__ void __launch_bounds__(kThreads) move_inside_f32(StateF32 s, int n, CursorArgs c, double dx, double dy, int wrap_pos)
{
    int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float4 p = s.pt[i];
    if (!inside((double)p.x, (double)p.y, c)) return;
    p.x = (float)ensure((double)p.x + dx, wrap_pos);
    p.y = (float)ensure((double)p.y + dy, wrap_pos);
    s.pt[i] = p;
}
__global__ void __launch_bounds__(kThreads) move_inside_f64(StateF64 s, int n, CursorArgs c, double dx, double dy, int wrap_pos)
{
    int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    double2 p = s.pos[i];
    if (!inside(p.x, p.y, c)) return;
    p.x = ensure(p.x + dx, wrap_pos);
    p.y = ensure(p.y + dy, wrap_pos);
    s.pos[i] = p;
}

// order-preserving compaction of the particles OUTSIDE the cursor: per-block keep counts -> scan -> scatter
template <typename S>
__global__ void __launch_bounds__(kThreads) keep_counts(S s, int n, CursorArgs c, int *block_counts)
{
    int i = blockIdx.x * kThreads + threadIdx.x;
    int keep = 0;
    if (i < n) {
        double x, y;
        load_pos(s, i, x, y);
        keep = inside(x, y, c) ? 0 : 1;
    }
    int cnt = __syncthreads_count(keep);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(1024) scan_blocks(int *block_counts, int nblocks, int *total)
{
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nblocks ? block_counts[i] : 0;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[w] = inc;
        __syncthreads();
        if (w == 0) {
            int ws = warp_sums[lane];
            int winc = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_sums[lane] = winc - ws;
        }
        __syncthreads();
        int carry = carry_s;
        int ex = carry + warp_sums[w] + inc - v;
        if (i < nblocks) block_counts[i] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(kThreads) compact_f32(StateF32 in, StateF32 out, int n, CursorArgs c, const int *block_offsets)
{
    __shared__ int warp_sums[kThreads / 32];
    int i = blockIdx.x * kThreads + threadIdx.x;
    float4 p = make_float4(0, 0, 0, 0);
    int keep = 0;
    if (i < n) {
        p = in.pt[i];
        keep = inside((double)p.x, (double)p.y, c) ? 0 : 1;
    }
    unsigned m = __ballot_sync(0xffffffffu, keep);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_sums[w] = __popc(m);
    __syncthreads();
    int off = block_offsets[blockIdx.x];
    for (int k = 0; k < w; k++) off += warp_sums[k];
    off += __popc(m & ((1u << lane) - 1u));
    if (keep) {
        out.pt[off] = p;
        out.vel[off] = in.vel[i];
    }
}
__global__ void __launch_bounds__(kThreads) compact_f64(StateF64 in, StateF64 out, int n, CursorArgs c, const int *block_offsets)
{
    __shared__ int warp_sums[kThreads / 32];
    int i = blockIdx.x * kThreads + threadIdx.x;
    double2 p = make_double2(0, 0);
    int keep = 0;
    if (i < n) {
        p = in.pos[i];
        keep = inside(p.x, p.y, c) ? 0 : 1;
    }
    unsigned m = __ballot_sync(0xffffffffu, keep);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_sums[w] = __popc(m);
    __syncthreads();
    int off = block_offsets[blockIdx.x];
    for (int k = 0; k < w; k++) off += warp_sums[k];
    off += __popc(m & ((1u << lane) - 1u));
    if (keep) {
        out.pos[off] = p;
        out.vel[off] = in.vel[i];
        out.type[off] = in.type[i];
        out.id[off] = in.id[i];
    }
}

int make_cursor(plife_handle *h, const plife_cursor *c, CursorArgs *a)
{
    if (!c) return edit_fail(h, PLIFE_ERR_INVALID, "cursor is NULL");
    if (c->shape < PLIFE_CURSOR_CIRCLE || c->shape > PLIFE_CURSOR_INFINITY || !isfinite(c->x) || !isfinite(c->y) || !(c->size >= 0))
        return edit_fail(h, PLIFE_ERR_INVALID, "bad cursor");
    a->x = c->x;
    a->y = c->y;
    a->empty = c->size == 0.0; // A/cursors/Cursor.java:17
    a->inv_size = a->empty ? 0.0 : 1.0 / c->size;
    a->shape = c->shape;
    a->wrap = c->wrap ? 1 : 0;
    return PLIFE_OK;
}

// plife_rebuild: new particle k = old particle src[k] (src[k] >= 0) or a new one; its type is type[k]; where place[k] >= 0 the
// position is placed[place[k]] and the velocity zero (Physics.setPosition, B/Physics.java:297-303)
__global__ void __launch_bounds__(kThreads) rebuild_f32(StateF32 in, StateF32 out, int n_new, const int32_t *__restrict__ src,
                                                        const int32_t *__restrict__ type, const int32_t *__restrict__ place,
                                                        const double2 *__restrict__ placed, uint32_t next_id)
{
    const int k = blockIdx.x * kThreads + threadIdx.x;
    if (k >= n_new) return;
    const int s = src[k], pl = place ? place[k] : -1;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 v = make_float2(0.f, 0.f);
    if (s >= 0) {
        p = in.pt[s];
        v = in.vel[s];
    } else {
        p.w = __uint_as_float(next_id + (uint32_t)pl);
    }
    p.z = __int_as_float(type[k]);
    if (pl >= 0) {
        const double2 q = placed[pl];
        p.x = (float)q.x;
        p.y = (float)q.y;
        v = make_float2(0.f, 0.f);
    }
    out.pt[k] = p;
    out.vel[k] = v;
}

__global__ void __launch_bounds__(kThreads) rebuild_f64(StateF64 in, StateF64 out, int n_new, const int32_t *__restrict__ src,
                                                        const int32_t *__restrict__ type, const int32_t *__restrict__ place,
                                                        const double2 *__restrict__ placed, uint32_t next_id)
{
    const int k = blockIdx.x * kThreads + threadIdx.x;
    if (k >= n_new) return;
    const int s = src[k], pl = place ? place[k] : -1;
    double2 p = make_double2(0, 0), v = make_double2(0, 0);
    uint32_t id = next_id + (uint32_t)(pl < 0 ? 0 : pl);
    if (s >= 0) {
        p = in.pos[s];
        v = in.vel[s];
        id = in.id[s];
    }
    if (pl >= 0) {
        p = placed[pl];
        v = make_double2(0, 0);
    }
    out.pos[k] = p;
    out.vel[k] = v;
    out.type[k] = type[k];
    out.id[k] = id;
}

int check(plife_handle *h)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (h->poisoned) return edit_fail(h, PLIFE_ERR_CUDA, "handle poisoned by an earlier CUDA error");
    if (h->slab.on) return edit_fail(h, PLIFE_ERR_STATE, "editing operations are not available in slab mode");
    if (cudaSetDevice(h->device) != cudaSuccess) return edit_fail(h, PLIFE_ERR_CUDA, "cudaSetDevice");
    return PLIFE_OK;
}

#define CUE(h, expr)                                                \
    do {                                                            \
        cudaError_t e_ = (expr);                                    \
        if (e_ != cudaSuccess) {                                    \
            h->poisoned = true;                                     \
            return edit_fail(h, PLIFE_ERR_CUDA, cudaGetErrorString(e_)); \
        }                                                           \
    } while (0)

} // namespace

extern "C" {

int plife_cursor_count(plife_handle *h, const plife_cursor *c, int64_t *out)
{
    int rc = check(h);
    if (rc) return rc;
    CursorArgs a;
    if ((rc = make_cursor(h, c, &a))) return rc;
    if (!out) return edit_fail(h, PLIFE_ERR_INVALID, "out is NULL");
    const int n = (int)h->n;
    unsigned long long *d_total = h->d_scalar;
    unsigned long long total = 0;
    CUE(h, cudaMemsetAsync(d_total, 0, sizeof total, h->stream));
    if (n > 0) {
        int nb = (n + kThreads - 1) / kThreads;
        if (nb > 148 * 16) nb = 148 * 16;
        if (h->precision == PLIFE_F32) count_inside<<<nb, kThreads, 0, h->stream>>>(h->s32[h->cur], n, a, d_total);
        else count_inside<<<nb, kThreads, 0, h->stream>>>(h->s64[h->cur], n, a, d_total);
        CUE(h, cudaGetLastError());
    }
    CUE(h, cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, h->stream));
    CUE(h, cudaStreamSynchronize(h->stream));
    *out = (int64_t)total;
    return PLIFE_OK;
}

int plife_cursor_move(plife_handle *h, const plife_cursor *c, double dx, double dy)
{
    int rc = check(h);
    if (rc) return rc;
    CursorArgs a;
    if ((rc = make_cursor(h, c, &a))) return rc;
    if (!isfinite(dx) || !isfinite(dy)) return edit_fail(h, PLIFE_ERR_INVALID, "move: delta not finite");
    const int n = (int)h->n;
    if (n == 0) return PLIFE_OK;
    const int nb = (n + kThreads - 1) / kThreads;
    const int wrap_pos = h->settings.wrap ? 1 : 0; // physics.ensurePosition (B/Physics.java:499-505)
    if (h->precision == PLIFE_F32) move_inside_f32<<<nb, kThreads, 0, h->stream>>>(h->s32[h->cur], n, a, dx, dy, wrap_pos);
    else move_inside_f64<<<nb, kThreads, 0, h->stream>>>(h->s64[h->cur], n, a, dx, dy, wrap_pos);
    CUE(h, cudaGetLastError());
    h->prebinned = false; // positions changed behind the fused binning
    h->has_sorted = false;
    return PLIFE_OK;
}

int plife_cursor_delete(plife_handle *h, const plife_cursor *c, int64_t *removed)
{
    int rc = check(h);
    if (rc) return rc;
    CursorArgs a;
    if ((rc = make_cursor(h, c, &a))) return rc;
    const int n = (int)h->n;
    if (removed) *removed = 0;
    if (n == 0) return PLIFE_OK;
    const int nb = (n + kThreads - 1) / kThreads;
    int *d_blocks = h->d_perm; // scratch: nb <= n ints, not live between steps
    int *d_total = reinterpret_cast<int *>(h->d_scalar + 1);
    const int src = h->cur, dst = h->cur ^ 1;
    if (h->precision == PLIFE_F32) keep_counts<<<nb, kThreads, 0, h->stream>>>(h->s32[src], n, a, d_blocks);
    else keep_counts<<<nb, kThreads, 0, h->stream>>>(h->s64[src], n, a, d_blocks);
    scan_blocks<<<1, 1024, 0, h->stream>>>(d_blocks, nb, d_total);
    if (h->precision == PLIFE_F32) compact_f32<<<nb, kThreads, 0, h->stream>>>(h->s32[src], h->s32[dst], n, a, d_blocks);
    else compact_f64<<<nb, kThreads, 0, h->stream>>>(h->s64[src], h->s64[dst], n, a, d_blocks);
    CUE(h, cudaGetLastError());
    int kept = 0;
    CUE(h, cudaMemcpyAsync(&kept, d_total, sizeof kept, cudaMemcpyDeviceToHost, h->stream));
    CUE(h, cudaStreamSynchronize(h->stream));
    h->cur = dst;
    h->n = h->n_phys = kept;
    h->slab.n_old = kept;
    h->prebinned = false;
    h->has_sorted = false;
    if (kept == 0) h->max_type = -1;
    if (removed) *removed = n - kept;
    return PLIFE_OK;
}

int plife_append(plife_handle *h, int64_t k, const double *pos_xy, const double *vel_xy, const int32_t *type)
{
    int rc = check(h);
    if (rc) return rc;
    if (k < 0 || (k > 0 && (!pos_xy || !type))) return edit_fail(h, PLIFE_ERR_INVALID, "append: bad arguments");
    if (k == 0) return PLIFE_OK;
    int max_type = h->max_type;
    for (int64_t i = 0; i < k; i++) {
        double x = pos_xy[2 * i], y = pos_xy[2 * i + 1];
        if (!(x >= 0 && x <= 1 && y >= 0 && y <= 1)) return edit_fail(h, PLIFE_ERR_INVALID, "append: position outside [0,1]^2");
        if (type[i] < 0 || type[i] >= h->m) return edit_fail(h, PLIFE_ERR_INVALID, "append: type outside [0,m)");
        if (type[i] > max_type) max_type = type[i];
    }
    const int64_t n = h->n;
    if (n + k > h->cap) {
        rc = edit_grow(h, n + k + (n + k) / 4);
        if (rc) return rc;
    }
    CUE(h, cudaStreamSynchronize(h->stream));
    if (h->precision == PLIFE_F32) {
        std::vector<float4> pt((size_t)k);
        std::vector<float2> vl((size_t)k);
        for (int64_t i = 0; i < k; i++) {
            uint32_t id = h->next_id++;
            float4 q;
            q.x = (float)pos_xy[2 * i];
            q.y = (float)pos_xy[2 * i + 1];
            memcpy(&q.z, &type[i], 4);
            memcpy(&q.w, &id, 4);
            pt[i] = q;
            vl[i] = vel_xy ? make_float2((float)vel_xy[2 * i], (float)vel_xy[2 * i + 1]) : make_float2(0.f, 0.f);
        }
        CUE(h, cudaMemcpy(h->s32[h->cur].pt + n, pt.data(), sizeof(float4) * k, cudaMemcpyHostToDevice));
        CUE(h, cudaMemcpy(h->s32[h->cur].vel + n, vl.data(), sizeof(float2) * k, cudaMemcpyHostToDevice));
    } else {
        std::vector<double2> vl((size_t)k, make_double2(0, 0));
        std::vector<uint32_t> ids((size_t)k);
        for (int64_t i = 0; i < k; i++) {
            ids[i] = h->next_id++;
            if (vel_xy) vl[i] = make_double2(vel_xy[2 * i], vel_xy[2 * i + 1]);
        }
        const StateF64 &s = h->s64[h->cur];
        CUE(h, cudaMemcpy(s.pos + n, pos_xy, sizeof(double2) * k, cudaMemcpyHostToDevice));
        CUE(h, cudaMemcpy(s.vel + n, vl.data(), sizeof(double2) * k, cudaMemcpyHostToDevice));
        CUE(h, cudaMemcpy(s.type + n, type, sizeof(int32_t) * k, cudaMemcpyHostToDevice));
        CUE(h, cudaMemcpy(s.id + n, ids.data(), sizeof(uint32_t) * k, cudaMemcpyHostToDevice));
    }
    h->n = h->n_phys = n + k;
    h->slab.n_old = h->n;
    h->max_type = max_type;
    h->prebinned = false;
    h->has_sorted = false;
    return PLIFE_OK;
}

// Physics.setParticleCount (shuffle-before-shrink, B/Physics.java:190-223, :278-280), ensureTypes (:266-272), setTypes (:509-511),
// ExtendedPhysics.setTypeCount / setTypeCountEqual (A/ExtendedPhysics.java:28-130): the caller plans the new array on the
// host - from the types alone, 4 bytes per particle - and the device applies the plan; particles never leave the GPU.
int plife_rebuild(plife_handle *h, int64_t n_new, const int32_t *src, const int32_t *type, const int32_t *place, int64_t n_placed,
                  const double *placed_xy)
{
    int rc = check(h);
    if (rc) return rc;
    if (n_new < 0 || n_placed < 0 || (n_new > 0 && (!src || !type)) || (n_placed > 0 && (!place || !placed_xy)))
        return edit_fail(h, PLIFE_ERR_INVALID, "rebuild: bad arguments");
    const int64_t n_old = h->n;
    int max_type = -1;
    for (int64_t k = 0; k < n_new; k++) {
        if (src[k] >= n_old) return edit_fail(h, PLIFE_ERR_INVALID, "rebuild: src index beyond the particle count");
        if (type[k] < 0 || type[k] >= h->m) return edit_fail(h, PLIFE_ERR_INVALID, "rebuild: type outside [0,m)");
        const int pl = place ? place[k] : -1;
        if (pl >= n_placed) return edit_fail(h, PLIFE_ERR_INVALID, "rebuild: place index beyond the placed positions");
        if (src[k] < 0 && pl < 0) return edit_fail(h, PLIFE_ERR_INVALID, "rebuild: a new particle needs a position");
        if (type[k] > max_type) max_type = type[k];
    }
    for (int64_t q = 0; q < n_placed; q++) {
        const double x = placed_xy[2 * q], y = placed_xy[2 * q + 1];
        if (!(x >= 0 && x <= 1 && y >= 0 && y <= 1)) return edit_fail(h, PLIFE_ERR_INVALID, "rebuild: position outside [0,1]^2");
    }
    if (n_new > h->cap) {
        rc = edit_grow(h, n_new + n_new / 4);
        if (rc) return rc;
    }
    int32_t *d_plan = nullptr;
    double2 *d_placed = nullptr;
    const size_t np = (size_t)(n_new ? n_new : 1);
    CUE(h, cudaMalloc((void **)&d_plan, sizeof(int32_t) * 3 * np));
    cudaError_t e = cudaMalloc((void **)&d_placed, sizeof(double2) * (size_t)(n_placed ? n_placed : 1));
    if (e == cudaSuccess && n_new) e = cudaMemcpyAsync(d_plan, src, sizeof(int32_t) * n_new, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && n_new) e = cudaMemcpyAsync(d_plan + np, type, sizeof(int32_t) * n_new, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && n_new && place) e = cudaMemcpyAsync(d_plan + 2 * np, place, sizeof(int32_t) * n_new, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && n_placed) e = cudaMemcpyAsync(d_placed, placed_xy, sizeof(double2) * n_placed, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && n_new) {
        const int nb = (int)((n_new + kThreads - 1) / kThreads);
        const int a = h->cur, b = h->cur ^ 1;
        if (h->precision == PLIFE_F32)
            rebuild_f32<<<nb, kThreads, 0, h->stream>>>(h->s32[a], h->s32[b], (int)n_new, d_plan, d_plan + np, place ? d_plan + 2 * np : nullptr, d_placed, h->next_id);
        else
            rebuild_f64<<<nb, kThreads, 0, h->stream>>>(h->s64[a], h->s64[b], (int)n_new, d_plan, d_plan + np, place ? d_plan + 2 * np : nullptr, d_placed, h->next_id);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_plan);
    cudaFree(d_placed);
    if (e != cudaSuccess) {
        h->poisoned = true;
        return edit_fail(h, PLIFE_ERR_CUDA, cudaGetErrorString(e));
    }
    h->cur ^= 1;
    h->n = h->n_phys = n_new;
    h->slab.n_old = n_new;
    h->next_id += (uint32_t)n_placed;
    h->max_type = max_type;
    h->prebinned = false;
    h->has_sorted = false;
    return PLIFE_OK;
}

} // extern "C"
