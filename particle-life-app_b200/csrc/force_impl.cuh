// Velocity + position pass of one step (Physics.updateVelocity / updatePosition,
// B/Physics.java:397-450) as ONE kernel: thread i owns sorted particle i, walks
// the 3x3 cells around floor(pos/rmax), accumulates the accelerator output,
// applies friction, integrates and wraps/clamps, and writes the new state into
// the other buffer (Jacobi: every force uses the old positions, which is what
// the reference's barrier between its two passes guarantees, :122-131).
//
// Included by force_f32.cu (fused multiply-add on) and force_f64.cu
// (compiled with -fmad=false so every operation rounds exactly like the
// reference's non-fused JOML arithmetic).
#pragma once

#include "plife_internal.h"
#include <type_traits>

namespace plife {

constexpr int kForceThreads = 128;

template <typename R>
struct Cand {
    R x, y;
    int type;
    uint32_t id;
};

// ---- state access policies -------------------------------------------------
struct IOF32 {
    using R = float;
    static constexpr int kMinBlocks = 16; // v1 kernel: 32 registers, full occupancy (it is latency-bound at the low densities it serves)
    const float4 *__restrict__ pt;        // sorted records in COMPUTE order (candidates and own position), type << kTypeShift
    const float2 *__restrict__ vel;       // velocities in PRE-sort order: the gather does not move them
    float4 *__restrict__ pt_out;
    float2 *__restrict__ vel_out;         // the other velocity buffer; the launcher swaps the two afterwards
    const int32_t *__restrict__ src;      // pre-sort slot of sorted particle i
    int first;                            // sorted-array index of target 0
    int stable;                           // 0: PLIFE_FLAG_UNSTABLE_SORT (results stay at the compute slot)
    StableKey key;                        // previous-array-order key (slab mode: arrivals; else the identity)
    __device__ __forceinline__ Cand<float> cand(int j) const
    {
        float4 q = __ldg(pt + j);
        return Cand<float>{q.x, q.y, __float_as_int(q.z) >> kTypeShift, __float_as_uint(q.w)};
    }
    __device__ __forceinline__ Cand<float> cand_pos(int j) const { return cand(j); } // one record: the type rides along
    __device__ __forceinline__ void cand_type(int, Cand<float> &) const {}
    __device__ __forceinline__ void self_vel(int i, float &vx, float &vy) const
    {
        float2 v = __ldg(vel + __ldg(src + i));
        vx = v.x;
        vy = v.y;
    }
    // where particle i's result goes: its slot in the reference's order (cells.cu: gather_f32)
    __device__ __forceinline__ int out_slot(int i, int cxy, const int32_t *__restrict__ cell_end, const Grid &g) const
    {
        if (!stable || g.ks == 0) return i;
        StableKey k = key; // (a copy: kernel parameters are read-only, modifying one in place would spill it to local memory)
        k.load();
        return reference_slot(i, cxy, src, cell_end, g, first, k);
    }
    __device__ __forceinline__ void store(int o, float x, float y, float vx, float vy, int type, uint32_t id) const
    {
        pt_out[o] = make_float4(x, y, __int_as_float(type), __uint_as_float(id));
        vel_out[o] = make_float2(vx, vy);
    }
};

struct IOF64 {
    using R = double;
    static constexpr int kMinBlocks = 8; // 64 registers: 8 CTAs per SM (70 registers would cost one)
    StateF64 in, out;
    __device__ __forceinline__ Cand<double> cand(int j) const
    {
        double2 p = __ldg(in.pos + j);
        return Cand<double>{p.x, p.y, __ldg(in.type + j), __ldg(in.id + j)};
    }
    __device__ __forceinline__ Cand<double> cand_pos(int j) const
    {
        double2 p = __ldg(in.pos + j);
        return Cand<double>{p.x, p.y, 0, 0u};
    }
    __device__ __forceinline__ void cand_type(int j, Cand<double> &q) const { q.type = __ldg(in.type + j); }
    __device__ __forceinline__ void self_vel(int i, double &vx, double &vy) const
    {
        double2 v = __ldg(in.vel + i);
        vx = v.x;
        vy = v.y;
    }
    __device__ __forceinline__ int out_slot(int i, int, const int32_t *, const Grid &) const { return i; } // fp64: the sorted order IS the reference's
    __device__ __forceinline__ void store(int o, double x, double y, double vx, double vy, int type, uint32_t id) const
    {
        out.pos[o] = make_double2(x, y);
        out.vel[o] = make_double2(vx, vy);
        out.type[o] = type;
        out.id[o] = id;
    }
};

// ---- small helpers -----------------------------------------------------------
// B/Physics.java:377-395 wrapContainerX/Y: a single +-n, not a modulo
__device__ __forceinline__ int wrap_container(int c, int n) { return c < 0 ? c + n : (c >= n ? c - n : c); }

__device__ __forceinline__ float rsqrt_fast(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// B/Range.java:46-57 (wrap): repeated +-1 until in [0,1).  Every intermediate
// step of that loop is exact except the one crossing zero, so the loop equals a
// single rounded `v - floor(v)`; like the loop it can return exactly 1.0 for a
// tiny negative input.
template <typename R>
__device__ __forceinline__ R range_wrap(R v)
{
    if (v < R(0) || v >= R(1)) v = v - floor(v);
    return v;
}
// B/Range.java:89-96 (clamp)
template <typename R>
__device__ __forceinline__ R range_clamp(R v)
{
    return v < R(0) ? R(0) : (v > R(1) ? R(1) : v);
}

// B/Range.java:74-81 wrapConnection on the raw difference b - a.
// fp64: literal.  fp32: the same decision, but the +-1 is applied to the operand
// that lies in [0.5, 1] first, where it is exact, so a seam pair
// is as accurate as an interior pair.
__device__ __forceinline__ double wrap_connection(double a, double b)
{
    double d = b - a;
    if (d < -0.5) return d + 1;
    else if (d >= 0.5) return d - 1;
    return d;
}
__device__ __forceinline__ float wrap_connection(float a, float b)
{
    float d = b - a;
    if (d < -0.5f) return b + (1.0f - a);
    else if (d >= 0.5f) return (b - 1.0f) - a;
    return d;
}

// in-range predicate of B/Physics.java:430-432: d2 != 0 && d2 <= rmax*rmax
__device__ __forceinline__ bool in_range(double dx, double dy, double r2)
{
    double d2 = dx * dx + dy * dy;
    return d2 != 0 && d2 <= r2;
}
constexpr float kTiny = 1e-36f; // keeps rsqrt finite for d2 == 0; far below any representable pair distance^2
__device__ __forceinline__ bool in_range(float dx, float dy, float r2)
{
    float d2 = fmaf(dx, dx, fmaf(dy, dy, kTiny));
    return (dx != 0.f || dy != 0.f) && d2 <= r2;
}

// Visitors whose in-range branch is long (the literal force: sqrt, divisions) declare kDeferred: the interior
// traversal then lets every lane scan forward to ITS next in-range candidate before the warp runs the long part
// together.  With the plain loop the long part executes whenever any lane of the warp has a hit - practically every
// iteration - with a third of the lanes active (ncu, fp64 C3: 13.7 of 32 threads active on average).  Order of the
// hits per lane is unchanged, so the accumulated result is bit-identical.
template <typename V, typename = void>
struct is_deferred : std::false_type {};
template <typename V>
struct is_deferred<V, std::void_t<decltype(V::kDeferred)>> : std::bool_constant<V::kDeferred> {};

// Two phases per lane.  Scan: every lane walks all candidates of its three row ranges (cheap, the whole warp active)
// and notes the in-range ones in a shared-memory list.  Process: the warp runs the long part over the lists, slot k of
// every lane together, so the lanes stay busy as long as their hit counts are similar (50 +- 7 at 16 particles/cell).
// A list entry is 2 bits of row + 14 bits of offset from that row's base; a lane whose list is full, or whose offset
// would overflow, makes the whole warp flush (vote), so any cell occupancy works.
constexpr int kHitSlots = 64;

template <typename IO, typename V>
__device__ __forceinline__ void traverse_listed(const IO &io, const int32_t *__restrict__ cell_end, const Grid &g, bool interior,
                                                typename IO::R xi, typename IO::R yi, int cx0, int cy0, V &v)
{
    using R = typename IO::R;
    __shared__ unsigned short s_hits[kHitSlots * kForceThreads];
    unsigned short *mine = s_hits + threadIdx.x; // slot k at mine[k * kForceThreads]
    const unsigned mask = __activemask();        // the lanes that entered: all loops below are uniform over them
    const int base0 = (cy0 + g.ly_shift - 1) * g.nx + cx0;
    int b0 = 0, b1 = 0, b2 = 0; // base of each row's offsets
    int nh = 0;
    auto process = [&]() {
        const int most = __reduce_max_sync(mask, nh);
        for (int k = 0; k < most; ++k) {
            if (k < nh) {
                const unsigned en = mine[k * kForceThreads];
                const int r = en >> 14;
                const int j = (r == 0 ? b0 : (r == 1 ? b1 : b2)) + (int)(en & 16383u);
                Cand<R> q = io.cand_pos(j);
                io.cand_type(j, q);
                v.hit(q, q.x - xi, q.y - yi);
            }
        }
        nh = 0;
    };
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        int j = 0, e = 0;
        if (interior) cell_span(cell_end, base0 + r * g.nx - 1, base0 + r * g.nx + 1, g.ks, j, e);
        int rb = j;
        if (r == 0) b0 = rb; else if (r == 1) b1 = rb; else b2 = rb;
        for (;;) {
            for (; j < e && nh < kHitSlots && j - rb < 16384; ++j) {
                const Cand<R> q = io.cand_pos(j);
                if (v.test(q.x - xi, q.y - yi)) {
                    mine[nh * kForceThreads] = (unsigned short)((r << 14) | (j - rb));
                    ++nh;
                }
            }
            if (!__any_sync(mask, j < e)) break; // every lane finished this row
            process();                           // somebody is full (or out of offset bits): all flush
            rb = j;                              // lists are empty: this row's offsets may restart here
            if (r == 0) b0 = rb; else if (r == 1) b1 = rb; else b2 = rb;
        }
    }
    process();
}

// ---- 3x3 traversal -------------------------------------------------------------
// Calls v.pair(j, cand, dx, dy) for the candidates of particle i in the
// reference's order (B/Physics.java:407-439).  Lanes whose 3x3 block needs no
// cell wrap ("interior", needs nx >= 4 so that |x_j - x_i| < 2*rmax <= 0.5 and
// wrapConnection is the identity) merge the three cells of a row into one
// contiguous index range; the candidate j == i is NOT filtered there (its
// d2 == 0 fails the in-range test).  All other lanes take the literal path.
// (fp32 handles with fine bins: the candidates of a cell are the reference's, their order inside the cell is by bin.)
template <typename IO, typename V>
__device__ __forceinline__ void traverse(const IO &io, const int32_t *__restrict__ cell_end, const Grid &g, int wrap,
                                         int i, typename IO::R xi, typename IO::R yi, int cxy, V &v)
{
    using R = typename IO::R;
    // :404-405 floor(x / containerSize) without the ==nx clamp: cached by the binning pass (cell_coords)
    const int cx0 = (cxy & 0xffff) >> g.ks;
    const int cy0 = scan_row(cxy >> 16, g);
    const bool interior = g.nx >= 4 && cx0 >= 1 && cx0 <= g.nx - 2 && cy0 >= 1 && cy0 <= g.ny - 2;
    if constexpr (is_deferred<V>::value) {
        traverse_listed(io, cell_end, g, interior, xi, yi, cx0, cy0, v);
    }
    if (interior && is_deferred<V>::value) {
    } else if (interior) {
#pragma unroll 1
        for (int oy = -1; oy <= 1; ++oy) {
            const int base = (cy0 + g.ly_shift + oy) * g.nx + cx0; // local cell of (cx0, cy0 + oy)
            int s, e;
            cell_span(cell_end, base - 1, base + 1, g.ks, s, e);
#pragma unroll 4
            for (int j = s; j < e; ++j) {
                Cand<R> q = io.cand(j);
                v.pair(j, q, q.x - xi, q.y - yi);
            }
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < 9; ++k) {
            const int ox = k % 3 - 1, oy = k / 3 - 1; // order of :88-98
            int cx = wrap_container(cx0 + ox, g.nx);  // :408
            int cy = wrap_container(cy0 + oy, g.ny);  // :409
            if (wrap) {
                cx = wrap_container(cx, g.nx); // :411
                cy = wrap_container(cy, g.ny); // :412
            } else if (cx < 0 || cx >= g.nx || cy < 0 || cy >= g.ny) {
                continue; // :414-416
            }
            const int ci = cx + local_row(cy, g) * g.nx; // :418 (local cell index)
            int s, e;                                    // :420-421
            cell_span(cell_end, ci, ci, g.ks, s, e);
            for (int j = s; j < e; ++j) {
                if (j == i) continue; // :424
                Cand<R> q = io.cand(j);
                R dx, dy;
                if (wrap) { // :464-467
                    dx = wrap_connection(xi, q.x);
                    dy = wrap_connection(yi, q.y);
                } else {
                    dx = q.x - xi;
                    dy = q.y - yi;
                }
                v.pair(j, q, dx, dy);
            }
        }
    }
}

// ---- accelerators (B/Accelerator.java:16): (a, pos/rmax) -> acceleration ------
// Literal forms, used for fp64 (bit-faithful to the reference's operation order)
// and for the builder-defined kinds in both precisions.
template <typename R, int KIND>
__device__ __forceinline__ void accelerate(R a, R px, R py, const R *prm, R &ox, R &oy)
{
    R dist = sqrt(px * px + py * py); // JOML length(), z == 0
    if (KIND == PLIFE_ACC_PARTICLE_LIFE || KIND == PLIFE_ACC_PARTICLE_LIFE_R || KIND == PLIFE_ACC_PARTICLE_LIFE_R2) {
        const R beta = prm[0];
        // A/Main.java:276-278
        // `dist < beta ? dist / beta - 1 : a * (1 - abs(1 + beta - 2 * dist) / (1 - beta))` with the operands of the one
        // division picked first: the same IEEE operations on the same values, but a warp whose lanes disagree about the
        // branch no longer executes two divisions
        const bool rep = dist < beta;
        const R q = (rep ? dist : fabs(R(1) + beta - R(2) * dist)) / (rep ? beta : R(1) - beta);
        R force = rep ? q - R(1) : a * (R(1) - q);
        R k;
        if (KIND == PLIFE_ACC_PARTICLE_LIFE) k = force / dist; // A/Main.java:279
        else if (KIND == PLIFE_ACC_PARTICLE_LIFE_R) k = force / (dist * dist);
        else k = force / (dist * dist * dist);
        ox = px * k;
        oy = py * k;
    } else if (KIND == PLIFE_ACC_ROTATOR_90) {
        R force = a * (R(1) - dist);
        R k = force / dist;
        ox = -py * k;
        oy = px * k;
    } else if (KIND == PLIFE_ACC_ROTATOR_ATTR) {
        R force = R(1) - dist;
        R angle = -a * R(3.14159265358979323846);
        R c = cos(angle), sn = sin(angle);
        R k = force / dist;
        ox = (c * px + sn * py) * k;
        oy = (-sn * px + c * py) * k;
    } else { // PLIFE_ACC_PLANETS
        R r = dist > R(0.01) ? dist : R(0.01);
        R k = R(0.01) / (r * r * r);
        ox = px * k;
        oy = py * k;
    }
}

// ---- visitors ---------------------------------------------------------------------
constexpr int kMatGlobal = 0; // Mt in global memory (any m)
constexpr int kMatShared = 1; // Mt copied to shared memory (m <= 64)
constexpr int kMatLaneTab = 2; // per-lane copy of the lane's own matrix row (staged kernel, m <= 32)

template <typename R, int MODE>
struct MatrixView { // transposed matrix Mt[other][own]
    const R *g;     // global copy
    const R *s;     // shared copy (MODE 1: Mt, MODE 2: tab[other][thread]), already multiplied by `scale`
    int m, own;
    R scale;
    uint32_t row; // shared-window byte address of this lane's column
    uint32_t mb;  // bytes per matrix row
    __device__ __forceinline__ void init()
    {
        if (MODE == kMatShared) {
            row = (uint32_t)__cvta_generic_to_shared(s + own);
            mb = (uint32_t)m * (uint32_t)sizeof(R);
        } else if (MODE == kMatLaneTab) {
            row = (uint32_t)__cvta_generic_to_shared(s + threadIdx.x);
        }
    }
    // `key` is the candidate's type (MODE 0/1) or type << kTypeShift (MODE 2)
    __device__ __forceinline__ R get(int key) const
    {
        if (MODE == kMatShared) return lds(row + (uint32_t)key * mb); // one IMAD + one LDS
        if (MODE == kMatLaneTab) return lds(row + (uint32_t)key);     // one IADD + one conflict-free LDS
        return __ldg(g + key * m + own) * scale;
    }
    static __device__ __forceinline__ float lds_(uint32_t a, float)
    {
        float v;
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ double lds_(uint32_t a, double)
    {
        double v;
        asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ R lds(uint32_t a) { return lds_(a, R(0)); }
};

// Literal visitor: accumulates term by term into the running velocity, exactly
// like `p.velocity.add(deltaV.mul(rmax*force*dt))` (B/Physics.java:437).
template <typename R, int KIND, int MODE>
struct LiteralForce {
    R vx, vy;
    R r2, invr, k2;
    const R *prm;
    MatrixView<R, MODE> M;
    // fp64 only: measured on the fp32 literal kinds (C5, staged candidates) a hit list made the step 10-40 % slower -
    // there the in-range branch is some 40 instructions, not the 150 of IEEE double sqrt and division
    static constexpr bool kDeferred = sizeof(R) == 8;
    __device__ __forceinline__ bool test(R dx, R dy) const { return in_range(dx, dy, r2); }
    __device__ __forceinline__ void hit(const Cand<R> &q, R dx, R dy)
    {
        R px = dx * invr, py = dy * invr; // :434 (JOML div = mul by reciprocal)
        R ox, oy;
        accelerate<R, KIND>(M.get(q.type), px, py, prm, ox, oy); // :435
        vx = vx + ox * k2;                                       // :437
        vy = vy + oy * k2;
    }
    __device__ __forceinline__ void pair(int, const Cand<R> &q, R dx, R dy)
    {
        if (test(dx, dy)) hit(q, dx, dy);
    }
    __device__ __forceinline__ void pair_rel(const float4 &q, float2 nself)
    {
        pair(-1, Cand<R>{q.x, q.y, __float_as_int(q.z), 0u}, q.x + nself.x, q.y + nself.y);
    }
    __device__ __forceinline__ void quad(const float4 &X, const float4 &Y, const float4 &Kq, float2 nself)
    {
        pair_rel(make_float4(X.x, Y.x, Kq.x, 0.f), nself);
        pair_rel(make_float4(X.y, Y.y, Kq.y, 0.f), nself);
        pair_rel(make_float4(X.z, Y.z, Kq.z, 0.f), nself);
        pair_rel(make_float4(X.w, Y.w, Kq.w, 0.f), nself);
    }
    __device__ __forceinline__ void finish(R &ovx, R &ovy) const
    {
        ovx = vx;
        ovy = vy;
    }
};

// ---- packed FP32 (sm_100a FADD2 / FFMA2: two fp32 lanes per instruction, each an ordinary IEEE fp32 add / fma) ----
// The force pass is bound by the SM's issue rate (one warp instruction per cycle and scheduler), not by the FP32
// pipe: a packed instruction does the work of two in ONE issue slot.
__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}

// fp32 fast visitor for kind 0 (A/Main.java:275-280), branch-free.  With b = beta*rmax (absolute distance units) the
// reference's f(d/rmax) is
//     f = (1/b) * [ min(d - b, 0) + a' * max(min(d - b, rmax - d), 0) ],  a' = a * 2*beta/(1-beta)
// exactly: the first term is the repulsion (d < b), the second the triangular lobe on [b, rmax] (rising side d - b up to
// its peak at (b + rmax)/2, falling side rmax - d), and both vanish beyond rmax, so the `d2 <= rmax^2` test
// (B/Physics.java:432) is implied (f is continuous and 0 at the cutoff).  The accumulated quantity is f*b/d (the
// acceleration is pos*(f/d), A/Main.java:279), computed straight from u = 1/d = rsqrt(d2):
//     r = 1 - b*u,  p = 1 - rmax*u,  f*b/d = min(r, 0) + a' * max(min(-p, r), 0)
// which needs neither d nor a final multiplication by 1/d.  The factor 1/b is folded into the final scale (fast_k), a'
// into the shared-memory matrix.  {r, p} is ONE packed fma, {ax, ay} += g * {dx, dy} another, and the staged loops form
// {dx, dy} with one packed add (the record's x and y sit in an aligned register pair after the LDS.128):
// 13.75 issued instructions per pair evaluation instead of 16.75.
template <int MODE>
struct FastParticleLife32 {
    float2 acc; // {ax, ay}
    float2 c1;  // {-b, -rmax}
    MatrixView<float, MODE> M;
    __device__ __forceinline__ FastParticleLife32(const ForceParams<float> &P, const MatrixView<float, MODE> &m)
        : acc(make_float2(0.f, 0.f)), c1(make_float2(-P.fast_b, -P.rmax)), M(m)
    {
        // keep the pair in registers: without this ptxas re-reads both halves from the constant bank in every loop trip
        asm volatile("" : "+f"(c1.x), "+f"(c1.y));
    }
    __device__ __forceinline__ FastParticleLife32(float2 acc_, float2 c1_, uint32_t row) : acc(acc_), c1(c1_)
    {
        M.g = nullptr;
        M.s = nullptr;
        M.m = 0;
        M.own = 0;
        M.scale = 1.0f;
        M.row = row;
        M.mb = 0u;
    }
    __device__ __forceinline__ float ax() const { return acc.x; }
    __device__ __forceinline__ float ay() const { return acc.y; }
    __device__ __forceinline__ void core(int key, float2 d)
    {
        const float d2 = fmaf(d.x, d.x, fmaf(d.y, d.y, kTiny));
        const float u = rsqrt_fast(d2);
        const float a = M.get(key);
        const float2 rp = ffma2(c1, make_float2(u, u), make_float2(1.0f, 1.0f)); // (the addend is an immediate)
        const float rep = fminf(rp.x, 0.0f);
        const float att = fmaxf(fminf(-rp.y, rp.x), 0.0f);
        const float g = fmaf(a, att, rep); // self / coincident: dx = dy = 0 kills the term
        acc = ffma2(make_float2(g, g), d, acc);
    }
    __device__ __forceinline__ void pair(int, const Cand<float> &q, float dx, float dy) { core(q.type, make_float2(dx, dy)); }
    // four candidates {x0..x3}, {y0..y3}, {key0..key3} (force_kernel_staged4): the same operations per pair as core(), in
    // candidate order, two candidates per packed instruction where the operands sit in register pairs
    __device__ __forceinline__ void quad(const float4 &X, const float4 &Y, const float4 &Kq, float2 nself)
    {
        const float2 nx2 = make_float2(nself.x, nself.x), ny2 = make_float2(nself.y, nself.y);
        const float2 tiny2 = make_float2(kTiny, kTiny), one2 = make_float2(1.0f, 1.0f);
        const float2 dxa = fadd2(make_float2(X.x, X.y), nx2), dxb = fadd2(make_float2(X.z, X.w), nx2);
        const float2 dya = fadd2(make_float2(Y.x, Y.y), ny2), dyb = fadd2(make_float2(Y.z, Y.w), ny2);
        const float2 sa = ffma2(dxa, dxa, ffma2(dya, dya, tiny2)), sb = ffma2(dxb, dxb, ffma2(dyb, dyb, tiny2));
        const float2 ua = make_float2(rsqrt_fast(sa.x), rsqrt_fast(sa.y)), ub = make_float2(rsqrt_fast(sb.x), rsqrt_fast(sb.y));
        const float2 aa = make_float2(M.get(__float_as_int(Kq.x)), M.get(__float_as_int(Kq.y)));
        const float2 ab = make_float2(M.get(__float_as_int(Kq.z)), M.get(__float_as_int(Kq.w)));
        const float2 cb = make_float2(c1.x, c1.x), cr = make_float2(c1.y, c1.y);
        const float2 ra = ffma2(cb, ua, one2), rb = ffma2(cb, ub, one2); // r = 1 - b*u
        const float2 pa = ffma2(cr, ua, one2), pb = ffma2(cr, ub, one2); // p = 1 - rmax*u
        const float2 ga = ffma2(aa, make_float2(fmaxf(fminf(-pa.x, ra.x), 0.0f), fmaxf(fminf(-pa.y, ra.y), 0.0f)),
                                make_float2(fminf(ra.x, 0.0f), fminf(ra.y, 0.0f)));
        const float2 gb = ffma2(ab, make_float2(fmaxf(fminf(-pb.x, rb.x), 0.0f), fmaxf(fminf(-pb.y, rb.y), 0.0f)),
                                make_float2(fminf(rb.x, 0.0f), fminf(rb.y, 0.0f)));
        acc.x = fmaf(ga.x, dxa.x, acc.x);
        acc.y = fmaf(ga.x, dya.x, acc.y);
        acc.x = fmaf(ga.y, dxa.y, acc.x);
        acc.y = fmaf(ga.y, dya.y, acc.y);
        acc.x = fmaf(gb.x, dxb.x, acc.x);
        acc.y = fmaf(gb.x, dyb.x, acc.y);
        acc.x = fmaf(gb.y, dxb.y, acc.x);
        acc.y = fmaf(gb.y, dyb.y, acc.y);
    }
};

struct NeighborDiag {
    int count;
    unsigned long long hash;
    template <typename R>
    __device__ __forceinline__ void add(const Cand<R> &q)
    {
        unsigned long long z = (unsigned long long)q.id + 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        hash += z ^ (z >> 31);
        count++;
    }
};
template <typename R>
struct NeighborVisitor {
    NeighborDiag d;
    R r2;
    __device__ __forceinline__ void pair(int, const Cand<R> &q, R dx, R dy)
    {
        if (in_range(dx, dy, r2)) d.add(q);
    }
};
template <typename R>
struct PairCountVisitor {
    int self;
    unsigned long long count;
    __device__ __forceinline__ void pair(int j, const Cand<R> &, R, R) { count += (j != self) ? 1u : 0u; }
};

// Fused K_BIN of the NEXT step: the integrate epilogue knows the new position, so it bins it
// (B/Physics.java:329-332 of the next update()) and saves a pass over the particle array.
// Slab mode: a particle whose new row belongs to another rank is appended to the migration
// message of that direction (two 16-byte records: {x,y,type,id}, {vx,vy,source slot,-}) and its
// slot is marked dead; record 0 of a message is the header {count, far, -, -}.
struct NextBin {
    int32_t *cell;  // bin word per particle, or nullptr: do not bin
    int32_t *count; // per-bin histogram (zeroed by K_SCAN of this step)
    float4 *mig[2]; // migration messages [down, up] (slab mode) or nullptr
    int mig_cap;
    int *late_err;  // slab mode, interior launch: a particle that leaves the slab raises this error bit instead (see slab.cu)
    template <typename R>
    __device__ __forceinline__ void add(int i, R x, R y, R vx, R vy, int type, uint32_t id, const Grid &g) const
    {
        if (!cell) return;
        const int cxy = cell_coords_fast((double)x, (double)y, g);
        const int c = container_of(cxy, g);
        if (c >= 0) {
            cell[i] = cxy;
            if (count) atomicAdd(count + c, 1); // (small mode: small_sort rebuilds the histogram from the bin words)
            return;
        }
        cell[i] = -1; // leaves this slab
        if (late_err) {
            atomicOr(late_err, 4); // kErrFar: the migration exchange of this step is already under way
            return;
        }
        if (!mig[0]) return;
        int cy = cxy >> 16;
        if (cy == g.ny) cy = g.ny - 1;
        int du = cy - g.row_hi; // rows above the slab (ring distance)
        if (du < 0) du += g.ny;
        int dd = g.row_lo - 1 - cy; // rows below the slab
        if (dd < 0) dd += g.ny;
        const int dir = du <= dd ? 1 : 0;
        const bool far = dir ? du >= g.rows_up : dd >= g.rows_dn;
        float4 *msg = dir ? mig[1] : mig[0]; // (a select, not an indexed load: the struct stays in the parameter bank)
        int *hdr = reinterpret_cast<int *>(msg);
        if (far) {
            atomicAdd(hdr + 1, 1); // more than one slab in one step: reported as PLIFE_ERR_STATE
            return;
        }
        const int k = atomicAdd(hdr, 1);
        if (k < mig_cap) {
            msg[1 + 2 * k] = make_float4((float)x, (float)y, __int_as_float(type), __uint_as_float(id));
            msg[2 + 2 * k] = make_float4((float)vx, (float)vy, __int_as_float(i), 0.f);
        }
    }
};

// The targets of a CTA.  Plain launch: CTA b owns sorted particles [128 b, 128 b + 128) below n.  Slab mode
// splits the pass in two launches over device-resident index ranges (P.tr = {s0, e0, s1, e1}): the interior rows
// while the halo is in flight, then the first and the last owned row.
template <typename R>
__device__ __forceinline__ bool cta_targets(const ForceParams<R> &P, int &i0, int &lim)
{
    const int b = blockIdx.x;
    if (P.tr) {
        const int s0 = P.tr[0], e0 = P.tr[1];
        const int nb0 = (max(e0 - s0, 0) + kForceThreads - 1) / kForceThreads;
        if (b < nb0) {
            i0 = s0 + b * kForceThreads;
            lim = e0;
        } else {
            i0 = P.tr[2] + (b - nb0) * kForceThreads;
            lim = P.tr[3];
        }
    } else {
        i0 = b * kForceThreads;
        lim = P.n_dev ? *P.n_dev : P.n;
    }
    return i0 < lim;
}

// ---- kernels ---------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void load_matrix_smem(R *sM, const R *__restrict__ gMt, int m, R scale)
{
    for (int k = threadIdx.x; k < m * m; k += blockDim.x) sM[k] = gMt[k] * scale;
    __syncthreads();
}

template <typename IO, int KIND, bool SMEM, bool FAST>
__global__ void __launch_bounds__(kForceThreads, IO::kMinBlocks) force_kernel(IO io, const int32_t *__restrict__ cell_end,
                                                             const int32_t *__restrict__ cell_sorted,
                                                             ForceParams<typename IO::R> P,
                                                             const typename IO::R *__restrict__ gMt, NextBin nb)
{
    using R = typename IO::R;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    R *sM = reinterpret_cast<R *>(smem_raw);
    int i0, lim;
    if (!cta_targets(P, i0, lim)) return;
    const R mscale = FAST ? P.fast_a_scale : R(1);
    if (SMEM) load_matrix_smem(sM, gMt, P.m, mscale);

    const int i = i0 + threadIdx.x;
    if (i >= lim) return;
    const int si = P.first + i;
    const Cand<R> self = io.cand(si);
    const int cxy = __ldg(cell_sorted + i);
    R vx, vy;
    io.self_vel(i, vx, vy);
    constexpr int MODE = SMEM ? kMatShared : kMatGlobal;
    MatrixView<R, MODE> M{gMt, sM, P.m, self.type, mscale, 0u, 0u};
    M.init();

    R nvx, nvy;
    if constexpr (FAST) {
        FastParticleLife32<MODE> v(P, M);
        traverse(io, cell_end, P.g, P.wrap, si, self.x, self.y, cxy, v);
        nvx = fmaf(P.fast_k, v.ax(), vx * P.mu); // friction first (:401-402), then the summed acceleration
        nvy = fmaf(P.fast_k, v.ay(), vy * P.mu);
    } else {
        LiteralForce<R, KIND, MODE> v{vx * P.mu, vy * P.mu, P.r2, P.invr, P.k2, P.accp, M};
        traverse(io, cell_end, P.g, P.wrap, si, self.x, self.y, cxy, v);
        v.finish(nvx, nvy);
    }
    // updatePosition (:443-450): pos = vel*dt + pos, then wrap or clamp (:499-505)
    R nx_ = nvx * P.dt + self.x;
    R ny_ = nvy * P.dt + self.y;
    if (P.wrap) {
        nx_ = range_wrap(nx_);
        ny_ = range_wrap(ny_);
    } else {
        nx_ = range_clamp(nx_);
        ny_ = range_clamp(ny_);
    }
    // this kernel only ever runs on the plain cell list (ks == 0: fp64, m > 32, fewer than 4 particles per cell - the launcher
    // refuses anything else), where the compute order is the reference's order
    const int o = i;
    io.store(o, nx_, ny_, nvx, nvy, self.type, self.id);
    nb.add(o, nx_, ny_, nvx, nvy, self.type, self.id, P.g);
}

// The v2 kernels below exist ONLY in the fp32 translation unit.  force_f64.cu is compiled with
// -fmad=false; if it also instantiated them, the library would hold two device images of the same
// kernel (fused and unfused arithmetic) and which one a launch binds to is decided at module-load time:
// results then differ by an ulp from process to process (found the hard way, profiles/r1_multigpu.md).
#ifdef PLIFE_FORCE_F32_TU
// ---- shared-memory staged candidates (fp32) ----------------------------------------------------------------------------
// The per-lane `LDG.128` of the kernel above costs 4 L1 data-pipe cycles per warp whatever the address pattern, and the
// shared matrix lookup bank-conflicts between lanes of different cells (profiles/r1_force_kernel.md); together they
// saturate L1TEX before the FP32 pipes.  The staged kernel below copies the three index ranges that cover the rows
// above / of / below a CTA's targets into shared memory - bin ids are row-major, so [binA + dy*nxk - K, binB + dy*nxk + K]
// is ONE contiguous index range per dy - and every lane copies its own matrix row into a [type][thread] table (bank = lane:
// conflict-free lookups).
//
// Fine bins: a lane walks the bins [own - K, own + K] of each row: every particle within rmax in x lies there
// (bin = floor(K x / rmax) is monotone in x), and a row needs (2 + 1/K) cell widths of candidates instead of 3.
// Lanes on the domain seam walk global memory with the literal 9-cell loop; CTAs whose ranges exceed the staging
// capacity (dense clusters) stream them through it in chunks.
constexpr int kTabMaxM = 32;
static_assert(kForceThreads * 4 == (1 << kTypeShift), "lane-table row stride");

__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 v;
    // volatile + memory clobber: the staging area is rewritten between barriers (chunked walks), and an asm the compiler
    // believes to be a pure function of its address may be moved across them
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}

// The literal 9-cell walk over global memory (seam lanes): B/Physics.java:412-437.  Records carry shifted types.
// (Kept as the plain one-candidate-at-a-time loop: batching its loads, inline or as a real function call, costs the staged
// main loop registers or spills and made the 16M-particle step 2 % slower for a 10 % gain at 10 000 particles.)
template <typename V>
__device__ __forceinline__ void traverse_global32(const IOF32 &io, const int32_t *__restrict__ cell_end, const Grid &g, int wrap,
                                                  int i, float xi, float yi, int cx0, int cy0, V &v)
{
#pragma unroll 1
    for (int k = 0; k < 9; ++k) {
        const int ox = k % 3 - 1, oy = k / 3 - 1;
        int cx = wrap_container(cx0 + ox, g.nx);
        int cy = wrap_container(cy0 + oy, g.ny);
        if (wrap) {
            cx = wrap_container(cx, g.nx);
            cy = wrap_container(cy, g.ny);
        } else if (cx < 0 || cx >= g.nx || cy < 0 || cy >= g.ny) {
            continue;
        }
        const int ci = cx + local_row(cy, g) * g.nx;
        int s, e;
        cell_span(cell_end, ci, ci, g.ks, s, e);
        for (int j = s; j < e; ++j) {
            if (j == i) continue;
            const float4 r = __ldg(io.pt + j);
            const Cand<float> q{r.x, r.y, __float_as_int(r.z), 0u}; // key = type << kTypeShift
            float dx, dy;
            if (wrap) {
                dx = wrap_connection(xi, q.x);
                dy = wrap_connection(yi, q.y);
            } else {
                dx = q.x - xi;
                dy = q.y - yi;
            }
            v.pair(j, q, dx, dy);
        }
    }
}

// the same walk handing the visitor PLAIN types (matrix in global / shared memory instead of the per-lane table)
template <typename V>
__device__ __forceinline__ void traverse_global32_plain(const IOF32 &io, const int32_t *__restrict__ cell_end, const Grid &g, int wrap,
                                                        int i, float xi, float yi, int cx0, int cy0, V &v)
{
#pragma unroll 1
    for (int k = 0; k < 9; ++k) {
        const int ox = k % 3 - 1, oy = k / 3 - 1;
        int cx = wrap_container(cx0 + ox, g.nx);
        int cy = wrap_container(cy0 + oy, g.ny);
        if (wrap) {
            cx = wrap_container(cx, g.nx);
            cy = wrap_container(cy, g.ny);
        } else if (cx < 0 || cx >= g.nx || cy < 0 || cy >= g.ny) {
            continue;
        }
        const int ci = cx + local_row(cy, g) * g.nx;
        int s, e;
        cell_span(cell_end, ci, ci, g.ks, s, e);
        for (int j = s; j < e; ++j) {
            if (j == i) continue;
            const Cand<float> q = io.cand(j);
            const float dx = wrap ? wrap_connection(xi, q.x) : q.x - xi;
            const float dy = wrap ? wrap_connection(yi, q.y) : q.y - yi;
            v.pair(j, q, dx, dy);
        }
    }
}

// ---- the staged kernel: candidates four wide -------------------------------------------------------------------------------
// At 16 particles per cell the pass is bound by the SM's ONE shared-memory data pipe (a 128-byte wavefront per cycle for all
// four schedulers) and by its issue slots, not by HBM.  The predecessor of this kernel staged plain 16-byte records (bulk
// copies) and read one per LDS.128: with fine bins the lanes of a warp then fetch up to 16 different records per load -
// 4.5 wavefronts per candidate plus one for the matrix lookup, the pipe 95 % busy (ncu: l1tex__data_pipe_lsu_wavefronts),
// the issue slots 71 % (profiles/r2_force_kernel.md).  Here the staging area holds the candidates in groups of four,
// {x0..x3 | y0..y3 | key0..key3} (48 bytes), aligned to the global sorted index, and a lane fetches a group with three
// LDS.128: 3.3 wavefronts per candidate (an LDS.128 costs the four quarter-warp phases whatever the addresses; the group
// stride of 3 x 16 bytes is coprime to the 8 bank groups, so nothing is added: 1 % excessive wavefronts).  dx, dy, d2,
// {r, p} and the weight are formed for two candidates per packed instruction (FADD2 / FFMA2): 52 instructions per group.
// The price: the ranges are not bulk-copied any more (the records are transposed on the way in: LDG.128 -> 3 x STS.32,
// conflict-free), and a lane's walk covers whole groups - the records next to its range come along (see walk_row4).
#ifndef PLIFE_STAGED4_MIN_BLOCKS
#define PLIFE_STAGED4_MIN_BLOCKS 10 // 48 registers: four candidates in flight per lane
#endif
constexpr int kGroupBytes = 48;
constexpr int kStageLead = 4; // room for the groups' alignment below a range start
constexpr int kStageTail = 4; // sentinels after a range: the unmasked walk reads its last group in full
constexpr float kFar = 1.0e9f; // x of a masked / sentinel candidate

// smem byte offset of candidate slot k of a row (slots are group-aligned: 4 per 48-byte group)
__device__ __forceinline__ uint32_t slot_x(uint32_t k) { return (k >> 2) * (uint32_t)kGroupBytes + (k & 3u) * 4u; }

// slots [0, n_slots) of the staging area at `base` <- sorted records [j0, j0 + n_slots), transposed into groups
// (LDG.128 -> 3 x STS.32, conflict-free); slots outside [lo, hi) become sentinels (far away, matrix key 0)
__device__ __forceinline__ void stage_in4(unsigned char *base, const float4 *__restrict__ pt, int j0, int n_slots, int lo, int hi)
{
    // thread t owns the slots t, t + 128, ...: slot k lives at (k / 4) * 48 + (k % 4) * 4, i.e. a fixed 32 groups further per trip
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(base) + slot_x(threadIdx.x);
    const float4 *src = pt + j0 + threadIdx.x;
    const uint32_t len = (uint32_t)(hi - lo);
    uint32_t rel = (uint32_t)(j0 + (int)threadIdx.x - lo); // record index relative to the range: in range <=> rel < len (unsigned)
    for (int k = threadIdx.x; k < n_slots; k += kForceThreads) {
        float4 q = make_float4(kFar, kFar, 0.f, 0.f);
        if (rel < len) q = __ldg(src);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst), "f"(q.x) : "memory");
        asm volatile("st.shared.f32 [%0+16], %1;" ::"r"(dst), "f"(q.y) : "memory");
        asm volatile("st.shared.f32 [%0+32], %1;" ::"r"(dst), "f"(q.z) : "memory");
        dst += (kForceThreads / 4) * (uint32_t)kGroupBytes;
        src += kForceThreads;
        rel += kForceThreads;
    }
}

// Candidates [s_rel, e_rel) (slot numbers) of the staging row at byte address rb, in order.  The first and the last
// group are read in full but the slots outside the range are masked: the records next to a range are NOT always far away
// in x - where the rest of a grid row is empty they belong to the adjacent grid row, i.e. to another of this lane's ranges.
template <typename V>
__device__ __forceinline__ void walk_row4(uint32_t rb, int s_rel, int e_rel, float2 nself, V &v)
{
    if (s_rel >= e_rel) return;
    uint32_t a = rb + (uint32_t)(s_rel >> 2) * kGroupBytes;
    const uint32_t al = rb + (uint32_t)((e_rel - 1) >> 2) * kGroupBytes; // the last group
    const int lead = s_rel & 3, cnt = e_rel & 3; // cnt == 0: the last group is full
    float4 X = lds128(a);
    if (lead > 0) X.x = kFar;
    if (lead > 1) X.y = kFar;
    if (lead > 2) X.z = kFar;
    if (a != al) {
        v.quad(X, lds128(a + 16u), lds128(a + 32u), nself);
#pragma unroll 1
        for (a += (uint32_t)kGroupBytes; a < al; a += (uint32_t)kGroupBytes) v.quad(lds128(a), lds128(a + 16u), lds128(a + 32u), nself);
        X = lds128(al);
    }
    if (cnt != 0) {
        X.w = kFar;
        if (cnt <= 2) X.z = kFar;
        if (cnt <= 1) X.y = kFar;
    }
    v.quad(X, lds128(al + 16u), lds128(al + 32u), nself);
}

// The masked walk of the fast visitor as a real function call: it serves the few CTAs that straddle a grid row (and the
// chunked walks of dense clusters), and as a call it costs the common path neither registers nor code.
template <int MODE>
__device__ __noinline__ float2 walk_row4_call(uint32_t rb, int s_rel, int e_rel, float2 nself, float2 acc, float2 c1, uint32_t row)
{
    FastParticleLife32<MODE> v(acc, c1, row);
    walk_row4(rb, s_rel, e_rel, nself, v);
    return v.acc;
}

template <int KIND, bool FAST>
__global__ void __launch_bounds__(kForceThreads, PLIFE_STAGED4_MIN_BLOCKS)
    force_kernel_staged4(IOF32 io, const int32_t *__restrict__ cell_end, const int32_t *__restrict__ cell_sorted, ForceParams<float> P,
                         const float *__restrict__ gM, int cap, NextBin nb)
{
    // cap: candidate slots per row (multiple of 4), lead and tail included
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tab = reinterpret_cast<float *>(smem_raw + (size_t)3 * cap * 12); // [m][kForceThreads]
    __shared__ int s_start[3], s_len[3], s_clean;

    int i0, lim;
    if (!cta_targets(P, i0, lim)) return;
    const int tid = threadIdx.x;
    const int i = i0 + tid;
    const bool valid = i < lim;
    const Grid g = P.g;
    const int K = 1 << g.ks, nxk = g.nxk();
    const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t row_bytes = (uint32_t)cap * 12u;

    if (tid == 0) { // the three row ranges of this CTA's targets
        const int b0 = container_of(__ldg(cell_sorted + i0), g);
        const int b1 = container_of(__ldg(cell_sorted + min(i0 + kForceThreads, lim) - 1), g);
        // clean: the targets share ONE grid row and the K bins of margin on either side stay inside it.  Every record next
        // to a lane's range then sits left or right of the lane's window in the same grid row, beyond rmax in x (or is a
        // sentinel past the staged range): the walk needs no masks, its extra candidates add exact zeros.
        const int x0 = b0 % nxk;
        s_clean = (x0 >= K && x0 + (b1 - b0) + K < nxk) ? 1 : 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int lo = max(b0 + (r - 1) * nxk - K, P.bin_lo);
            const int hi = min(b1 + (r - 1) * nxk + K, P.bin_hi);
            int st = 0, ln = 0;
            if (lo <= hi) {
                st = __ldg(cell_end + lo - 1);
                ln = __ldg(cell_end + hi) - st;
            }
            s_start[r] = st;
            s_len[r] = ln;
        }
    }
    Cand<float> self{0.f, 0.f, 0, 0u};
    int cxy = 0;
    const int si = P.first + i;
    float vx = 0.f, vy = 0.f;
    if (valid) {
        self = io.cand(si);
        cxy = __ldg(cell_sorted + i);
        io.self_vel(i, vx, vy);
    }
    // the six bounds of this lane's candidate ranges depend on the bin word only: fetched now, in the shadow of the staging
    const int cx0 = (cxy & 0xffff) >> g.ks, cy0 = scan_row(cxy >> 16, g);
    const bool interior = valid && g.nx >= 4 && cx0 >= 1 && cx0 <= g.nx - 2 && cy0 >= 1 && cy0 <= g.ny - 2;
    const int fb = (cxy & 0xffff) + (cy0 + g.ly_shift) * nxk; // own bin (interior lanes: no clamp needed)
    int s[3] = {0, 0, 0}, e[3] = {0, 0, 0};
    if (interior) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int base = fb + (r - 1) * nxk;
            s[r] = __ldg(cell_end + base - K - 1);
            e[r] = __ldg(cell_end + base + K);
        }
    }
    { // this lane's matrix row -> tab[other][tid] (bank = lane: conflict-free lookups)
        const float mscale = FAST ? P.fast_a_scale : 1.0f;
        const float *rowp = gM + self.type * P.m;
        if ((P.m & 3) == 0) {
            for (int t = 0; t < P.m; t += 4) {
                const float4 r4 = __ldg(reinterpret_cast<const float4 *>(rowp + t));
                tab[(t + 0) * kForceThreads + tid] = r4.x * mscale;
                tab[(t + 1) * kForceThreads + tid] = r4.y * mscale;
                tab[(t + 2) * kForceThreads + tid] = r4.z * mscale;
                tab[(t + 3) * kForceThreads + tid] = r4.w * mscale;
            }
        } else {
            for (int t = 0; t < P.m; ++t) tab[t * kForceThreads + tid] = __ldg(rowp + t) * mscale;
        }
    }
    __syncthreads(); // s_start / s_len / s_clean
    const int room = cap - kStageLead - kStageTail;
    const bool staged_ok = s_len[0] <= room && s_len[1] <= room && s_len[2] <= room;
    if (staged_ok) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int j0 = s_start[r] & ~3; // group boundaries follow the global sorted index
            stage_in4(smem_raw + (size_t)r * row_bytes, io.pt, j0, ((s_start[r] - j0 + s_len[r] + 3) & ~3) + kStageTail, s_start[r],
                      s_start[r] + s_len[r]);
        }
    }
    // the slot of the result in the reference's order is a rank over the pre-sort indices of the cell: global loads that need
    // nothing from the force pass, issued while the staging stores drain
    int o = 0;
    if (valid) o = io.out_slot(i, cxy, cell_end, g);
    if (staged_ok) {
        __syncthreads();
        if (!valid) return; // (in chunked mode every thread is needed at the barriers)
    }

    MatrixView<float, kMatLaneTab> M{nullptr, tab, P.m, self.type, 1.0f, 0u, 0u};
    M.init();
    const float2 nself = make_float2(-self.x, -self.y);
    float nvx, nvy;
    // walk(v, masked_row): `masked_row(rb, s_rel, e_rel)` is how this visitor walks one row with masks
    auto walk = [&](auto &v, auto masked_row) {
        if (staged_ok && s_clean) {
            if (interior) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int j0 = s_start[r] & ~3;
                    uint32_t a = stage_addr + (uint32_t)r * row_bytes + (uint32_t)((s[r] - j0) >> 2) * kGroupBytes;
                    const uint32_t a1 = stage_addr + (uint32_t)r * row_bytes + (uint32_t)((e[r] - j0 + 3) >> 2) * kGroupBytes;
#pragma unroll 1
                    for (; a < a1; a += (uint32_t)kGroupBytes) v.quad(lds128(a), lds128(a + 16u), lds128(a + 32u), nself);
                }
            }
        } else if (staged_ok) {
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
                const int j0 = s_start[r] & ~3;
                const int sr = r == 0 ? s[0] : (r == 1 ? s[1] : s[2]), er = r == 0 ? e[0] : (r == 1 ? e[1] : e[2]);
                masked_row(stage_addr + (uint32_t)r * row_bytes, sr - j0, er - j0);
            }
        } else { // dense cluster: the row ranges stream through the whole staging area in chunks (CTA-uniform loops)
            const int chunk = 3 * cap - kStageTail; // multiple of 4
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
                const int row_lo = s_start[r], row_hi = row_lo + s_len[r];
                const int sr = r == 0 ? s[0] : (r == 1 ? s[1] : s[2]), er = r == 0 ? e[0] : (r == 1 ? e[1] : e[2]);
#pragma unroll 1
                for (int c0 = row_lo & ~3; c0 < row_hi; c0 += chunk) {
                    const int c1 = min(c0 + chunk, row_hi); // this chunk's records: [max(c0, row_lo), c1)
                    __syncthreads();                        // the previous chunk has been consumed
                    stage_in4(smem_raw, io.pt, c0, (c1 - c0 + 3) & ~3, row_lo, c1);
                    __syncthreads();
                    masked_row(stage_addr, max(sr, c0) - c0, min(er, c1) - c0);
                }
            }
        }
        if (valid && !interior) traverse_global32(io, cell_end, g, P.wrap, si, self.x, self.y, cx0, cy0, v);
    };
    if constexpr (FAST) {
        FastParticleLife32<kMatLaneTab> v(P, M);
        { // {-b, -rmax} through a volatile shared-memory round trip of the thread's own slot: ptxas cannot rematerialise them
          // from the constant bank inside the loop any more (two of 54 instructions per group)
            __shared__ float2 s_c1[kForceThreads];
            s_c1[tid] = v.c1;
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(&s_c1[tid]);
            asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.c1.x), "=f"(v.c1.y) : "r"(a) : "memory");
        }
        walk(v, [&](uint32_t rb, int s_rel, int e_rel) { v.acc = walk_row4_call<kMatLaneTab>(rb, s_rel, e_rel, nself, v.acc, v.c1, M.row); });
        nvx = fmaf(P.fast_k, v.ax(), vx * P.mu);
        nvy = fmaf(P.fast_k, v.ay(), vy * P.mu);
    } else {
        LiteralForce<float, KIND, kMatLaneTab> v{vx * P.mu, vy * P.mu, P.r2, P.invr, P.k2, P.accp, M};
        walk(v, [&](uint32_t rb, int s_rel, int e_rel) { walk_row4(rb, s_rel, e_rel, nself, v); });
        v.finish(nvx, nvy);
    }
    if (!valid) return;
    float nx_ = fmaf(nvx, P.dt, self.x);
    float ny_ = fmaf(nvy, P.dt, self.y);
    if (P.wrap) {
        nx_ = range_wrap(nx_);
        ny_ = range_wrap(ny_);
    } else {
        nx_ = range_clamp(nx_);
        ny_ = range_clamp(ny_);
    }
    io.store(o, nx_, ny_, nvx, nvy, self.type, self.id);
    nb.add(o, nx_, ny_, nvx, nvy, self.type, self.id, g);
}

inline cudaError_t dispatch_force_staged(const IOF32 &io, const int32_t *cell_end, const int32_t *cell_sorted,
                                         const ForceParams<float> &P, int nblocks, const float *gM, int kind, int cap, NextBin nbin,
                                         cudaStream_t stream)
{
    if (nblocks <= 0) return cudaSuccess;
    const int cap4 = ((cap + 3) & ~3) + kStageLead + kStageTail; // candidate slots per staging row
    const size_t sbytes = (size_t)3 * cap4 * 12 + (size_t)P.m * kForceThreads * 4;
#define PLIFE_LAUNCH_STAGED(KIND, FAST)                                                                          \
    do {                                                                                                         \
        auto kfn = force_kernel_staged4<KIND, FAST>;                                                             \
        if (sbytes > 48 * 1024) {                                                                                \
            cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbytes);  \
            if (e != cudaSuccess) return e;                                                                      \
        }                                                                                                        \
        kfn<<<nblocks, kForceThreads, sbytes, stream>>>(io, cell_end, cell_sorted, P, gM, cap4, nbin);           \
    } while (0)
    switch (kind) {
    case PLIFE_ACC_PARTICLE_LIFE: PLIFE_LAUNCH_STAGED(PLIFE_ACC_PARTICLE_LIFE, true); break;
    case PLIFE_ACC_PARTICLE_LIFE_R: PLIFE_LAUNCH_STAGED(PLIFE_ACC_PARTICLE_LIFE_R, false); break;
    case PLIFE_ACC_PARTICLE_LIFE_R2: PLIFE_LAUNCH_STAGED(PLIFE_ACC_PARTICLE_LIFE_R2, false); break;
    case PLIFE_ACC_ROTATOR_90: PLIFE_LAUNCH_STAGED(PLIFE_ACC_ROTATOR_90, false); break;
    case PLIFE_ACC_ROTATOR_ATTR: PLIFE_LAUNCH_STAGED(PLIFE_ACC_ROTATOR_ATTR, false); break;
    case PLIFE_ACC_PLANETS: PLIFE_LAUNCH_STAGED(PLIFE_ACC_PLANETS, false); break;
    default: return cudaErrorInvalidValue;
    }
#undef PLIFE_LAUNCH_STAGED
    return cudaGetLastError();
}

// ---- small particle counts: one warp per cell, candidates spread over the lanes (fp32) -------------------------------
// With ten thousand particles (BASELINE config 1: the app's own default) the staged kernel has one warp per scheduler
// running a 4 500-instruction dependent stream, and one lane in six sits on the periodic seam and walks global memory one
// L2 round trip per candidate: 43 us for 0.3 us worth of arithmetic.  Here a warp owns a CELL: it copies the particles
// of the cell's 9 neighbour cells - in the reference's order and with its cell wrapping (B/Physics.java:407-421), so seam
// cells and the duplicate visits of tiny grids need no special case - into its slice of shared memory, then LPT lanes share
// one target: each takes every LPT-th candidate with the exact per-pair min-image (wrap_connection), and a shuffle adds the
// partial sums.  A target whose un-clamped coordinates are not its cell's (x or y exactly 1.0, the fat strip) takes the
// literal walk.  Only used on the plain cell list (ks == 0), so the compute order is the reference's order.
constexpr int kCellsLpt = 4; // lanes per target

template <int KIND, bool FAST>
__global__ void __launch_bounds__(kForceThreads) force_kernel_cells(IOF32 io, const int32_t *__restrict__ cell_end,
                                                                   const int32_t *__restrict__ cell_sorted, ForceParams<float> P,
                                                                   const float *__restrict__ gMt, int capw, int wpc, NextBin nb)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int kWarps = kForceThreads / 32;
    float4 *buf = reinterpret_cast<float4 *>(smem_raw) + (size_t)(threadIdx.x >> 5) * capw; // this warp's candidates
    float *sM = reinterpret_cast<float *>(smem_raw + (size_t)kWarps * capw * 16);
    const Grid g = P.g;
    load_matrix_smem(sM, gMt, P.m, FAST ? P.fast_a_scale : 1.0f);

    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int cell = w / wpc, part = w % wpc;
    if (cell >= g.nx * g.ny) return;
    const int ccx = cell % g.nx, ccy = cell / g.nx;
    int cs, ce;
    cell_span(cell_end, cell, cell, 0, cs, ce);
    const int nt = ce - cs;
    constexpr int TPP = 32 / kCellsLpt; // targets per pass
    if (part * TPP >= nt) return;
    // the 9 neighbour cells, lane k < 9 holds cell k of the reference's order (:88-98, :407-421)
    int s_k = 0, len_k = 0;
    if (lane < 9) {
        const int ox = lane % 3 - 1, oy = lane / 3 - 1;
        int cx = wrap_container(ccx + ox, g.nx), cy = wrap_container(ccy + oy, g.ny);
        bool use = true;
        if (P.wrap) {
            cx = wrap_container(cx, g.nx);
            cy = wrap_container(cy, g.ny);
        } else if (cx < 0 || cx >= g.nx || cy < 0 || cy >= g.ny) {
            use = false;
        }
        if (use) {
            int e_k;
            cell_span(cell_end, cx + cy * g.nx, cx + cy * g.nx, 0, s_k, e_k);
            len_k = e_k - s_k;
        }
    }
    int off_k = len_k; // exclusive prefix of the lengths over lanes 0..8
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, off_k, o);
        if (lane >= o) off_k += t;
    }
    const int T = __shfl_sync(0xffffffffu, off_k, 8); // all candidates of the cell's targets
    off_k -= len_k;
    const bool single = T <= capw - 2 * kCellsLpt * 4;

    const int target = lane % TPP, sub = lane / TPP;
    bool loaded = false;
#pragma unroll 1
    for (int t0 = part * TPP; t0 < nt; t0 += wpc * TPP) {
        const int t = t0 + target;
        const bool active = t < nt;
        const int i = cs + (active ? t : 0); // sorted index == reference slot (ks == 0, single GPU)
        const float4 me = __ldg(io.pt + i);
        const int own = __float_as_int(me.z) >> kTypeShift;
        const int cxy = __ldg(cell_sorted + i);
        const bool plain = (cxy & 0xffff) == ccx && (cxy >> 16) == ccy; // else: un-clamped coordinates differ from the cell's
        MatrixView<float, kMatShared> M{nullptr, sM, P.m, own, 1.0f, 0u, 0u};
        M.init();
        float vx, vy;
        io.self_vel(i, vx, vy);
        FastParticleLife32<kMatShared> vf(P, M);
        LiteralForce<float, KIND, kMatShared> vl{0.f, 0.f, P.r2, P.invr, P.k2, P.accp, M}; // this lane's share of the increments
#pragma unroll 1
        for (int c0 = 0; c0 < T; c0 += capw - 2 * kCellsLpt * 4) {
            const int clen = min(capw - 2 * kCellsLpt * 4, T - c0);
            if (!(single && loaded)) {
                __syncwarp();
                for (int base = 0; base < clen + kCellsLpt * 4; base += 32) { // uniform trip count: the shuffles need every lane
                    const int idx = base + lane, f = c0 + idx;
                    int src = -1;
#pragma unroll
                    for (int k = 0; k < 9; ++k) { // which of the 9 cells holds flat index f
                        const int ok = __shfl_sync(0xffffffffu, off_k, k), sk = __shfl_sync(0xffffffffu, s_k, k),
                                  lk = __shfl_sync(0xffffffffu, len_k, k);
                        if (idx < clen && f >= ok && f < ok + lk) src = sk + f - ok;
                    }
                    float4 q = make_float4(1.0e9f, 1.0e9f, 0.f, 0.f); // padding: far away
                    if (src >= 0) {
                        q = __ldg(io.pt + src);
                        q.z = __int_as_float(__float_as_int(q.z) >> kTypeShift);
                    }
                    if (idx < clen + kCellsLpt * 4) buf[idx] = q;
                }
                loaded = true;
                __syncwarp();
            }
            if (active && plain) {
                for (int j = sub; j < clen; j += kCellsLpt * 4) { // 4 candidates of this lane per trip (padded)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 q = buf[j + u * kCellsLpt];
                        const float dx = P.wrap ? wrap_connection(me.x, q.x) : q.x - me.x; // :464-467
                        const float dy = P.wrap ? wrap_connection(me.y, q.y) : q.y - me.y;
                        const Cand<float> cq{q.x, q.y, __float_as_int(q.z), 0u};
                        if constexpr (FAST) vf.pair(-1, cq, dx, dy);
                        else vl.pair(-1, cq, dx, dy);
                    }
                }
            }
        }
        float ax, ay;
        if constexpr (FAST) {
            ax = vf.ax();
            ay = vf.ay();
        } else {
            ax = vl.vx;
            ay = vl.vy;
        }
#pragma unroll
        for (int o = TPP; o < 32; o <<= 1) {
            ax += __shfl_xor_sync(0xffffffffu, ax, o);
            ay += __shfl_xor_sync(0xffffffffu, ay, o);
        }
        if (!active || sub != 0) continue;
        float nvx, nvy;
        if (!plain) { // rare: the literal 9-cell walk around the un-clamped coordinates
            const int cx0 = cxy & 0xffff, cy0 = cxy >> 16;
            MatrixView<float, kMatGlobal> Mg{gMt, nullptr, P.m, own, FAST ? P.fast_a_scale : 1.0f, 0u, 0u};
            if constexpr (FAST) {
                FastParticleLife32<kMatGlobal> v2(P, Mg);
                traverse_global32_plain(io, cell_end, g, P.wrap, i, me.x, me.y, cx0, cy0, v2);
                ax = v2.ax();
                ay = v2.ay();
            } else {
                LiteralForce<float, KIND, kMatGlobal> v2{0.f, 0.f, P.r2, P.invr, P.k2, P.accp, Mg};
                traverse_global32_plain(io, cell_end, g, P.wrap, i, me.x, me.y, cx0, cy0, v2);
                ax = v2.vx;
                ay = v2.vy;
            }
        }
        if constexpr (FAST) {
            nvx = fmaf(P.fast_k, ax, vx * P.mu);
            nvy = fmaf(P.fast_k, ay, vy * P.mu);
        } else {
            nvx = vx * P.mu + ax;
            nvy = vy * P.mu + ay;
        }
        float nx_ = fmaf(nvx, P.dt, me.x);
        float ny_ = fmaf(nvy, P.dt, me.y);
        if (P.wrap) {
            nx_ = range_wrap(nx_);
            ny_ = range_wrap(ny_);
        } else {
            nx_ = range_clamp(nx_);
            ny_ = range_clamp(ny_);
        }
        io.store(i, nx_, ny_, nvx, nvy, own, __float_as_uint(me.w));
        nb.add(i, nx_, ny_, nvx, nvy, own, __float_as_uint(me.w), g);
    }
}

inline cudaError_t dispatch_force_cells(const IOF32 &io, const int32_t *cell_end, const int32_t *cell_sorted, const ForceParams<float> &P,
                                        const float *gMt, int kind, NextBin nbin, cudaStream_t stream)
{
    const int ncell = P.g.nx * P.g.ny;
    // warps per cell: enough to fill the machine several times over; a warp whose first pass lies beyond the cell's
    // particles leaves at once, so fuller cells simply keep more of their warps
    int wpc = (4800 + ncell - 1) / ncell;
    if (wpc > 16) wpc = 16;
    if (wpc < 1) wpc = 1;
    int capw = (int)(9.0f * P.g.rho * 1.5f) + 64;
    capw = (capw + 31) / 32 * 32;
    if (capw > 1024) capw = 1024;
    const int nwarps = ncell * wpc;
    const int nblocks = (nwarps + kForceThreads / 32 - 1) / (kForceThreads / 32);
    const size_t sbytes = (size_t)(kForceThreads / 32) * capw * 16 + (size_t)P.m * P.m * 4;
#define PLIFE_LAUNCH_CELLS(KIND, FAST)                                                                           \
    do {                                                                                                         \
        auto kfn = force_kernel_cells<KIND, FAST>;                                                               \
        if (sbytes > 48 * 1024) {                                                                                \
            cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbytes);  \
            if (e != cudaSuccess) return e;                                                                      \
        }                                                                                                        \
        kfn<<<nblocks, kForceThreads, sbytes, stream>>>(io, cell_end, cell_sorted, P, gMt, capw, wpc, nbin);     \
    } while (0)
    switch (kind) {
    case PLIFE_ACC_PARTICLE_LIFE: PLIFE_LAUNCH_CELLS(PLIFE_ACC_PARTICLE_LIFE, true); break;
    case PLIFE_ACC_PARTICLE_LIFE_R: PLIFE_LAUNCH_CELLS(PLIFE_ACC_PARTICLE_LIFE_R, false); break;
    case PLIFE_ACC_PARTICLE_LIFE_R2: PLIFE_LAUNCH_CELLS(PLIFE_ACC_PARTICLE_LIFE_R2, false); break;
    case PLIFE_ACC_ROTATOR_90: PLIFE_LAUNCH_CELLS(PLIFE_ACC_ROTATOR_90, false); break;
    case PLIFE_ACC_ROTATOR_ATTR: PLIFE_LAUNCH_CELLS(PLIFE_ACC_ROTATOR_ATTR, false); break;
    case PLIFE_ACC_PLANETS: PLIFE_LAUNCH_CELLS(PLIFE_ACC_PLANETS, false); break;
    default: return cudaErrorInvalidValue;
    }
#undef PLIFE_LAUNCH_CELLS
    return cudaGetLastError();
}

#endif // PLIFE_FORCE_F32_TU

template <typename IO>
__global__ void __launch_bounds__(kForceThreads) neighbors_kernel(IO io, const int32_t *__restrict__ cell_end,
                                                                 const int32_t *__restrict__ cell_sorted,
                                                                 ForceParams<typename IO::R> P, int32_t *__restrict__ cnt,
                                                                 unsigned long long *__restrict__ hash)
{
    using R = typename IO::R;
    const int i = blockIdx.x * kForceThreads + threadIdx.x;
    if (i >= P.n) return;
    const Cand<R> self = io.cand(P.first + i);
    NeighborVisitor<R> v{{0, 0ull}, P.r2};
    const int cxy = __ldg(cell_sorted + i);
    traverse(io, cell_end, P.g, P.wrap, P.first + i, self.x, self.y, cxy, v);
    const int o = io.out_slot(i, cxy, cell_end, P.g); // reported in the reference's particle order
    cnt[o] = v.d.count;
    hash[o] = v.d.hash;
}

template <typename IO>
__global__ void __launch_bounds__(kForceThreads) pair_count_kernel(IO io, const int32_t *__restrict__ cell_end,
                                                                  const int32_t *__restrict__ cell_sorted,
                                                                  ForceParams<typename IO::R> P,
                                                                  unsigned long long *__restrict__ total)
{
    using R = typename IO::R;
    __shared__ unsigned long long sh[kForceThreads / 32];
    int i0 = 0, lim = 0;
    cta_targets(P, i0, lim);
    const int i = i0 + threadIdx.x;
    unsigned long long c = 0;
    if (i < lim) {
        const Cand<R> self = io.cand(P.first + i);
        PairCountVisitor<R> v{P.first + i, 0ull};
        traverse(io, cell_end, P.g, P.wrap, P.first + i, self.x, self.y, __ldg(cell_sorted + i), v);
        c = v.count;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < kForceThreads / 32; w++) s += sh[w];
        atomicAdd(total, s);
    }
}

// dispatch over accelerator kind / matrix placement
template <typename IO, bool FAST_OK>
cudaError_t dispatch_force(const IO &io, const int32_t *cell_end, const int32_t *cell_sorted,
                           const ForceParams<typename IO::R> &P, int nblocks, const typename IO::R *gMt, int kind, NextBin nbin,
                           cudaStream_t stream)
{
    using R = typename IO::R;
    if (nblocks <= 0) return cudaSuccess;
    const int nb = nblocks;
    const bool smem = P.use_smem_matrix != 0;
    const size_t sbytes = smem ? sizeof(R) * (size_t)P.m * P.m : 0;
#define PLIFE_LAUNCH(KIND, SM, FAST)                                                                            \
    do {                                                                                                        \
        auto kfn = force_kernel<IO, KIND, SM, FAST>;                                                            \
        if (sbytes > 48 * 1024) {                                                                               \
            cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sbytes); \
            if (e != cudaSuccess) return e;                                                                     \
        }                                                                                                       \
        kfn<<<nb, kForceThreads, sbytes, stream>>>(io, cell_end, cell_sorted, P, gMt, nbin);                    \
    } while (0)
#define PLIFE_KIND(KIND)                       \
    do {                                       \
        if (smem) PLIFE_LAUNCH(KIND, true, false); \
        else PLIFE_LAUNCH(KIND, false, false);     \
    } while (0)
    switch (kind) {
    case PLIFE_ACC_PARTICLE_LIFE:
        if (FAST_OK) {
            if (smem) PLIFE_LAUNCH(PLIFE_ACC_PARTICLE_LIFE, true, FAST_OK);
            else PLIFE_LAUNCH(PLIFE_ACC_PARTICLE_LIFE, false, FAST_OK);
        } else {
            PLIFE_KIND(PLIFE_ACC_PARTICLE_LIFE);
        }
        break;
    case PLIFE_ACC_PARTICLE_LIFE_R: PLIFE_KIND(PLIFE_ACC_PARTICLE_LIFE_R); break;
    case PLIFE_ACC_PARTICLE_LIFE_R2: PLIFE_KIND(PLIFE_ACC_PARTICLE_LIFE_R2); break;
    case PLIFE_ACC_ROTATOR_90: PLIFE_KIND(PLIFE_ACC_ROTATOR_90); break;
    case PLIFE_ACC_ROTATOR_ATTR: PLIFE_KIND(PLIFE_ACC_ROTATOR_ATTR); break;
    case PLIFE_ACC_PLANETS: PLIFE_KIND(PLIFE_ACC_PLANETS); break;
    default: return cudaErrorInvalidValue;
    }
#undef PLIFE_KIND
#undef PLIFE_LAUNCH
    return cudaGetLastError();
}

} // namespace plife
