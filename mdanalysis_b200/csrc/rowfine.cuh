// rowfine.cuh -- InterRDF pair histogram for orthorhombic cells, fine cells along x.
//
// Successor of rowquad.cuh for the headline workload.  Reference behaviour:
// package/MDAnalysis/analysis/rdf.py:268-291 (capped_distance + np.histogram per frame),
// lib/nsgrid.pyx:483-644 (which pairs a search sees).  Same exactness argument as
// rowtile.cuh / rowquad.cuh: the FP32 value only bins pairs that are provably away from
// every bin edge; every other pair is re-evaluated in the reference's float64 arithmetic.
//
// What ncu said about rowquad.cuh (profiles/r01h_ncu_rdf_row_quad_kernel.txt): the
// shared-memory data pipe was 73 % busy (12 wavefronts of LDS.128 + 9 of RED per four
// candidates), 61 issue slots per four candidates, and 73 % of the candidates were misses
// (a 5 x 5 x 5 block of cells holds 3.7 x the volume of the cutoff sphere).  So:
//
//  * Cells are FINE ALONG X only (cutoff / 8 wide; y and z keep cutoff / 2).  The atoms of
//    a (y, z) row are then sorted by x to within one fine cell, and a query's window in a
//    stencil row is [x - h, x + h] rounded out to fine cells with h^2 = cutoff^2 - (gap of
//    the query to the row's cells in y and z)^2 -- per LANE, from its own position.  The
//    candidates per query drop from 880 to ~550 (236 of them are hits).
//  * A row is staged as ONE run in "extended" order: the cells left of the periodic edge
//    are staged with x - L, those right of it with x + L (added while staging), so the
//    walk has no per-run loop and no per-run image arithmetic.  The y / z translation of
//    the row still goes onto the query.
//  * Decision by clamping instead of predicates: the bin coordinate is clamped to
//    nbins + 1 (FMNMX), so a miss lands in a scratch row of the histogram; the histogram
//    is bin-major with one column per lane (word = bin * 32 + lane): every RED of a warp
//    is conflict free (1 wavefront instead of 2.2) and misses need no private dummy word.
//    A pair inside a decision band is detected by the sign of fr^2 - band^2 (packed FFMA,
//    two LOP3 and one branch per four candidates); it is taken OUT of the bin the FP32
//    path put it into and queued for the float64 evaluation.
//  * 48 issue slots per four candidates (was 61).
//  * Triclinic cells (MIC = MIC_TRICLINIC; only on the library's own grid, where the reference's
//    stencil is complete, rdf.cu).  The walk is Cartesian and needs nothing but the row's full
//    translation ib b + ic c on the query.  The planner works in the cell coordinates of
//    geom.cuh cell_of(): x' = (x - t_xy y - t_xz z) / cs_xx, y' = (y - t_yz z) / cs_yy,
//    z' = z / cs_zz, in which the lattice images are whole grids (t_xy c_y = 0 is required), a
//    stencil row is the unit square [dy, dy + 1) x [dz, dz + 1) relative to the query's cell,
//    and the Cartesian separation of a pair is dz = cs_zz dz', dy = cs_yy dy' + t_yz dz,
//    dx = cs_xx dx' + t_xy dy + t_xz dz.  Interval arithmetic over the square gives the gaps
//    (-> h as before) and the range of the shear t_xy dy + t_xz dz, by which the window in x'
//    is widened: [x' + (-h - s_max) / cs_xx, x' + (h - s_min) / cs_xx].
#pragma once
#include "rowquad.cuh"

namespace mdagpu {

#ifdef MDAGPU_FINE_TIMING
__device__ unsigned long long g_fine_phase[16];
#define FINE_MARK(k) do { if (threadIdx.x == 0) { long long t_ = clock64(); atomicAdd(&g_fine_phase[k], (unsigned long long)(t_ - fine_t0)); fine_t0 = t_; } } while (0)
#else
#define FINE_MARK(k) do { } while (0)
#endif

constexpr int RF_CAP = 3584;            // staged candidates per CTA (three planes of floats: 42 kB)
constexpr int RF_MAXROWS = 25;          // (2W+1)^2, W <= 2
constexpr int RF_MAXEXT = 112;          // staged fine cells per row: home segment + 2 WX
constexpr int RF_SLOW = 40;             // queued exact evaluations per warp
constexpr int RF_LIST = 2816;           // window records per batch of queries
constexpr int RF_WMAX = 64;             // quads per window record (longer windows are split)

struct FineParams {
    int W;              // stencil half-width in rows along y and z (cells >= cutoff / W)
    int WX;             // stencil half-width in fine cells along x (cells >= cutoff / WX)
    int R;              // home fine cells per CTA
    float reach;        // cutoff + safety margin: the per-lane windows are cut at this distance
    int debug;          // profiling aid: 1 = skip the walk (results are wrong)
};

// One stencil row as staged: atoms of the fine cells [a0, a1) of a grid row in extended
// order -- up to three contiguous runs of the sorted atom array (image -L, in range, image +L).
struct FineRow {
    uint32_t g[3];      // global slot of the first atom of each run
    uint32_t end[3];    // extended position where each run ends
    float xs[3];        // translation added to x while staging
    uint32_t len;       // staged candidates of the row
    uint32_t base;      // tile position of the row (or its piece) in the current group
    float ay, by, az, bz;   // gap of a query to the row's cells: gy = ay * q.y + by, gz = az * q.z + bz
    float sy, sz;       // lattice translation of the row along y, z (goes onto the query)
    int64_t rowbase;    // global cell index of the grid row's cell 0
};

struct FineConst {
    float inv_dr, c0, cap, nbl2;     // bin coordinate um = d * inv_dr + c0; clamp; -(band_lim^2)
    uint32_t hb, stride, wgt;        // RED address = hb + stride * bits(um + magic)
};

struct FineQuad { float4 cx, cy, cz; };                           // four staged candidates: x, y, z planes

template <int CAP>
__device__ __forceinline__ FineQuad fine_load(uint32_t ptr)
{
    constexpr uint32_t PL = CAP * 4;              // bytes between the planes
    FineQuad c;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c.cx.x), "=f"(c.cx.y), "=f"(c.cx.z), "=f"(c.cx.w) : "r"(ptr));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c.cy.x), "=f"(c.cy.y), "=f"(c.cy.z), "=f"(c.cy.w) : "r"(ptr + PL));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c.cz.x), "=f"(c.cz.y), "=f"(c.cz.z), "=f"(c.cz.w) : "r"(ptr + 2u * PL));
    return c;
}

// Four candidates of one query: distance, bin coordinate, histogram update.  Returns the AND of
// the four sign words (negative iff every pair is safely inside a bin); a[] = the RED addresses.
template <bool HOME>
__device__ __forceinline__ int fine_quad(const FineQuad &c, const float2 &nqx2, const float2 &nqy2, const float2 &nqz2,
                                         int ahead, const FineConst &kc, uint32_t (&a)[4], float2 (&sg)[2])
{
    const float2 inv2 = make_float2(kc.inv_dr, kc.inv_dr), c02 = make_float2(kc.c0, kc.c0);
    const float2 magic2 = make_float2(12582912.0f, 12582912.0f), nmagic2 = make_float2(-12582912.0f, -12582912.0f);
    const float2 neg1 = make_float2(-1.0f, -1.0f), nbl2 = make_float2(kc.nbl2, kc.nbl2);
    float2 tt[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float2 x2 = h ? make_float2(c.cx.z, c.cx.w) : make_float2(c.cx.x, c.cx.y);
        const float2 y2 = h ? make_float2(c.cy.z, c.cy.w) : make_float2(c.cy.x, c.cy.y);
        const float2 z2 = h ? make_float2(c.cz.z, c.cz.w) : make_float2(c.cz.x, c.cz.y);
        const float2 dx = __fadd2_rn(x2, nqx2), dy = __fadd2_rn(y2, nqy2), dz = __fadd2_rn(z2, nqz2);
        const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __fmul2_rn(dz, dz)));
        float2 d;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d.x) : "f"(d2.x));
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d.y) : "f"(d2.y));
        // bin coordinate um = d * inv_dr + c0: bin k <=> um in [k - 0.5, k + 0.5), clamped to
        // the scratch row nbins + 1 (an exact integer: never inside a band)
        float2 u = __ffma2_rn(d, inv2, c02);
        u.x = fminf(u.x, kc.cap);
        u.y = fminf(u.y, kc.cap);
        if (HOME) {
            if (!(2 * h < ahead)) u.x = kc.cap;
            if (!(2 * h + 1 < ahead)) u.y = kc.cap;
        }
        // the magic number add rounds to the bin (its low mantissa bits), fr = um - bin in
        // [-0.5, 0.5]; sg = fr^2 - band_lim^2 is negative iff the pair is safely inside a bin
        tt[h] = __fadd2_rn(u, magic2);
        const float2 fr = __ffma2_rn(__fadd2_rn(tt[h], nmagic2), neg1, u);
        sg[h] = __ffma2_rn(fr, fr, nbl2);
    }
    a[0] = kc.hb + kc.stride * __float_as_uint(tt[0].x); a[1] = kc.hb + kc.stride * __float_as_uint(tt[0].y);
    a[2] = kc.hb + kc.stride * __float_as_uint(tt[1].x); a[3] = kc.hb + kc.stride * __float_as_uint(tt[1].y);
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[0]), "r"(kc.wgt) : "memory");
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[1]), "r"(kc.wgt) : "memory");
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[2]), "r"(kc.wgt) : "memory");
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[3]), "r"(kc.wgt) : "memory");
    return __float_as_int(sg[0].x) & __float_as_int(sg[0].y) & __float_as_int(sg[1].x) & __float_as_int(sg[1].y);
}

// Walks quads [qa, qe) of the tile for TWO queries that share the window (nqa, nqb = minus the
// queries, translated): the shared-memory data pipe was 96 % busy with one query per lane (three
// LDS.128 of ~8 wavefronts each per four pairs -- the lanes of a warp read unrelated windows);
// two queries get eight pairs out of the same loads.  HOME: only candidates behind each query's
// own position (`ahead` = own position of the first query - position of quad qa; the second
// query sits right behind the first).  A pair inside a decision band is handed to
// slow_fn(which query, quad, k, addr).
template <bool HOME, int CAP, class SlowFn>
__device__ __forceinline__ void fine_walk2(uint32_t planes, uint32_t qa, uint32_t qe,
                                           float3 nqa, float3 nqb, int ahead, const FineConst &kc, SlowFn slow_fn)
{
    const float2 ax = make_float2(nqa.x, nqa.x), ay = make_float2(nqa.y, nqa.y), az = make_float2(nqa.z, nqa.z);
    const float2 bx = make_float2(nqb.x, nqb.x), by = make_float2(nqb.y, nqb.y), bz = make_float2(nqb.z, nqb.z);
    const uint32_t end = planes + 16u * qe;
    // (the divergent region stays INSIDE the loop: a lane that left the loop for its band pair
    // would walk the rest of its window on its own)
    auto slow4 = [&](int which, uint32_t at, const uint32_t (&a)[4], const float2 (&sg)[2]) {
        const uint32_t quad = (at - planes) >> 4;
        if (__float_as_int(sg[0].x) >= 0) slow_fn(which, quad, 0, a[0]);
        if (__float_as_int(sg[0].y) >= 0) slow_fn(which, quad, 1, a[1]);
        if (__float_as_int(sg[1].x) >= 0) slow_fn(which, quad, 2, a[2]);
        if (__float_as_int(sg[1].y) >= 0) slow_fn(which, quad, 3, a[3]);
    };
    if (HOME) {
        for (uint32_t ptr = planes + 16u * qa; ptr < end; ptr += 16u, ahead -= 4) {
            const FineQuad c = fine_load<CAP>(ptr);
            uint32_t a[4], b[4];
            float2 sa[2], sb[2];
            const int na = fine_quad<true>(c, ax, ay, az, ahead, kc, a, sa);
            const int nb = fine_quad<true>(c, bx, by, bz, ahead + 1, kc, b, sb);
            if ((na & nb) >= 0) {
                if (na >= 0) slow4(0, ptr, a, sa);
                if (nb >= 0) slow4(1, ptr, b, sb);
            }
        }
    } else {
        // The windows of a warp's lanes lie anywhere in the tile: eight lanes of a quarter warp hit
        // the eight 16-byte bank groups at random (measured: 9 wavefronts per LDS.128, the data pipe
        // 88 % busy).  The order within a window is free, so lane l starts at the first quad of its
        // window that lies in bank group l mod 8 and wraps around at the end: lanes that advance
        // together stay in different groups (until they wrap, windows of a warp being equally long).
        const uint32_t first = planes + 16u * qa;
        const uint32_t n = qe - qa;
        uint32_t rot = ((threadIdx.x & 7u) - (first >> 4)) & 7u;
        if (rot >= n) rot = 0u;
        uint32_t ptr = first + 16u * rot;
        for (uint32_t left = n; left; --left) {
            const FineQuad c = fine_load<CAP>(ptr);
            uint32_t a[4], b[4];
            float2 sa[2], sb[2];
            const int na = fine_quad<false>(c, ax, ay, az, 0, kc, a, sa);
            const int nb = fine_quad<false>(c, bx, by, bz, 0, kc, b, sb);
            if ((na & nb) >= 0) {
                if (na >= 0) slow4(0, ptr, a, sa);
                if (nb >= 0) slow4(1, ptr, b, sb);
            }
            ptr += 16u;
            if (ptr == end) ptr = first;
        }
    }
}

// ---- planner -> walker hand-over (global scratch), one unit per (frame, (y, z) row, x segment) ----
struct FineRowA { uint32_t g0, g1, g2, base; };                    // global slot of the first atom of each run; tile position of the row
struct FineRowB { uint32_t end0, end1, len, pad; };                // extended positions where the runs end; staged candidates
struct FineRowC { float sy, sz, xl, sx; };                         // lattice translation of the row (y, z; x for triclinic cells); box edge along x
struct FineUnitHead {
    uint32_t nrec, hstart;      // window records; where the home-row records start (a multiple of 32)
    uint32_t qs0, nq;           // first query slot, queries
    uint32_t fill;              // staged candidates (rows padded to quads)
    uint32_t own0;              // tile position of the first query (SELF: the queries are in the tile)
    uint32_t pad[2];
};
template <int NR> struct FineUnit {
    FineUnitHead h;
    FineRowA a[NR];
    FineRowB b[NR];
    FineRowC c[NR];
};
// bits of *flag set by the planner: the walker cannot take this call (the caller falls back)
constexpr unsigned FINE_OVER_TILE = 2u, FINE_OVER_LIST = 4u, FINE_OVER_QUERIES = 8u;

// Planner: everything the walk needs to know, computed with nothing but bookkeeping resident
// (no tile: six CTAs per SM hide the chains of dependent loads that cost the fused kernel
// 73 000 cycles per CTA): row descriptors, and the (pair of queries, row) windows as records
//   quads | first quad << 7 | pair << 17 | row << 25
// counting-sorted by length, longest first, home-row records behind the others.
template <bool SELF, int T, int MIC = MIC_ORTHO_LAZY>
__global__ void __launch_bounds__(T, 6)
fine_plan_kernel(const float4 *__restrict__ q_atoms, const uint32_t *__restrict__ q_start,
                 const uint32_t *__restrict__ c_start, const FrameGeom *__restrict__ fgs, int per_frame,
                 const __grid_constant__ FrameGeom fg_const, FineParams fp, int cap, int list_cap,
                 FineUnit<SELF ? 13 : 25> *__restrict__ units, uint32_t *__restrict__ lists, unsigned *__restrict__ flag,
                 const unsigned *__restrict__ bad0, const unsigned *__restrict__ bad1)
{
    // a cell list with folded atoms (engine.cuh: inconsistent) is not planned: the host sends the
    // piece to the per-pair kernels when it reads the flags, one piece later
    if ((bad0 && *bad0) || (bad1 && *bad1)) return;
    constexpr int NW = T / 32;
    constexpr int NR = SELF ? 13 : 25;                             // rows of the stencil (W = 2)
    constexpr int NB = 2 * RF_WMAX;                                // (home?, length) buckets
    constexpr int W = 2, side = 5, nrows = NR;
    extern __shared__ uint32_t eoff[];                             // extended position of every staged fine cell [NR][next + 1]
    __shared__ FrameGeom s_fg;
    __shared__ uint32_t s_raw[NR][8];                              // c_start values the row descriptors are made of
    __shared__ uint32_t s_qraw[RF_MAXEXT + 1];                     // first query slot of every home fine cell
    __shared__ float4 s_gap[NR];                                   // gap of a query to the row's cells: gy = ay * q.y + by, ...
    __shared__ uint32_t s_base[NR + 1], s_len[NR];
    __shared__ uint32_t s_cnt[NW][NB];
    __shared__ uint32_t s_tot[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t f = blockIdx.y;
    {
        const uint32_t *src = per_frame ? reinterpret_cast<const uint32_t *>(fgs + f) : reinterpret_cast<const uint32_t *>(&fg_const);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&s_fg);
        for (int w = tid; w < (int)(sizeof(FrameGeom) / 4); w += T) dst[w] = src[w];
    }
    for (int w = tid; w < NW * NB; w += T) (&s_cnt[0][0])[w] = 0u;
    // where this CTA is: one thread divides (r02c line profile: the three runtime divisions, done by
    // every thread, were 5.6 % of the kernel's instructions)
    __shared__ int s_where[4];                                     // runs, x run, y, z
    if (tid == 0) {
        const GridGeom &gg = (per_frame ? fgs + f : &fg_const)->grid;
        const int runs0 = (gg.ncells[0] + fp.R - 1) / fp.R;
        const int rest = (int)blockIdx.x / runs0;
        s_where[0] = runs0;
        s_where[1] = (int)blockIdx.x - rest * runs0;
        s_where[2] = rest % gg.ncells[1];
        s_where[3] = rest / gg.ncells[1];
    }
    __syncthreads();
    const int WX = fp.WX, R = fp.R;
    const int nx = s_fg.grid.ncells[0], ny = s_fg.grid.ncells[1], nz = s_fg.grid.ncells[2];
    const int runs = s_where[0];
    FineUnit<NR> &unit = units[(uint64_t)f * gridDim.x + blockIdx.x];
    if (blockIdx.x >= (unsigned)(runs * ny * nz)) {                // frames may have fewer cells (uniform)
        if (tid == 0) unit.h.nq = 0u;
        return;
    }
    const int xr = s_where[1], y = s_where[2], z = s_where[3];
    const int c0h = xr * R, c1h = min(c0h + R, nx), nh = c1h - c0h;
    const int a0 = c0h - WX, next = nh + 2 * WX;                   // staged fine cells [a0, a0 + next)
    const int64_t cbase = per_frame ? s_fg.cell_base : (int64_t)f * s_fg.grid.ncell_total;
    const BoxParams &box = s_fg.box;
    const float wy = (float)s_fg.grid.cs[4], wz = (float)s_fg.grid.cs[8];
    const float inv_wx = (float)(1.0 / s_fg.grid.cs[0]);
    const int64_t home_row = cbase + ((int64_t)z * ny + y) * nx;
    auto row_dydz = [&](int r, int &dy, int &dz) {
        if (SELF) {
            if (r <= W) { dy = r; dz = 0; }
            else { dz = 1 + (r - (W + 1)) / side; dy = (r - (W + 1)) % side - W; }
        } else { dz = r / side - W; dy = r % side - W; }
    };
    auto row_cells = [&](int r, int &ib, int &ic) -> int64_t {     // global cell index of the grid row's cell 0
        int dy, dz;
        row_dydz(r, dy, dz);
        int zz = z + dz, yy = y + dy;
        ib = ic = 0;
        if (zz < 0) { zz += nz; ic = -1; } else if (zz >= nz) { zz -= nz; ic = 1; }
        if (yy < 0) { yy += ny; ib = -1; } else if (yy >= ny) { yy -= ny; ib = 1; }
        return cbase + ((int64_t)zz * ny + yy) * nx;
    };
    // ---- step 1: the c_start values every row descriptor is made of, the query offsets ----
    for (int idx = tid; idx < nrows * 8 + nh + 1; idx += T) {
        if (idx < nrows * 8) {
            const int r = idx >> 3, j = idx & 7;
            int ib, ic;
            const uint32_t *S = c_start + row_cells(r, ib, ic);
            const bool home = SELF && r == 0;
            const int a1 = home ? c1h : a0 + next;                 // a home row is only walked behind the query
            const int m0 = a0 < 0 ? 0 : a0, m1 = a1 > nx ? nx : a1;
            // j: 0 first atom of the left-image run, 1 of the in-range run, 2 of the right-image run (= row start),
            //    3 row end, 4 end of the in-range run, 5 end of the right-image run
            const int at = j == 0 ? (a0 < 0 ? nx + a0 : 0) : j == 1 ? m0 : j == 2 ? 0 : j == 3 ? nx : j == 4 ? m1
                                                                  : (a1 > nx ? a1 - nx : 0);
            if (j < 6) s_raw[r][j] = S[at];
        } else {
            const int e = idx - nrows * 8;
            s_qraw[e] = q_start[home_row + c0h + e];
        }
    }
    __syncthreads();
    const uint32_t qs0 = s_qraw[0], qs1 = s_qraw[nh], nq_all = qs1 - qs0;
    if (nq_all == 0u) {                                            // no queries (uniform)
        if (tid == 0) unit.h.nq = 0u;
        return;
    }
    // A thread plans for a PAIR of neighbouring queries (2 * pair, 2 * pair + 1: the walker evaluates
    // both against the same candidates) and for one half of the stencil rows.
    constexpr int HALF = T / 2, NRH = (NR + 1) / 2;
    const uint32_t pair = (uint32_t)(tid % HALF), npairs = (nq_all + 1u) >> 1;
    const int rlo = (tid / HALF) * NRH;
    const bool has_a = pair < npairs && tid < 2 * HALF, has_b = has_a && 2u * pair + 1u < nq_all;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), qb = q;
    if (has_a) q = q_atoms[qs0 + 2u * pair];
    if (has_b) qb = q_atoms[qs0 + 2u * pair + 1u];
    // ---- step 2: row descriptors; extended position of every staged fine cell ----
    auto row_lens = [&](int r, uint32_t &nL, uint32_t &nM, uint32_t &nR) {
        const bool home = SELF && r == 0;
        const int a1 = home ? c1h : a0 + next;
        nL = a0 < 0 ? s_raw[r][3] - s_raw[r][0] : 0u;
        nM = s_raw[r][4] - s_raw[r][1];
        nR = a1 > nx ? s_raw[r][5] - s_raw[r][2] : 0u;
    };
    if (warp == 0) {
        // one lane per row; the tile position of a row is the prefix of the padded lengths
        const int r = lane;
        uint32_t nL = 0, nM = 0, nR = 0;
        int dy = 0, dz = 0, ib = 0, ic = 0;
        if (r < nrows) {
            row_dydz(r, dy, dz);
            (void)row_cells(r, ib, ic);
            row_lens(r, nL, nM, nR);
        }
        const uint32_t len = nL + nM + nR, mine = ru4(len);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (r <= nrows) s_base[r] = incl - mine;                   // r == nrows: the total
        if (r < nrows) {
            s_len[r] = len;
            // gap to the cells of the row, in the query's (unwrapped) frame
            s_gap[r] = make_float4(dy > 0 ? -1.0f : (dy < 0 ? 1.0f : 0.0f),
                                   dy > 0 ? (float)(y + dy) * wy : (dy < 0 ? -(float)(y + dy + 1) * wy : 0.0f),
                                   dz > 0 ? -1.0f : (dz < 0 ? 1.0f : 0.0f),
                                   dz > 0 ? (float)(z + dz) * wz : (dz < 0 ? -(float)(z + dz + 1) * wz : 0.0f));
            unit.a[r] = { s_raw[r][0], s_raw[r][1], s_raw[r][2], incl - mine };
            unit.b[r] = { nL, nL + nM, len, 0u };
            if (MIC == MIC_TRICLINIC)                               // ib b + ic c; a = (h0, 0, 0)
                unit.c[r] = { (float)ib * box.h[4] + (float)ic * box.h[7], (float)ic * box.h[8], box.h[0],
                              (float)ib * box.h[3] + (float)ic * box.h[6] };
            else
                unit.c[r] = { (float)ib * box.L[1], (float)ic * box.L[2], box.L[0], 0.0f };
        }
    }
    const int ne1 = next + 1;
    // (a warp takes whole rows: the row's constants are computed once, its cells need no division)
    for (int r = warp; r < nrows; r += NW) {
        int ib, ic;
        const uint32_t *S = c_start + row_cells(r, ib, ic);
        uint32_t nL, nM, nR;
        row_lens(r, nL, nM, nR);
        const uint32_t g0 = s_raw[r][0], g1 = s_raw[r][1], g2 = s_raw[r][2], top = nL + nM + nR;
        uint32_t *eo = eoff + r * ne1;
        for (int e0 = lane; e0 < ne1; e0 += 128) {
            uint32_t val[4];
            const int groups = min(4, (ne1 - (e0 - lane) + 31) >> 5);   // uniform: 32-cell groups left in the row
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (u < groups) {
                    const int a = a0 + min(e0 + 32 * u, ne1 - 1);
                    val[u] = S[a < 0 ? nx + a : (a <= nx ? a : a - nx)];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + 32 * u, a = a0 + e;
                if (u < groups && e < ne1)
                    eo[e] = min(a < 0 ? val[u] - g0 : (a <= nx ? nL + (val[u] - g1) : nL + nM + (val[u] - g2)), top);
            }
        }
    }
    __syncthreads();
    const uint32_t fill = s_base[nrows];
    unsigned over = 0u;
    if (fill > (uint32_t)cap) over |= FINE_OVER_TILE;
    if (nq_all > (uint32_t)T) over |= FINE_OVER_QUERIES;
    // ---- this thread's windows (the union of the two queries' windows): first quad | quads << 10 ----
    uint32_t win[NRH];
    {
        const uint32_t pos = 2u * pair;
        auto home_cell = [&](uint32_t at) {                        // fine cell of the query at position `at`
            int e0 = 0, e1 = nh;                                   // s_qraw[e0] <= qs0 + at < s_qraw[e1]
            while (e1 - e0 > 1) {
                const int mid = (e0 + e1) >> 1;
                if (s_qraw[mid] <= qs0 + at) e0 = mid; else e1 = mid;
            }
            return e0;
        };
        // (the cell from the coordinate -- the sort key's own float64 arithmetic may differ by one at a
        // cell face, so the guess is checked against the cell starts and the search decides otherwise;
        // the two searches were 6.7 % of the kernel's instructions)
        auto home_cell_of = [&](const float4 &qq, uint32_t at) {
            int e = (int)(qq.x * inv_wx) - c0h;
            if (MIC == MIC_TRICLINIC) e = (int)((qq.x - (float)s_fg.grid.cs[3] * qq.y - (float)s_fg.grid.cs[6] * qq.z) * inv_wx) - c0h;
            e = min(max(e, 0), nh - 1);
            const uint32_t slot = qs0 + at;
            return (s_qraw[e] <= slot && slot < s_qraw[e + 1]) ? e : home_cell(at);
        };
        const int ea = has_a ? home_cell_of(q, pos) : 0, eb = has_b ? home_cell_of(qb, pos + 1u) : ea;
        const float reach2 = fp.reach * fp.reach;
        auto window = [&](const float4 &qq, int e0, const uint32_t *eo, const float4 &g, uint32_t &lo, uint32_t &hi) {
            const float gy = fmaf(g.x, qq.y, g.y), gz = fmaf(g.z, qq.z, g.w);
            const float h2 = fmaf(-gy, gy, fmaf(-gz, gz, reach2));
            float h;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(fmaxf(h2, 0.0f)));
            int k = (int)(h * inv_wx) + 1;
            k = k > WX ? WX : k;
            lo = eo[e0 + WX - k];
            hi = h2 > 0.0f ? eo[e0 + WX + k + 1] : lo;
        };
        // triclinic cells: the query in cell coordinates (see the head of the file)
        const float t_xy = (float)s_fg.grid.cs[3], t_xz = (float)s_fg.grid.cs[6], t_yz = (float)s_fg.grid.cs[7];
        const float inv_wy = (float)(1.0 / s_fg.grid.cs[4]), inv_wz = (float)(1.0 / s_fg.grid.cs[8]);
        const float wx_f = (float)s_fg.grid.cs[0];
        struct CellPos { float xc, fy, fz; };
        auto cell_pos = [&](const float4 &qq) {
            CellPos c;
            c.fz = qq.z * inv_wz - (float)z;
            c.fy = (qq.y - t_yz * qq.z) * inv_wy - (float)y;
            c.xc = (qq.x - t_xy * qq.y - t_xz * qq.z) * inv_wx;
            return c;
        };
        const CellPos cpa = MIC == MIC_TRICLINIC ? cell_pos(q) : CellPos{ 0.f, 0.f, 0.f };
        const CellPos cpb = MIC == MIC_TRICLINIC ? cell_pos(qb) : CellPos{ 0.f, 0.f, 0.f };
        auto window_tri = [&](const CellPos &c, int dy, int dz, const uint32_t *eo, uint32_t &lo, uint32_t &hi) {
            // Cartesian dz, dy over the row's unit square
            const float zlo = wz * ((float)dz - c.fz), zhi = zlo + wz;
            const float ya = wy * ((float)dy - c.fy);
            const float ylo = ya + fminf(t_yz * zlo, t_yz * zhi), yhi = ya + wy + fmaxf(t_yz * zlo, t_yz * zhi);
            const float gz = fmaxf(0.0f, fmaxf(zlo, -zhi)), gy = fmaxf(0.0f, fmaxf(ylo, -yhi));
            const float h2 = fmaf(-gy, gy, fmaf(-gz, gz, reach2));
            lo = hi = 0u;
            if (!(h2 > 0.0f)) return;
            float h;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(h) : "f"(h2));
            h *= 1.000001f;
            // shear of x over the part of the square a pair within reach can lie in
            const float zl = fmaxf(zlo, -fp.reach), zh = fminf(zhi, fp.reach), yl = fmaxf(ylo, -fp.reach), yh = fminf(yhi, fp.reach);
            const float slo = fminf(t_xy * yl, t_xy * yh) + fminf(t_xz * zl, t_xz * zh);
            const float shi = fmaxf(t_xy * yl, t_xy * yh) + fmaxf(t_xz * zl, t_xz * zh);
            const int cl = (int)floorf(c.xc + (-h - shi) * inv_wx - 1e-3f), ch = (int)floorf(c.xc + (h - slo) * inv_wx + 1e-3f);
            const int el = min(max(cl - a0, 0), next), eh = min(max(ch - a0 + 1, 0), next);
            lo = eo[el];
            hi = eo[eh];
        };
        (void)wx_f;
#pragma unroll
        for (int i = 0; i < NRH; ++i) {
            const int r = rlo + i;
            win[i] = 0u;
            if (r < NR && has_a && !over) {
                const bool home = SELF && r == 0;
                const uint32_t *eo = eoff + r * ne1;
                const float4 g = s_gap[r];
                uint32_t lo, hi;
                if (home) {
                    lo = eo[ea];
                    hi = eo[WX] + pos + (has_b ? 1u : 0u);         // the own extended position of the pair's last query
                } else {
                    int dy = 0, dz = 0;
                    if (MIC == MIC_TRICLINIC) { row_dydz(r, dy, dz); window_tri(cpa, dy, dz, eo, lo, hi); }
                    else window(q, ea, eo, g, lo, hi);
                    if (has_b) {
                        uint32_t lo2, hi2;
                        if (MIC == MIC_TRICLINIC) window_tri(cpb, dy, dz, eo, lo2, hi2);
                        else window(qb, eb, eo, g, lo2, hi2);
                        if (lo2 < hi2) {
                            if (lo < hi) { lo = min(lo, lo2); hi = max(hi, hi2); }
                            else { lo = lo2; hi = hi2; }
                        }
                    }
                }
                if (lo < hi) {
                    const uint32_t qa = (s_base[r] >> 2) + (lo >> 2), qe = (s_base[r] >> 2) + ((hi + 3u) >> 2);
                    win[i] = qa | ((qe - qa) << 10);
                }
            }
        }
    }
    // records of this thread: emit(bucket, record); windows longer than RF_WMAX quads are split
    auto records = [&](auto emit) {
#pragma unroll
        for (int i = 0; i < NRH; ++i) {
            if (!win[i]) continue;
            const int r = rlo + i;
            const bool home = SELF && r == 0;
            const uint32_t qa = win[i] & 1023u, qe = qa + (win[i] >> 10);
            for (uint32_t j = qa; j < qe; j += RF_WMAX) {
                const uint32_t len = min((uint32_t)RF_WMAX, qe - j);
                emit((home ? (uint32_t)RF_WMAX : 0u) + (uint32_t)RF_WMAX - len,
                     len | (j << 7) | (pair << 17) | ((uint32_t)r << 25));
            }
        }
    };
    // ---- counting sort by length, longest first (per warp counters: the lanes of a warp only
    // contend with each other).  (Tried: eight streams by first quad mod 8, interleaved, so that a
    // quarter warp reads eight different bank groups -- conflicts gone, but the lengths within a warp
    // matched less well: 24.7 instead of 30.6 active lanes.  The walker rotates its windows instead.) ----
    records([&](uint32_t b, uint32_t) { atomicAdd(&s_cnt[warp][b], 1u); });
    __syncthreads();
    if (warp == 0) {
        // exclusive prefix in (bucket, warp) order, four buckets per lane; lanes 16.. hold the
        // home-row buckets, which start at a multiple of 32 (a warp never mixes the two walks)
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int w = 0; w < NW; ++w) sum += s_cnt[w][4 * lane + k];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const uint32_t plain = __shfl_sync(0xffffffffu, incl, 15), all = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t hstart = (plain + 31u) & ~31u;
        uint32_t run = incl - sum + (lane >= 16 ? hstart - plain : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t v = s_cnt[w][4 * lane + k];
                s_cnt[w][4 * lane + k] = run;
                run += v;
            }
        if (lane == 0) { s_tot[0] = hstart; s_tot[1] = all + (hstart - plain); }
    }
    __syncthreads();
    const uint32_t hstart = s_tot[0], nrec = s_tot[1];
    if (nrec > (uint32_t)list_cap) over |= FINE_OVER_LIST;
    uint32_t *list = lists + ((uint64_t)f * gridDim.x + blockIdx.x) * (uint64_t)list_cap;
    if (!over) records([&](uint32_t b, uint32_t rec) { list[atomicAdd(&s_cnt[warp][b], 1u)] = rec; });
    __syncthreads();
    if (!over) {
        // (cursors have advanced to the ends of their runs: the last plain cursor is the plain total)
        const uint32_t plain = s_cnt[NW - 1][RF_WMAX - 1];
        for (uint32_t i = plain + tid; i < hstart; i += T) list[i] = 0u;      // padding in front of the home-row records
    }
    if (tid == 0) {
        unit.h.nrec = over ? 0u : nrec;
        unit.h.hstart = hstart;
        unit.h.qs0 = qs0;
        unit.h.nq = over ? 0u : nq_all;
        unit.h.fill = fill;
        unit.h.own0 = eoff[WX];                                    // row 0 starts at tile position 0
        if (over) atomicOr(flag, over);
    }
}

// Walker: stages the rows of its unit (x, y, z planes, extended order) and walks the planner's
// records, 32 at a time per warp.
template <bool SELF, class Sink, int CAP, int MIN_CTAS, int MIC = MIC_ORTHO_LAZY>
__global__ void __launch_bounds__(Sink::THREADS, MIN_CTAS)
fine_walk_kernel(const float4 *__restrict__ q_atoms, const float4 *__restrict__ c_atoms,
                 const FineUnit<SELF ? 13 : 25> *__restrict__ units, const uint32_t *__restrict__ lists, int list_cap,
                 const __grid_constant__ BoxParams box, SearchParams sp, Sink sink,
                 const unsigned *__restrict__ skip0, const unsigned *__restrict__ skip1, const unsigned *__restrict__ skip2)
{
    // the planner declined (skip0) or a cell list has folded atoms (skip1, skip2): nothing of this
    // piece is binned here -- the host reads the same flags and runs the piece through the other kernels
    if ((skip0 && *skip0) || (skip1 && *skip1) || (skip2 && *skip2)) return;
    constexpr int T = Sink::THREADS, NW = T / 32;
    constexpr int NR = SELF ? 13 : 25;
    // dynamic shared memory: [histogram] [x, y, z planes of the staged rows]
    extern __shared__ __align__(16) unsigned char fine_dyn[];
    float *tile = reinterpret_cast<float *>(fine_dyn + (((size_t)sink.hist_words() * sizeof(uint32_t) + 15) & ~(size_t)15));
    __shared__ __align__(16) FineUnit<NR> s_u;
    __shared__ __align__(16) float4 s_q[SELF ? 1 : T];              // queries of another group (SELF: they are in the tile)
    __shared__ uint2 s_slow[NW][RF_SLOW];                          // queued (query slot, candidate slot)
    __shared__ uint32_t s_nslow[NW];
    __shared__ uint32_t s_next;                                    // next unclaimed work item
    __shared__ FineConst s_kc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t f = blockIdx.y, uidx = f * gridDim.x + blockIdx.x;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(units + uidx);
        uint4 *dst = reinterpret_cast<uint4 *>(&s_u);
        for (int w = tid; w < (int)(sizeof(FineUnit<NR>) / 16); w += T) dst[w] = src[w];
    }
    sink.begin();
    if (lane == 0) s_nslow[warp] = 0u;
    if (tid == 0) {
        s_kc.inv_dr = sink.inv_dr;
        s_kc.c0 = sink.c0;
        s_kc.cap = (float)(sink.nbins + 1);
        s_kc.nbl2 = -(sink.band_lim * sink.band_lim);
        s_kc.wgt = SELF ? 2u : 1u;
        s_kc.stride = 4u << sink.ls_shift;
        s_next = (uint32_t)NW;
    }
    __syncthreads();
    const uint32_t nq = s_u.h.nq, nrec = s_u.h.nrec, fill = s_u.h.fill, qs0 = s_u.h.qs0, hstart = s_u.h.hstart;
    if (nq == 0u) return;                                          // nothing here (uniform)
    const uint32_t *list = lists + uidx * (uint64_t)list_cap;
    const uint32_t tile0 = smem_u32(tile);
    // the first records are on their way while the tile is staged
    uint32_t rec = 32u * (uint32_t)warp + lane < nrec ? list[32u * (uint32_t)warp + lane] : 0u;
    if (!SELF && (uint32_t)tid < nq) s_q[SELF ? 0 : tid] = q_atoms[qs0 + tid];
    // ---- stage: transpose into planes, translate x, sentinels into the padding.  A warp takes whole
    // (row, run) pieces -- up to three contiguous runs of the sorted atoms per row -- so the row and
    // the run of an element are not searched for per element (r02c line profile: 12 % of the kernel's
    // instructions were staging, a third of them the search); eight loads in flight per lane ----
    for (int rs = warp; rs < 3 * NR; rs += NW) {
        const int r = rs / 3, sxi = rs - 3 * r;
        const FineRowA fa = s_u.a[r];
        const FineRowB fb = s_u.b[r];
        const uint32_t p0 = sxi == 0 ? 0u : (sxi == 1 ? fb.end0 : fb.end1);
        const uint32_t p1 = sxi == 0 ? fb.end0 : (sxi == 1 ? fb.end1 : fb.len);
        const uint32_t g = (sxi == 0 ? fa.g0 : (sxi == 1 ? fa.g1 : fa.g2)) - p0;   // row position p -> global slot g + p
        const float xs = sxi == 1 ? 0.0f : (sxi == 0 ? -s_u.c[r].xl : s_u.c[r].xl);
        const uint32_t dst0 = tile0 + 4u * fa.base;
        for (uint32_t p = p0 + lane; p < p1; p += 256u) {
            float4 c[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (p + 32u * u < p1) c[u] = c_atoms[g + p + 32u * u];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (p + 32u * u < p1) {
                    const uint32_t dst = dst0 + 4u * (p + 32u * u);
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst), "f"(c[u].x + xs) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + CAP * 4u), "f"(c[u].y) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + 2u * CAP * 4u), "f"(c[u].z) : "memory");
                }
            }
        }
        // the padding behind the row (rows start at multiples of four): with the row's last run
        if (sxi == 2) {
            const uint32_t pp = fb.len + (uint32_t)lane;
            if (pp < ru4(fb.len) && fa.base + pp < fill) {
                const uint32_t dst = dst0 + 4u * pp;
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst), "f"(QUAD_FAR) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + CAP * 4u), "f"(QUAD_FAR) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + 2u * CAP * 4u), "f"(QUAD_FAR) : "memory");
            }
        }
    }
    if (SELF && sp.include_self && (uint32_t)tid < nq) {           // capped_distance(x, x) reports (i, i)
        const float4 q = q_atoms[qs0 + tid];
        const double d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, q.x, q.y, q.z, box);
        if (d2 <= sp.d2_max && d2 > sp.d2_min) sink.template hit_w<false>(__float_as_int(q.w), __float_as_int(q.w), d2, f);
    }
    // (constants are read back from shared memory: a kernel parameter would be re-fetched from
    // the constant bank inside the hot loop, one issue slot per use)
    FineConst kc = s_kc;
    kc.hb = sink.hist_addr() + 4u * ((uint32_t)lane & ((1u << sink.ls_shift) - 1u)) - kc.stride * 0x4B400000u;
    uint2 *slowq = s_slow[warp];
    uint32_t *nslow = &s_nslow[warp];
    // evaluates the queued pairs, 32 at a time (all lanes of the warp)
    auto drain = [&]() {
        __syncwarp();
        const uint32_t n = min(*nslow, (uint32_t)RF_SLOW);
        for (uint32_t i = lane; i < n; i += 32) {
            const uint2 e = slowq[i];
            quad_exact<SELF, Sink, MIC>(sink, &box, sp.d2_max, sp.d2_min, q_atoms, c_atoms, e.x, e.y, f);
        }
        __syncwarp();
        if (lane == 0) *nslow = 0u;
        __syncwarp();
    };
    __syncthreads();                                               // the tile is staged
    // ---- walk: warp w starts with item w; whoever finishes claims the next unclaimed one (and
    // fetches its records before it walks the current ones) ----
    const uint32_t own0 = s_u.h.own0;
    const int nitems = (int)((nrec + 31u) >> 5);
    for (int item = warp; item < nitems;) {
        uint32_t nxt = 0;
        if (lane == 0) nxt = atomicAdd(&s_next, 1u);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        const uint32_t rec_next = 32u * nxt + lane < nrec ? list[32u * nxt + lane] : 0u;
        const bool home = SELF && 32u * (uint32_t)item >= hstart;  // uniform
        const uint32_t len = rec & 127u, qa = (rec >> 7) & 1023u, pair = (rec >> 17) & 255u;
        const int r = (int)(rec >> 25);
        const FineRowC fc = s_u.c[r];
        const uint32_t qi = 2u * pair;
        const bool has_b = qi + 1u < nq;
        float3 qa3, qb3;
        if (SELF) {
            const uint32_t at = tile0 + 4u * (own0 + qi);
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qa3.x) : "r"(at));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qa3.y) : "r"(at + CAP * 4u));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qa3.z) : "r"(at + 2u * CAP * 4u));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qb3.x) : "r"(at + 4u));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qb3.y) : "r"(at + CAP * 4u + 4u));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qb3.z) : "r"(at + 2u * CAP * 4u + 4u));
        } else {
            const float4 q0 = s_q[SELF ? 0 : qi], q1 = s_q[SELF ? 0 : min(qi + 1u, nq - 1u)];
            qa3 = make_float3(q0.x, q0.y, q0.z);
            qb3 = make_float3(q1.x, q1.y, q1.z);
        }
        if (!has_b) qb3 = make_float3(-QUAD_FAR, -QUAD_FAR, -QUAD_FAR);      // nobody there: every candidate misses
        const uint32_t qslot = qs0 + qi;
        auto slow_fn = [&](int which, uint32_t quad, int k, uint32_t addr) {
            // out of the bin the FP32 path chose; float64 decides (a sentinel is never inside a band)
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(0u - kc.wgt) : "memory");
            const FineRowA fa = s_u.a[r];
            const FineRowB fb = s_u.b[r];
            const uint32_t p = 4u * quad + (uint32_t)k - fa.base;
            const uint32_t cslot = p < fb.end0 ? fa.g0 + p : (p < fb.end1 ? fa.g1 + (p - fb.end0) : fa.g2 + (p - fb.end1));
            const uint32_t at = atomicAdd(nslow, 1u);
            if (at < (uint32_t)RF_SLOW) slowq[at] = make_uint2(qslot + (uint32_t)which, cslot);
            else quad_exact<SELF, Sink, MIC>(sink, &box, sp.d2_max, sp.d2_min, q_atoms, c_atoms, qslot + (uint32_t)which, cslot, f);
        };
        const float3 na3 = make_float3(MIC == MIC_TRICLINIC ? -(qa3.x - fc.sx) : -qa3.x, -(qa3.y - fc.sy), -(qa3.z - fc.sz));
        const float3 nb3 = make_float3(MIC == MIC_TRICLINIC ? -(qb3.x - fc.sx) : -qb3.x, -(qb3.y - fc.sy), -(qb3.z - fc.sz));
        if (home)                                                  // candidates behind each query's own position only
            fine_walk2<true, CAP>(tile0, qa, qa + len, na3, nb3, (int)(own0 + qi) - (int)(4u * qa), kc, slow_fn);
        else
            fine_walk2<false, CAP>(tile0, qa, qa + len, na3, nb3, 0, kc, slow_fn);
        __syncwarp();
        if (*nslow >= (uint32_t)(RF_SLOW - 8)) drain();            // (32 at a time: an evaluation by two lanes costs as much)
        item = (int)nxt;
        rec = rec_next;
    }
    __syncwarp();
    if (*nslow) drain();
    sink.end();
}

}  // namespace mdagpu
