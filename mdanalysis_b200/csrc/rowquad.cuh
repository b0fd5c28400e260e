// rowquad.cuh -- the InterRDF pair histogram for orthorhombic cells (the headline kernel).
//
// Same decomposition and the same exactness argument as rowtile.cuh: one CTA per run of
// home cells along x, one thread per query, the FP32 value only bins pairs that are
// provably away from every decision boundary, everything else is re-evaluated in float64
// from the original coordinates.  Reference: package/MDAnalysis/analysis/rdf.py:268-291
// (capped_distance + np.histogram) and lib/nsgrid.pyx:483-644.  What is different:
//
//  * ALL stencil rows of the CTA are staged ONCE (13 rows for a self RDF with the 5x5x5
//    stencil: 15-41 kB), by 4-byte cp.async copies issued by all threads; after one
//    __syncthreads every lane walks its windows in all rows with no further synchronisation.
//    (Row-by-row stages cost a barrier each: 27 % of the warp samples sat there.  Rows that
//    do not fit the tile -- dense systems -- are staged in groups of rows / pieces of a row.)
//    A 224-thread CTA takes a whole row of home cells when its queries fit, 128-thread CTAs
//    take runs of R cells otherwise (rdf.cu).
//  * The copies TRANSPOSE: x[], y[], z[] planes, every run of the row (x-wrap left / in
//    range / x-wrap right) padded to a multiple of four with far-away sentinels.  One
//    LDS.128 per plane feeds four candidates, and two candidates share every arithmetic
//    instruction through the packed FP32 pipe of sm_100 (FADD2 / FMUL2 / FFMA2: two
//    results per issue slot) -- the kernel is issue bound, not pipe bound.
//  * A lane walks its own window rounded out to whole quads.  The extra candidates are
//    harmless: an atom outside the lane's cells [cell - W, cell + W] is farther than the
//    cutoff under the image of its run (cells are >= cutoff / W * (1 + 1e-4) + 1e-3 wide,
//    rdf.cu: grid_geometry_ortho_free) and fails the range test like any distant atom.
//    (Not so on the REFERENCE's grid for triclinic cells, where its own stencil can be
//    incomplete and the window is part of the result: those stay on rowtile.cuh.  On the
//    library's own triclinic grid -- rdf.cu: grid_geometry_triclinic_free, taken only where
//    the reference's stencil is complete -- the same argument holds and this kernel runs with
//    MIC = MIC_TRICLINIC: the row's translation is ib b + ic c, all three components.)
//  * The cutoff is the top histogram edge (InterRDF passes range[1] as max_cutoff,
//    rdf.py:270-272), so "inside the cutoff" is "bin coordinate below nbins" and the band
//    around the cutoff is the band around the last edge: one range test, one band test.
//  * ptxas never predicates a shared-memory atomic (it branches around it, three extra
//    issue slots per candidate): lanes without a hit add to a private dummy word instead.
//  * Pairs inside a decision band are not evaluated where they are found (one lane busy,
//    31 waiting, ~250 issue slots): they are queued per warp and evaluated 32 at a time.
#pragma once
#include "rowtile.cuh"

namespace mdagpu {

constexpr int WQ_CAP_128 = 1280;          // padded candidates the tile holds: 128-thread CTAs (15 kB, 7 CTAs/SM)
constexpr int WQ_CAP_224 = 3584;          // 224-thread CTAs own a whole row of home cells (42 kB: the 13 rows of C2, ~3380 padded candidates, fit with 3.5 sigma to spare)
constexpr int WQ_SLOW = 48;               // queued exact evaluations per warp
constexpr int WQ_PARTS = 2;               // row parts a chunk of queries is split into (work items of the walk; 3: 32.4 ms, 4: 32.7 ms, 2: 32.0 ms)
constexpr float QUAD_FAR = 1.0e18f;       // sentinel coordinate: (2e18)^2 * 3 is finite in FP32

__device__ __forceinline__ uint32_t ru4(uint32_t v) { return (v + 3u) & ~3u; }

// Bound on |FP32 distance of this kernel - the reference's float64 distance| for a pair near
// a decision boundary.  Per axis, with L the box edge (coordinates are wrapped into [0, L)):
//   q' = fl(q - t), t = 0 or +-L exactly: |q'| <= 2 L                       2 L * 2^-24
//   fl(c - q') for a pair within the cutoff                              cutoff * 2^-24
//   the reference's own fl(q - c), |q - c| < L, part of ITS value             L * 2^-24
// and the three axes combine as a Euclidean norm (|d(v) - d(v')| <= |v - v'|).
static inline double rowquad_abs_err(const BoxParams &b, double cutoff)
{
    double s = 0.0;
    for (int k = 0; k < 3; ++k) {
        // triclinic cells: S = sum of the lattice components along the axis bounds a wrapped
        // coordinate; t = fl(ib b + ic c) (S), q' = fl(q - t) (2 S), fl(q' +- a) (3 S), the
        // reference's own fl(q - c) (S): 7 S + cutoff, rounded up
        const double e = b.mode == BOX_TRICLINIC
            ? (8.0 * (fabs((double)b.h[k]) + fabs((double)b.h[3 + k]) + fabs((double)b.h[6 + k])) + 2.0 * cutoff) * 6e-8
            : (3.0 * fabs((double)b.L[k]) + 2.0 * cutoff) * 6e-8;
        s += e * e;
    }
    return sqrt(s);
}

struct QuadConst {
    float lim, band_lim, inv_dr, c0;
    uint32_t hb, trash, wgt;
};

// Padded layout of a row: run s starts at P[s] (multiple of 4), P[3] = padded length.
struct QuadPad {
    uint32_t P[4];
};
__device__ __forceinline__ QuadPad quad_pad(const RowInfo &ri)
{
    QuadPad qp;
    qp.P[0] = 0u;
    qp.P[1] = ru4(ri.end[0]);
    qp.P[2] = qp.P[1] + ru4(ri.end[1] - ri.end[0]);
    qp.P[3] = qp.P[2] + ru4(ri.end[2] - ri.end[1]);
    return qp;
}

// Walks quads [qa, qe) of one staged run for one query.  `planes` is the shared-memory
// address of the x plane of the buffer (y and z follow at CAP floats each).  HOME: only candidates whose tile position
// is behind the query's own (`ahead` = own position - position of quad qa).  Pairs inside
// a decision band are handed to slow_fn(quad, k).
template <bool HOME, int CAP, class SlowFn>
__device__ __forceinline__ void quad_walk(uint32_t planes, uint32_t qa, uint32_t qe,
                                          float nqx, float nqy, float nqz, int ahead,
                                          const QuadConst &kc, SlowFn slow_fn)
{
    constexpr uint32_t PL = CAP * 4;              // bytes between the planes
    const float2 nqx2 = make_float2(nqx, nqx), nqy2 = make_float2(nqy, nqy), nqz2 = make_float2(nqz, nqz);
    const float2 inv2 = make_float2(kc.inv_dr, kc.inv_dr), c02 = make_float2(kc.c0, kc.c0);
    const float2 magic2 = make_float2(12582912.0f, 12582912.0f), nmagic2 = make_float2(-12582912.0f, -12582912.0f);
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    const uint32_t end = planes + 16u * qe;
    for (uint32_t ptr = planes + 16u * qa; ptr < end; ptr += 16u, ahead -= 4) {
        float4 cx, cy, cz;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cx.x), "=f"(cx.y), "=f"(cx.z), "=f"(cx.w) : "r"(ptr));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cy.x), "=f"(cy.y), "=f"(cy.z), "=f"(cy.w) : "r"(ptr + PL));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cz.x), "=f"(cz.y), "=f"(cz.z), "=f"(cz.w) : "r"(ptr + 2u * PL));
        float2 um[2], tt[2], fr[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 x2 = h ? make_float2(cx.z, cx.w) : make_float2(cx.x, cx.y);
            const float2 y2 = h ? make_float2(cy.z, cy.w) : make_float2(cy.x, cy.y);
            const float2 z2 = h ? make_float2(cz.z, cz.w) : make_float2(cz.x, cz.y);
            const float2 dx = __fadd2_rn(x2, nqx2), dy = __fadd2_rn(y2, nqy2), dz = __fadd2_rn(z2, nqz2);
            const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __fmul2_rn(dz, dz)));
            float2 d;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d.x) : "f"(d2.x));
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d.y) : "f"(d2.y));
            // bin coordinate um = d * inv_dr + c0: bin k <=> um in [k - 0.5, k + 0.5); the magic
            // number add rounds to the bin (its low mantissa bits), fr = um - bin in [-0.5, 0.5]
            um[h] = __ffma2_rn(d, inv2, c02);
            tt[h] = __fadd2_rn(um[h], magic2);
            fr[h] = __ffma2_rn(__fadd2_rn(tt[h], nmagic2), neg1, um[h]);
        }
        const float u[4] = { um[0].x, um[0].y, um[1].x, um[1].y };
        const float t4[4] = { tt[0].x, tt[0].y, tt[1].x, tt[1].y };
        const float f4[4] = { fr[0].x, fr[0].y, fr[1].x, fr[1].y };
        // Decision + histogram update of the four candidates, in PTX so that each costs the
        // compares, one select and the reduction (see the file header for the dummy word).
        //   in   = um < lim  [and, HOME, behind the query]       inside the cutoff (+ band)
        //   keep = in and |fr| < band_lim                         safely inside a bin: count it
        //   slow = in and not keep                                float64 decides
        uint32_t any_slow;
#define QUAD_DECIDE(HOME_TESTS, AHEAD)                                                             \
        asm volatile(                                                                              \
            "{\n\t"                                                                                \
            ".reg .pred i0, i1, i2, i3, k0, k1, k2, k3, s0, s1, s2, s3;\n\t"                        \
            ".reg .f32 g0, g1, g2, g3;\n\t"                                                         \
            ".reg .b32 a0, a1, a2, a3;\n\t"                                                         \
            "setp.lt.f32 i0, %1, %9;\n\t"                                                           \
            "setp.lt.f32 i1, %2, %9;\n\t"                                                           \
            "setp.lt.f32 i2, %3, %9;\n\t"                                                           \
            "setp.lt.f32 i3, %4, %9;\n\t" HOME_TESTS                                                \
            "abs.f32 g0, %5;\n\t"                                                                   \
            "abs.f32 g1, %6;\n\t"                                                                   \
            "abs.f32 g2, %7;\n\t"                                                                   \
            "abs.f32 g3, %8;\n\t"                                                                   \
            "setp.lt.and.f32 k0, g0, %10, i0;\n\t"                                                  \
            "setp.lt.and.f32 k1, g1, %10, i1;\n\t"                                                  \
            "setp.lt.and.f32 k2, g2, %10, i2;\n\t"                                                  \
            "setp.lt.and.f32 k3, g3, %10, i3;\n\t"                                                  \
            "xor.pred s0, i0, k0;\n\t"                                                              \
            "xor.pred s1, i1, k1;\n\t"                                                              \
            "xor.pred s2, i2, k2;\n\t"                                                              \
            "xor.pred s3, i3, k3;\n\t"                                                              \
            "selp.b32 a0, %11, %16, k0;\n\t"                                                        \
            "selp.b32 a1, %12, %16, k1;\n\t"                                                        \
            "selp.b32 a2, %13, %16, k2;\n\t"                                                        \
            "selp.b32 a3, %14, %16, k3;\n\t"                                                        \
            "red.shared.add.u32 [a0], %17;\n\t"                                                     \
            "red.shared.add.u32 [a1], %17;\n\t"                                                     \
            "red.shared.add.u32 [a2], %17;\n\t"                                                     \
            "red.shared.add.u32 [a3], %17;\n\t"                                                     \
            "or.pred s0, s0, s1;\n\t"                                                               \
            "or.pred s2, s2, s3;\n\t"                                                               \
            "or.pred s0, s0, s2;\n\t"                                                               \
            "selp.u32 %0, 1, 0, s0;\n\t"                                                            \
            "}"                                                                                    \
            : "=r"(any_slow)                                                                       \
            : "f"(u[0]), "f"(u[1]), "f"(u[2]), "f"(u[3]), "f"(f4[0]), "f"(f4[1]), "f"(f4[2]), "f"(f4[3]), \
              "f"(kc.lim), "f"(kc.band_lim),                                                       \
              "r"(kc.hb + 4u * __float_as_uint(t4[0])), "r"(kc.hb + 4u * __float_as_uint(t4[1])),  \
              "r"(kc.hb + 4u * __float_as_uint(t4[2])), "r"(kc.hb + 4u * __float_as_uint(t4[3])),  \
              "r"(AHEAD), "r"(kc.trash), "r"(kc.wgt))
        if constexpr (HOME) {
            QUAD_DECIDE("setp.lt.and.s32 i0, %15, 0, i0;\n\t"
                        "setp.lt.and.s32 i1, %15, 1, i1;\n\t"
                        "setp.lt.and.s32 i2, %15, 2, i2;\n\t"
                        "setp.lt.and.s32 i3, %15, 3, i3;\n\t", ahead);
        } else {
            QUAD_DECIDE("", 0);
        }
#undef QUAD_DECIDE
        if (any_slow) {
            const uint32_t quad = (ptr - planes) >> 4;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (u[k] < kc.lim && (!HOME || k > ahead) && !(fabsf(f4[k]) < kc.band_lim)) slow_fn(quad, k);
        }
    }
}

// Exact evaluation of one queued pair (slots in the sorted arrays) in the reference's
// float64 arithmetic; the histogram sink bins it against the exact thresholds.
template <bool SELF, class Sink, int MIC = MIC_ORTHO_LAZY>
__device__ __noinline__ void quad_exact(Sink sink, const BoxParams *box, double d2_max, double d2_min,
                                        const float4 *__restrict__ q_atoms, const float4 *__restrict__ c_atoms,
                                        uint32_t qslot, uint32_t cslot, uint64_t f)
{
    const float4 q = q_atoms[qslot], c = c_atoms[cslot];
    const double d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, c.x, c.y, c.z, *box);
    if (d2 <= d2_max && d2 > d2_min) sink.template hit_w<SELF>(__float_as_int(q.w), __float_as_int(c.w), d2, f);
}

__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

template <bool SELF, class Sink, int CAP, int MIN_CTAS, int MIC_ = MIC_ORTHO_LAZY>
__global__ void __launch_bounds__(Sink::THREADS, MIN_CTAS)
row_quad_kernel(const float4 *__restrict__ q_atoms, const uint32_t *__restrict__ q_start,
                const float4 *__restrict__ c_atoms, const uint32_t *__restrict__ c_start,
                const FrameGeom *__restrict__ fgs, int per_frame, RowParams rp, SearchParams sp, Sink sink)
{
    constexpr int T = Sink::THREADS, NW = T / 32;
    constexpr int MIC = MIC_;
    static_assert(MIC == MIC_ORTHO_LAZY || MIC == MIC_TRICLINIC, "cells with a periodic image per cell pair");
    __shared__ RowBook bk;
    // x, y, z planes of the staged rows: dynamic shared memory behind the per-warp histograms
    extern __shared__ __align__(16) unsigned char quad_dyn[];
    float *tile = reinterpret_cast<float *>(quad_dyn + (((size_t)sink.nbins * NW * sizeof(uint32_t) + 15) & ~(size_t)15));
    __shared__ uint32_t s_base[ROW_MAXROWS + 1];                   // padded tile coordinate where each row starts
    __shared__ uint4 s_pad[ROW_MAXROWS];                           // {P1, P2, P3, -}: padded starts of runs 1, 2 and the row's end
    __shared__ float4 s_shift[ROW_MAXROWS];                        // the row's lattice translation (x is 0 for orthorhombic cells)
    __shared__ uint2 s_slow[NW][WQ_SLOW];                          // queued (query slot, candidate slot)
    __shared__ uint32_t s_nslow[NW];
    __shared__ uint32_t s_trash[T];
    __shared__ uint32_t s_next;                                    // next unclaimed work item of the group (see the walk)
    __shared__ QuadConst s_kc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t f = blockIdx.y;
    sink.begin();
    // (constants are read back from shared memory: a kernel parameter would be re-fetched from
    // the constant bank inside the hot loop, one issue slot per use)
    if (tid == 0) {
        s_kc.lim = (float)sink.nbins - sink.band_lim;              // see quad_walk: um < nbins - 0.5 + band
        s_kc.band_lim = sink.band_lim;
        s_kc.inv_dr = sink.inv_dr;
        s_kc.c0 = sink.c0;
        s_kc.wgt = SELF ? 2u : 1u;
    }
    if (lane == 0) s_nslow[warp] = 0u;
    const RowJob job = row_prepare<SELF, T>(bk, q_start, c_start, fgs, per_frame, rp, f);
    if (job.empty) return;
    const int W = job.W, nh = job.nh, NC = job.NC, nrows = job.nrows;
    const BoxParams &box = bk.fg.box;
    const float Lx = MIC == MIC_TRICLINIC ? box.h[0] : box.L[0];   // a = (a_x, 0, 0) in either case
    const uint32_t qs0 = bk.qst[0], qs1 = bk.qst[nh];
    for (int r = tid; r < nrows; r += T) {
        const QuadPad qp = quad_pad(bk.row[r]);
        s_pad[r] = make_uint4(qp.P[1], qp.P[2], qp.P[3], 0u);
        const float fb = (float)bk.row[r].ib, fc = (float)bk.row[r].ic;
        if (MIC == MIC_TRICLINIC)                                  // ib b + ic c
            s_shift[r] = make_float4(fb * box.h[3] + fc * box.h[6], fb * box.h[4] + fc * box.h[7], fb * box.h[5] + fc * box.h[8], 0.0f);
        else
            s_shift[r] = make_float4(0.0f, fb * box.L[1], fc * box.L[2], 0.0f);
    }
    __syncthreads();
    QuadConst kc = s_kc;
    kc.hb = sink.hb;
    kc.trash = smem_u32(&s_trash[tid]);
    uint2 *slowq = s_slow[warp];
    uint32_t *nslow = &s_nslow[warp];
    const uint32_t tile0 = smem_u32(tile);

    // evaluates the queued pairs, 32 at a time (all lanes of the warp)
    auto drain = [&]() {
        __syncwarp();
        const uint32_t n = min(*nslow, (uint32_t)WQ_SLOW);
        for (uint32_t i = lane; i < n; i += 32) {
            const uint2 e = slowq[i];
            quad_exact<SELF, Sink, MIC>(sink, &box, sp.d2_max, sp.d2_min, q_atoms, c_atoms, e.x, e.y, f);
        }
        __syncwarp();
        if (lane == 0) *nslow = 0u;
        __syncwarp();
    };

    // Rows are staged in groups [r0, r1) that fit the tile (all of them at once for a
    // liquid-like density); a row longer than the tile is staged in pieces [c0, c0 + CAP).
    int r0 = 0;
    uint32_t c0 = 0;
    while (r0 < nrows) {
        // ---- group layout (every thread computes the same) ----
        int r1 = r0;
        uint32_t fill = 0, piece = 0;                              // piece != 0: only [c0, c0 + piece) of row r0
        {
            const uint32_t len0 = s_pad[r0].z - c0;
            if (len0 > (uint32_t)CAP) { piece = CAP; fill = CAP; r1 = r0 + 1; }
            else {
                piece = c0 ? len0 : 0u;
                fill = len0; r1 = r0 + 1;
                while (r1 < nrows && !c0) {
                    const uint32_t len = s_pad[r1].z;
                    if (fill + len > (uint32_t)CAP) break;
                    fill += len; ++r1;
                }
            }
        }
        if (tid == 0) {
            uint32_t acc = 0;
            for (int r = r0; r < r1; ++r) { s_base[r] = acc; acc += (r == r0 && piece) ? piece : s_pad[r].z; }
        }
        __syncthreads();                                           // previous group consumed, s_base visible
        if (tid == 0) s_next = (uint32_t)NW;                       // nobody claims items of the previous group any more; visible after the barrier below
        // ---- stage: transposing 4-byte async copies, sentinels into the padding ----
        for (int r = r0; r < r1; ++r) {
            const RowInfo &ri = bk.row[r];
            const uint4 pd = s_pad[r];
            const uint32_t lo = r == r0 ? c0 : 0u;
            const uint32_t hi = (r == r0 && piece) ? c0 + piece : pd.z;
            const uint32_t dst0 = tile0 + 4u * s_base[r] - 4u * lo;
            // run by run: padded positions [P, P + ru4(n)) clipped to [lo, hi)
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
                const uint32_t n = ri.end[s] - (s ? ri.end[s - 1] : 0u);
                if (n == 0u) continue;
                const uint32_t P = s == 0 ? 0u : (s == 1 ? pd.x : pd.y);
                const uint32_t k0 = lo > P ? lo - P : 0u;
                const uint32_t k1 = min(ru4(n), hi > P ? hi - P : 0u);
                const uint32_t g = ri.g[s];
                for (uint32_t k = k0 + tid; k < k1; k += T) {
                    const uint32_t dst = dst0 + 4u * (P + k);
                    if (k < n) {
                        const float *src = reinterpret_cast<const float *>(c_atoms) + 4ull * (g + k);
                        cp_async4(dst, src);
                        cp_async4(dst + CAP * 4u, src + 1);
                        cp_async4(dst + 2u * CAP * 4u, src + 2);
                    } else {
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst), "f"(QUAD_FAR) : "memory");
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + CAP * 4u), "f"(QUAD_FAR) : "memory");
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst + 2u * CAP * 4u), "f"(QUAD_FAR) : "memory");
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- every lane walks its windows in the staged rows; no synchronisation inside ----
        // Work items = (chunk of 32 queries, part of the group's rows), part-major.  Warp w starts with
        // item w; whoever finishes takes the next unclaimed item, so a warp whose queries had short
        // windows takes over the second row part of somebody else's chunk instead of waiting at the
        // barrier below (ncu: 14 % of the warp samples sat there with one static chunk per warp).
        const int nchunk = (int)((qs1 - qs0 + 31u) >> 5);
        const int nparts = (r1 - r0 >= 2 * WQ_PARTS) ? WQ_PARTS : 1;
        const int nitems = nchunk * nparts;
        for (int item = warp; item < nitems;) {
            const int part = item / nchunk, chunk = item - part * nchunk;
            const int ra = r0 + (r1 - r0) * part / nparts, rb = r0 + (r1 - r0) * (part + 1) / nparts;
            const uint32_t qslot = qs0 + 32u * (uint32_t)chunk + lane;
            if (qslot < qs1) {
                const float4 q = q_atoms[qslot];
                int cxl = 0;                                       // home cell of the query, relative to x0
                while (cxl + 1 < nh && bk.qst[cxl + 1] <= qslot) ++cxl;
                if (SELF && sp.include_self && ra == 0 && c0 == 0) {     // capped_distance(x, x) reports (i, i)
                    const double d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, q.x, q.y, q.z, box);
                    if (d2 <= sp.d2_max && d2 > sp.d2_min) sink.template hit_w<false>(__float_as_int(q.w), __float_as_int(q.w), d2, f);
                }
                const uint32_t mypos = qslot - qs0;                // position in the home row
#pragma unroll 1
                for (int r = ra; r < rb; ++r) {
                    const RowInfo &ri = bk.row[r];
                    const uint4 pd = s_pad[r];
                    const bool home = SELF && r == 0;
                    const int ncell = home ? nh + W : NC;
                    const uint32_t lo = bk.off[r][cxl];
                    const uint32_t hi = bk.off[r][min(cxl + (home ? W : 2 * W) + 1, ncell)];
                    // staged part of the row in padded coordinates, and where it sits in the tile
                    const uint32_t p_lo = r == r0 ? c0 : 0u;
                    const uint32_t p_hi = (r == r0 && piece) ? c0 + piece : pd.z;
                    const uint32_t planes = tile0 + 4u * s_base[r];
                    // the row's lattice translation goes onto the query: q - (c + t) = (q - t) - c
                    const float4 sh = s_shift[r];
                    const float qxr = MIC == MIC_TRICLINIC ? q.x - sh.x : q.x;
                    const float qy = q.y - sh.y, qz = q.z - sh.z;
#pragma unroll 1
                    for (int s = 0; s < 3; ++s) {
                        const uint32_t sb = s ? ri.end[s - 1] : 0u, se = ri.end[s];
                        const uint32_t a = max(lo, sb), e = min(hi, se);
                        if (a >= e) continue;
                        const uint32_t P = s == 0 ? 0u : (s == 1 ? pd.x : pd.y);
                        const uint32_t shift = P - sb;             // row coordinate -> padded coordinate
                        const uint32_t ta = max(a + shift, p_lo), te = min(e + shift, p_hi);
                        if (ta >= te) continue;
                        const uint32_t qa = (ta - p_lo) >> 2, qe = (te - p_lo + 3u) >> 2;
                        const float qx = qxr + (float)(1 - s) * Lx; // left run: image -a, i.e. q + a
                        // a pair inside a band: padded -> global slot (a sentinel is never inside one)
                        const uint32_t g0 = ri.g[s] + p_lo - P;
                        auto slow_fn = [&](uint32_t quad, int k) {
                            const uint32_t cslot = g0 + 4u * quad + (uint32_t)k;
                            const uint32_t at = atomicAdd(nslow, 1u);
                            if (at < (uint32_t)WQ_SLOW) slowq[at] = make_uint2(qslot, cslot);
                            else quad_exact<SELF, Sink, MIC>(sink, &box, sp.d2_max, sp.d2_min, q_atoms, c_atoms, qslot, cslot, f);
                        };
                        // home row, in-range run: only candidates behind the query itself
                        // (padded == row coordinate there: the left run of a home row is empty)
                        if (home && s == 1)
                            quad_walk<true, CAP>(planes, qa, qe, -qx, -qy, -qz, (int)mypos - (int)(p_lo + 4u * qa), kc, slow_fn);
                        else
                            quad_walk<false, CAP>(planes, qa, qe, -qx, -qy, -qz, 0, kc, slow_fn);
                    }
                }
            }
            drain();
            uint32_t nxt = 0;
            if (lane == 0) nxt = atomicAdd(&s_next, 1u);
            item = (int)__shfl_sync(0xffffffffu, nxt, 0);
        }
        // next group
        if (piece) {
            c0 += piece;
            if (c0 >= s_pad[r0].z) { c0 = 0; ++r0; }
        } else {
            r0 = r1;
        }
    }
    sink.end();
}

}  // namespace mdagpu
