// rowtile.cuh -- the fast cutoff-pair traversal: one CTA per run of home cells along x,
// candidates of a whole neighbour row staged once in shared memory, image shift
// pre-applied while staging.
//
// Reference behaviour reproduced (package/MDAnalysis/lib/nsgrid.pyx):
//   search       :499-580  27-cell stencil around the query's cell
//   self_search  :582-644  home cell (later atoms) + 13-cell half shell, each unordered
//                          cell pair once
//   calc_distsq  :483-497  float32 difference, float64 minimum image
//
// Why this is exact although the inner loop is FP32:
//  * Cells are sorted x-fastest, so the (up to 2W+1) cells a query must see in a
//    neighbour row (dy, dz) are ONE contiguous run of the sorted atom array (three
//    runs when the row wraps around the periodic x edge).  The run is staged in
//    shared memory with the lattice translation of that (row, segment) already
//    added: c' = c + a*h0 + b*h1 + c*h2, a,b,c in {-1,0,1}.  The translation is a
//    property of the CELL pair, not of the atom pair: with >= 4 cells per periodic
//    dimension (the reference's own rule, nsgrid.pyx:361-367) a wrapped stencil
//    neighbour always needs exactly this image and a non-wrapped one never needs
//    any (host-checked for triclinic cells, see rowtile_supported()).
//  * The FP32 value |q - c'|^2 only FILTERS: it rejects pairs that are outside the
//    cutoff by more than a rigorous error band, and bins pairs that are inside a
//    histogram bin by more than that band.  Every pair inside a band is
//    re-evaluated from the ORIGINAL float32 coordinates with the reference's
//    float64 arithmetic (geom.cuh: pair_d2_exact / tric_image_d2), and that value
//    alone takes the decision.  Pair lists always carry the exact float64 distance.
//
// Work decomposition: blockIdx.x = (z * ny + y) * runs + run, blockIdx.y = frame.
// A thread owns one query atom of the run (cell-sorted order, so a warp mostly
// shares one home cell and all its lanes read the same staged candidate: LDS.128
// broadcast); its window inside the staged row is [cell - W, cell + W].
#pragma once
#include "engine.cuh"

namespace mdagpu {

constexpr int ROW_TILE = 1024;       // staged candidates per piece (16 kB)
constexpr int ROW_MAXR = 24;         // max home cells per CTA
constexpr int ROW_MAXC = ROW_MAXR + 4 + 1;

struct RowParams {
    int W;              // window half-width in cells (1: cells >= cutoff, 2: cells >= cutoff/2)
    int R;              // home cells per CTA along x
};

// Can the per-cell-pair image prediction be used for this frame geometry?
//  MIC_NONE: always.  Otherwise every dimension must be periodic with >= 2W+2 cells
//  (>= 4 for W = 1), and for triclinic cells the images that are NOT predicted must be
//  provably beyond the cutoff: (1 - (W+1)/n_d) * width_d > cutoff, width_d > 2 cutoff.
static inline bool rowtile_supported(int mic, const GridGeom &g, const BoxParams &b, double cutoff, int W)
{
    if (mic == MIC_NONE) return true;
    if (mic == MIC_ORTHO_ROUND) return false;
    for (int k = 0; k < 3; ++k)
        if (!g.periodic[k] || g.ncells[k] < 2 * W + 2) return false;
    if (mic == MIC_ORTHO_LAZY) {
        for (int k = 0; k < 3; ++k)
            if (!((double)b.L[k] * (1.0 - (double)(W + 1) / g.ncells[k]) >= 0.5 * (double)b.L[k])) return false;
        return true;
    }
    // triclinic: perpendicular widths of the cell
    const double h0[3] = { b.h[0], b.h[1], b.h[2] }, h1[3] = { b.h[3], b.h[4], b.h[5] }, h2[3] = { b.h[6], b.h[7], b.h[8] };
    auto cross = [](const double *u, const double *v, double *o) {
        o[0] = u[1] * v[2] - u[2] * v[1]; o[1] = u[2] * v[0] - u[0] * v[2]; o[2] = u[0] * v[1] - u[1] * v[0];
    };
    auto norm = [](const double *u) { return sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]); };
    double c12[3], c20[3], c01[3];
    cross(h1, h2, c12); cross(h2, h0, c20); cross(h0, h1, c01);
    const double vol = fabs(h0[0] * c12[0] + h0[1] * c12[1] + h0[2] * c12[2]);
    const double width[3] = { vol / norm(c12), vol / norm(c20), vol / norm(c01) };
    for (int k = 0; k < 3; ++k) {
        if (!(width[k] > 2.0 * cutoff * (1.0 + 1e-5))) return false;
        if (!((1.0 - (double)(W + 1) / g.ncells[k]) * width[k] > cutoff * (1.0 + 1e-5))) return false;
    }
    return true;
}

// Bound on |FP32 distance from staged coordinates - the reference's float64 distance|.
// Per axis k, with S_k the sum of the lattice components (box edge for orthorhombic cells):
//   translation built in FP32 (two roundings of a value <= S_k)      2 * S_k * 2^-24
//   staged coordinate c + translation rounded once (|.| <= 2 S_k)    2 * S_k * 2^-24
//   the reference's own rounding of q - c (|.| <= S_k), which is
//   part of ITS value but not of the FP32 estimate                       S_k * 2^-24
//   q - c' for a pair near the cutoff                                 cutoff * 2^-24
// i.e. (5 S_k + cutoff) * 6e-8, summed over the axes (>= the Euclidean combination).
// Without a box the FP32 estimate starts from the very same rounded difference.
static inline double rowtile_abs_err(int mic, const BoxParams &b, double cutoff)
{
    if (mic == MIC_NONE) return 0.0;
    double s = 0.0;
    for (int k = 0; k < 3; ++k) {
        const double S = b.mode == BOX_TRICLINIC
            ? fabs((double)b.h[k]) + fabs((double)b.h[3 + k]) + fabs((double)b.h[6 + k]) : fabs((double)b.L[k]);
        s += (5.0 * S + cutoff) * 6e-8;
    }
    return s;
}

template <int MIC, bool SELF, class Sink>
__global__ void __launch_bounds__(Sink::THREADS)
row_tile_kernel(const float4 *__restrict__ q_atoms, const uint32_t *__restrict__ q_start,
                const float4 *__restrict__ c_atoms, const uint32_t *__restrict__ c_start,
                const FrameGeom *__restrict__ fgs, int per_frame, RowParams rp, SearchParams sp, Sink sink)
{
    constexpr int T = Sink::THREADS;
    __shared__ FrameGeom sfg;
    __shared__ float4 tile[ROW_TILE];
    __shared__ uint32_t s_off[ROW_MAXC + 1];      // tile offset of every staged cell (virtual x order)
    __shared__ uint32_t s_qst[ROW_MAXR + 1];      // first query slot of every home cell of the run
    const uint64_t f = blockIdx.y;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(fgs + (per_frame ? f : 0));
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sfg);
        for (int w = threadIdx.x; w < (int)(sizeof(FrameGeom) / 4); w += T) dst[w] = src[w];
    }
    sink.begin();
    __syncthreads();
    const int W = rp.W, R = rp.R;
    const int nx = sfg.grid.ncells[0], ny = sfg.grid.ncells[1], nz = sfg.grid.ncells[2];
    const int runs = (nx + R - 1) / R;
    if (blockIdx.x >= (unsigned)(runs * ny * nz)) return;          // frames may have fewer cells (uniform)
    const int xr = blockIdx.x % runs, y = (blockIdx.x / runs) % ny, z = blockIdx.x / (runs * ny);
    const int x0 = xr * R, x1 = min(x0 + R, nx);
    const int64_t base = per_frame ? sfg.cell_base : (int64_t)f * sfg.grid.ncell_total;
    const BoxParams box = sfg.box;
    // lattice rows as FP32 vectors (orthorhombic: the box edges)
    float hx[3], hy[3], hz[3];
    if (MIC == MIC_TRICLINIC) {
        hx[0] = box.h[0]; hx[1] = box.h[1]; hx[2] = box.h[2];
        hy[0] = box.h[3]; hy[1] = box.h[4]; hy[2] = box.h[5];
        hz[0] = box.h[6]; hz[1] = box.h[7]; hz[2] = box.h[8];
    } else {
        hx[0] = box.L[0]; hx[1] = 0.f; hx[2] = 0.f;
        hy[0] = 0.f; hy[1] = box.L[1]; hy[2] = 0.f;
        hz[0] = 0.f; hz[1] = 0.f; hz[2] = box.L[2];
    }
    const int64_t home_row = base + ((int64_t)z * ny + y) * nx;
    for (int j = threadIdx.x; j <= x1 - x0; j += T) s_qst[j] = q_start[home_row + x0 + j];
    __syncthreads();
    const uint32_t qs0 = s_qst[0], qs1 = s_qst[x1 - x0];

    for (uint32_t qbase = qs0; qbase < qs1; qbase += T) {
        const uint32_t qslot = qbase + threadIdx.x;
        const bool active = qslot < qs1;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        int cxl = 0;                                           // home cell of the query, relative to x0
        if (active) {
            q = q_atoms[qslot];
            while (cxl + 1 < x1 - x0 && s_qst[cxl + 1] <= qslot) ++cxl;
        }
        const int qi = __float_as_int(q.w);
        if (SELF && sp.include_self && active) {               // capped_distance(x, x) reports (i, i)
            const double d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, q.x, q.y, q.z, box);
            if (d2 <= sp.d2_max && d2 > sp.d2_min) sink.template hit_w<false>(qi, qi, d2, f);
        }
        // rows of the stencil: SELF visits the half shell {dz > 0} U {dz = 0, dy > 0} U home row
        for (int dz = SELF ? 0 : -W; dz <= W; ++dz) {
            int zz = z + dz, ic = 0;
            if (zz < 0) { if (!sfg.grid.periodic[2]) continue; zz += nz; ic = -1; }
            else if (zz >= nz) { if (!sfg.grid.periodic[2]) continue; zz -= nz; ic = 1; }
            for (int dy = (SELF && dz == 0) ? 0 : -W; dy <= W; ++dy) {
                int yy = y + dy, ib = 0;
                if (yy < 0) { if (!sfg.grid.periodic[1]) continue; yy += ny; ib = -1; }
                else if (yy >= ny) { if (!sfg.grid.periodic[1]) continue; yy -= ny; ib = 1; }
                const bool home = SELF && dz == 0 && dy == 0;
                const int64_t row = base + ((int64_t)zz * ny + yy) * nx;
                // virtual cells vlo .. vhi (inclusive) of this row are staged
                const int vlo = home ? x0 : x0 - W, vhi = x1 - 1 + W;
                const int ncell = vhi - vlo + 1;
                __syncthreads();                               // previous row fully consumed
                if (threadIdx.x < 32) {                        // warp 0: offsets of the staged cells
                    uint32_t run = 0;
                    for (int j0 = 0; j0 < ncell; j0 += 32) {
                        const int j = j0 + (int)threadIdx.x;
                        uint32_t cnt = 0;
                        if (j < ncell) {
                            int v = vlo + j;
                            bool ok = true;
                            if (v < 0) { ok = sfg.grid.periodic[0]; v += nx; }
                            else if (v >= nx) { ok = sfg.grid.periodic[0]; v -= nx; }
                            if (ok) cnt = c_start[row + v + 1] - c_start[row + v];
                        }
                        uint32_t incl = cnt;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                            if ((int)threadIdx.x >= o) incl += u;
                        }
                        if (j < ncell) s_off[j] = run + incl - cnt;
                        run += __shfl_sync(0xffffffffu, incl, 31);
                    }
                    if (threadIdx.x == 0) s_off[ncell] = run;
                }
                __syncthreads();
                const uint32_t total = s_off[ncell];
                // the three contiguous segments of the row: left wrap (image -h0), main, right wrap (+h0)
                const int nleft = vlo < 0 ? -vlo : 0, nright = vhi >= nx ? vhi - nx + 1 : 0;
                const uint32_t left_end = s_off[nleft], right_begin = s_off[ncell - nright];
                const uint32_t g_left = (nleft && sfg.grid.periodic[0]) ? c_start[row + nx - nleft] : 0u;
                const uint32_t g_main = c_start[row + (vlo < 0 ? 0 : vlo)];
                const uint32_t g_right = c_start[row];
                // translation of the row (b, c) and of the x segments (a)
                const float ryx = (float)ib * hy[0] + (float)ic * hz[0];
                const float ryy = (float)ib * hy[1] + (float)ic * hz[1];
                const float ryz = (float)ib * hy[2] + (float)ic * hz[2];
                // window of this thread in tile coordinates
                uint32_t lo = 0, hi = 0;
                if (active) {
                    const int jl = home ? cxl : cxl;           // first staged cell of the window
                    lo = s_off[jl];
                    hi = s_off[min(jl + (home ? W : 2 * W) + 1, ncell)];
                    if (home) lo = s_off[cxl] + (qslot - s_qst[cxl]) + 1;      // later atoms of the own cell
                }
                for (uint32_t p0 = 0; p0 < total; p0 += ROW_TILE) {
                    const uint32_t pn = min((uint32_t)ROW_TILE, total - p0);
                    if (p0) __syncthreads();
                    for (uint32_t k = threadIdx.x; k < pn; k += T) {
                        const uint32_t t = p0 + k;
                        uint32_t slot;
                        float sx = ryx, sy = ryy, sz = ryz;
                        if (t < left_end) { slot = g_left + t; sx -= hx[0]; sy -= hx[1]; sz -= hx[2]; }
                        else if (t < right_begin) slot = g_main + (t - left_end);
                        else { slot = g_right + (t - right_begin); sx += hx[0]; sy += hx[1]; sz += hx[2]; }
                        const float4 c = c_atoms[slot];
                        tile[k] = make_float4(c.x + sx, c.y + sy, c.z + sz, c.w);
                    }
                    __syncthreads();
                    if (!active) continue;
                    const uint32_t a = max(lo, p0), b = min(hi, p0 + pn);
#pragma unroll 4
                    for (uint32_t t = a; t < b; ++t) {
                        const float4 c = tile[t - p0];
                        const float tx = q.x - c.x, ty = q.y - c.y, tz = q.z - c.z;
                        const float d2f = fmaf(tx, tx, fmaf(ty, ty, tz * tz));
                        if (d2f > sp.d2_filter) continue;
                        if (Sink::FAST && sink.template try_fast_w<SELF>(qi, __float_as_int(c.w), d2f)) continue;
                        // exact re-evaluation from the original coordinates
                        const uint32_t slot = t < left_end ? g_left + t
                                            : (t < right_begin ? g_main + (t - left_end) : g_right + (t - right_begin));
                        const float4 co = c_atoms[slot];
                        double d2;
                        if (MIC == MIC_TRICLINIC) {
                            const int ia = t < left_end ? -1 : (t >= right_begin ? 1 : 0);
                            D3 d = sep_f32(q.x, q.y, q.z, co.x, co.y, co.z), img;
                            d2 = tric_image_d2(d, box.h, -ia, -ib, -ic, img);
                        } else {
                            d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, co.x, co.y, co.z, box);
                        }
                        if (d2 <= sp.d2_max && d2 > sp.d2_min) sink.template hit_w<SELF>(qi, __float_as_int(co.w), d2, f);
                    }
                }
            }
        }
    }
    sink.end();
}

}  // namespace mdagpu
