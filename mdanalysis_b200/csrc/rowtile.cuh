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
//    runs when the row wraps around the periodic x edge).  The lattice translation
//    of a (row, run) is a property of the CELL pair, not of the atom pair: with
//    >= 2W+2 cells per periodic dimension a wrapped stencil neighbour always needs
//    exactly this image and a non-wrapped one never needs any (host-checked for
//    triclinic cells, see rowtile_supported()).  It is applied to the QUERY once per
//    row / run:  q - (c + t) = (q - t) - c,  so the staged candidates are a verbatim
//    copy of global memory.
//  * The FP32 value |(q - t) - c|^2 only FILTERS: it rejects pairs that are outside
//    the cutoff by more than a rigorous error band, and bins pairs that are inside a
//    histogram bin by more than that band.  Every pair inside a band is re-evaluated
//    from the ORIGINAL float32 coordinates with the reference's float64 arithmetic
//    (geom.cuh: pair_d2_exact / tric_image_d2), and that value alone takes the
//    decision.  Pair lists always carry the exact float64 distance.
//
// Work decomposition: blockIdx.x = (z * ny + y) * runs + run, blockIdx.y = frame.
// A thread owns one query atom of the run (cell-sorted order); its window inside a
// staged row is the cells [cell - W, cell + W].  Rows are streamed through a
// two-buffer shared-memory ring by 1-D bulk copies (TMA, cp.async.bulk) that one
// thread issues one row ahead; an mbarrier per buffer signals arrival, so the copy
// of row r+1 runs under the arithmetic of row r and costs no issue slots.
#pragma once
#include "engine.cuh"
#include "tma.cuh"

namespace mdagpu {

constexpr int ROW_TILE = 512;        // staged candidates per buffer (8 kB, two buffers)
constexpr int ROW_MAXR = 24;         // max home cells per CTA
constexpr int ROW_MAXC = ROW_MAXR + 4;     // staged cells per row: R + 2W, W <= 2
constexpr int ROW_MAXROWS = 25;      // (2W+1)^2

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

// One stencil row as staged in shared memory: up to three contiguous runs of the sorted
// atom array (x-wrap on the left with image -h0, the in-range cells, x-wrap on the right
// with image +h0), laid out back to back in "tile coordinates".
struct RowInfo {
    uint32_t g[3];      // first global slot of the left / main / right run
    uint32_t end[3];    // tile coordinate where each run ends; end[2] = atoms in the row
    int ib, ic;         // lattice image of the row (wrap along y, z)
};

// Exact re-evaluation of one pair from the original coordinates (rare: pairs inside an
// error band of a decision boundary, or every accepted pair of a sink without an FP32
// path).  (a, ib, ic) is the lattice image predicted for the cell pair.
template <int MIC, bool SELF, class Sink>
__device__ __noinline__ void row_exact(Sink sink, const BoxParams *box, double d2_max, double d2_min,
                                       float4 q, float4 co, int a, int ib, int ic, uint64_t f)
{
    double d2;
    if (MIC == MIC_TRICLINIC) {
        D3 d = sep_f32(q.x, q.y, q.z, co.x, co.y, co.z), img;
        d2 = tric_image_d2(d, box->h, a, -ib, -ic, img);
    } else {
        d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, co.x, co.y, co.z, *box);
    }
    if (d2 <= d2_max && d2 > d2_min) sink.template hit_w<SELF>(__float_as_int(q.w), __float_as_int(co.w), d2, f);
}

// Bookkeeping of one CTA (shared memory): the frame geometry, the run descriptors of every
// stencil row and the tile coordinate of every staged cell.
struct RowBook {
    FrameGeom fg;
    RowInfo row[ROW_MAXROWS];
    uint32_t off[ROW_MAXROWS][ROW_MAXC + 1];   // tile coordinate of every staged cell
    uint32_t qst[ROW_MAXR + 1];                // first query slot of every home cell
};

// What row_prepare() derives for the CTA (identical in every thread).
struct RowJob {
    int nrows, nh, NC, W;
    bool empty;                                // the CTA lies outside this frame's grid
};

// Phases shared by the row kernels: copy the frame geometry, count the atoms of every staged
// cell, prefix them per row and describe the (up to three) runs of every row.
// SELF visits the half shell: the home row, then {dz = 0, dy > 0}, then {dz > 0}.
template <bool SELF, int T>
__device__ __forceinline__ RowJob row_prepare(RowBook &bk, const uint32_t *__restrict__ q_start,
                                              const uint32_t *__restrict__ c_start,
                                              const FrameGeom *__restrict__ fgs, int per_frame,
                                              RowParams rp, uint64_t f)
{
    constexpr int NW = T / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(fgs + (per_frame ? f : 0));
        uint32_t *dst = reinterpret_cast<uint32_t *>(&bk.fg);
        for (int w = tid; w < (int)(sizeof(FrameGeom) / 4); w += T) dst[w] = src[w];
    }
    __syncthreads();
    RowJob job;
    const int W = rp.W, R = rp.R;
    const int nx = bk.fg.grid.ncells[0], ny = bk.fg.grid.ncells[1], nz = bk.fg.grid.ncells[2];
    const int runs = (nx + R - 1) / R;
    job.W = W;
    job.empty = blockIdx.x >= (unsigned)(runs * ny * nz);          // frames may have fewer cells (uniform)
    job.nrows = job.nh = job.NC = 0;
    if (job.empty) return job;
    const int xr = blockIdx.x % runs, y = (blockIdx.x / runs) % ny, z = blockIdx.x / (runs * ny);
    const int x0 = xr * R, x1 = min(x0 + R, nx), nh = x1 - x0;
    const int64_t base = per_frame ? bk.fg.cell_base : (int64_t)f * bk.fg.grid.ncell_total;
    const int px = bk.fg.grid.periodic[0], py = bk.fg.grid.periodic[1], pz = bk.fg.grid.periodic[2];
    const int side = 2 * W + 1;
    const int nrows = SELF ? (W + 1) + W * side : side * side;
    const int NC = nh + 2 * W;                                     // staged cells of a full row
    job.nrows = nrows; job.nh = nh; job.NC = NC;
    auto row_dydz = [&](int r, int &dy, int &dz) {
        if (SELF) {
            if (r <= W) { dy = r; dz = 0; }
            else { dz = 1 + (r - (W + 1)) / side; dy = (r - (W + 1)) % side - W; }
        } else { dz = r / side - W; dy = r % side - W; }
    };
    // ---- phase 1: atoms per staged cell ------------------------------------------------
    for (int idx = tid; idx < nrows * NC; idx += T) {
        const int r = idx / NC, j = idx - r * NC;
        int dy, dz;
        row_dydz(r, dy, dz);
        const bool home = SELF && r == 0;
        int zz = z + dz, yy = y + dy;
        bool ok = true;
        if (zz < 0) { ok = ok && pz; zz += nz; } else if (zz >= nz) { ok = ok && pz; zz -= nz; }
        if (yy < 0) { ok = ok && py; yy += ny; } else if (yy >= ny) { ok = ok && py; yy -= ny; }
        int v = (home ? x0 : x0 - W) + j;
        if (home && j >= nh + W) ok = false;
        if (v < 0) { ok = ok && px; v += nx; } else if (v >= nx) { ok = ok && px; v -= nx; }
        uint32_t cnt = 0;
        if (ok) {
            const int64_t cell = base + ((int64_t)zz * ny + yy) * nx + v;
            cnt = c_start[cell + 1] - c_start[cell];
        }
        bk.off[r][j] = cnt;
    }
    const int64_t home_row = base + ((int64_t)z * ny + y) * nx;
    for (int j = tid; j <= nh; j += T) bk.qst[j] = q_start[home_row + x0 + j];
    __syncthreads();
    // ---- phase 2: exclusive prefix per row, run descriptors ----------------------------------
    for (int r = warp; r < nrows; r += NW) {
        const bool home = SELF && r == 0;
        const int ncell = home ? nh + W : NC;
        const uint32_t cnt = lane < ncell ? bk.off[r][lane] : 0u;
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        __syncwarp();
        if (lane <= ncell) bk.off[r][lane] = incl - cnt;              // lane == ncell: the row total
        __syncwarp();
        if (lane == 0) {
            int dy, dz;
            row_dydz(r, dy, dz);
            int zz = z + dz, yy = y + dy, ib = 0, ic = 0;
            if (zz < 0) { zz += nz; ic = -1; } else if (zz >= nz) { zz -= nz; ic = 1; }
            if (yy < 0) { yy += ny; ib = -1; } else if (yy >= ny) { yy -= ny; ib = 1; }
            const int vlo = home ? x0 : x0 - W, vhi = vlo + ncell - 1;
            const int nleft = vlo < 0 ? -vlo : 0, nright = vhi >= nx ? vhi - nx + 1 : 0;
            RowInfo ri;
            ri.ib = ib; ri.ic = ic;
            ri.end[0] = bk.off[r][nleft];
            ri.end[1] = bk.off[r][ncell - nright];
            ri.end[2] = bk.off[r][ncell];
            ri.g[0] = ri.g[1] = ri.g[2] = 0u;
            if (ri.end[2]) {                                             // a populated row is inside the grid
                const int64_t row = base + ((int64_t)zz * ny + yy) * nx;
                if (ri.end[0]) ri.g[0] = c_start[row + nx - nleft];
                if (ri.end[1] > ri.end[0]) ri.g[1] = c_start[row + (vlo < 0 ? 0 : vlo)];
                if (ri.end[2] > ri.end[1]) ri.g[2] = c_start[row];
            }
            bk.row[r] = ri;
        }
    }
    __syncthreads();
    return job;
}

template <int MIC, bool SELF, class Sink>
__global__ void __launch_bounds__(Sink::THREADS)
row_tile_kernel(const float4 *__restrict__ q_atoms, const uint32_t *__restrict__ q_start,
                const float4 *__restrict__ c_atoms, const uint32_t *__restrict__ c_start,
                const FrameGeom *__restrict__ fgs, int per_frame, RowParams rp, SearchParams sp, Sink sink)
{
    constexpr int T = Sink::THREADS;
    __shared__ RowBook bk;
    __shared__ __align__(128) float4 tile[2][ROW_TILE];
    __shared__ uint64_t bar[2];
    const int tid = threadIdx.x;
    const uint64_t f = blockIdx.y;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    sink.begin();
    const RowJob job = row_prepare<SELF, T>(bk, q_start, c_start, fgs, per_frame, rp, f);
    if (job.empty) return;
    const int W = job.W, nh = job.nh, NC = job.NC, nrows = job.nrows;
    const FrameGeom &sfg = bk.fg;
    RowInfo *s_row = bk.row;
    uint32_t (*s_off)[ROW_MAXC + 1] = bk.off;
    uint32_t *s_qst = bk.qst;
    const BoxParams &box = sfg.box;
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
    const uint32_t qs0 = s_qst[0], qs1 = s_qst[nh];
    // stage (row r, tile offset p) -> buffer b: one thread issues up to three bulk copies
    auto issue = [&](int r, uint32_t p, int b) {
        const RowInfo &ri = s_row[r];
        const uint32_t pe = min(p + (uint32_t)ROW_TILE, ri.end[2]);
        mbar_expect_tx(&bar[b], (pe - p) * (uint32_t)sizeof(float4));
        uint32_t sb = 0;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const uint32_t se = ri.end[s];
            const uint32_t a = max(sb, p), e = min(se, pe);
            if (a < e) bulk_g2s(&tile[b][a - p], c_atoms + ri.g[s] + (a - sb), (e - a) * (uint32_t)sizeof(float4), &bar[b]);
            sb = se;
        }
    };
    auto next_stage = [&](int &r, uint32_t &p) {
        p += ROW_TILE;
        if (p >= s_row[r].end[2]) {
            p = 0;
            do { ++r; } while (r < nrows && s_row[r].end[2] == 0u);
        }
    };
    uint32_t si = 0;                                               // stages issued so far (buffer = si & 1)
    for (uint32_t qbase = qs0; qbase < qs1; qbase += T) {
        const uint32_t qslot = qbase + tid;
        const bool active = qslot < qs1;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        int cxl = 0;                                               // home cell of the query, relative to x0
        if (active) {
            q = q_atoms[qslot];
            while (cxl + 1 < nh && s_qst[cxl + 1] <= qslot) ++cxl;
        }
        const int qi = __float_as_int(q.w);
        if (SELF && sp.include_self && active) {                   // capped_distance(x, x) reports (i, i)
            const double d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, q.x, q.y, q.z, box);
            if (d2 <= sp.d2_max && d2 > sp.d2_min) sink.template hit_w<false>(qi, qi, d2, f);
        }
        int r = 0;
        uint32_t p = 0;
        while (r < nrows && s_row[r].end[2] == 0u) ++r;
        if (r < nrows && tid == 0) issue(r, p, (int)(si & 1u));
        while (r < nrows) {
            int rn = r;
            uint32_t pn = p;
            next_stage(rn, pn);
            if (rn < nrows && tid == 0) issue(rn, pn, (int)((si + 1u) & 1u));
            const RowInfo &ri = s_row[r];
            const bool home = SELF && r == 0;
            // window of this thread in tile coordinates
            uint32_t lo = 0, hi = 0;
            if (active) {
                const int ncell = home ? nh + W : NC;
                lo = s_off[r][cxl];
                hi = s_off[r][min(cxl + (home ? W : 2 * W) + 1, ncell)];
                if (home) lo += (qslot - s_qst[cxl]) + 1;          // later atoms of the own cell
            }
            const uint32_t pe = min(p + (uint32_t)ROW_TILE, ri.end[2]);
            lo = max(lo, p);
            hi = min(hi, pe);
            // the row's lattice translation goes onto the query: q - (c + t) = (q - t) - c
            const float qrx = q.x - ((float)ri.ib * hy[0] + (float)ri.ic * hz[0]);
            const float qry = q.y - ((float)ri.ib * hy[1] + (float)ri.ic * hz[1]);
            const float qrz = q.z - ((float)ri.ib * hy[2] + (float)ri.ic * hz[2]);
            mbar_wait(&bar[si & 1u], (si >> 1) & 1u);
            const float4 *tb = tile[si & 1u];
            uint32_t sb = 0;
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
                const uint32_t se = ri.end[s];
                const uint32_t a = max(lo, sb), e = min(hi, se);
                const float sgn = (float)(1 - s);                  // left run: image -h0, i.e. q + h0
                const float qx = qrx + sgn * hx[0], qy = qry + sgn * hx[1], qz = qrz + sgn * hx[2];
                uint32_t t = a;
                if constexpr (Sink::FAST) {
                    // groups of four: straight-line FP32 filter + binning, one test for the rare exact path
                    for (; t + 4 <= e; t += 4) {
                        bool slow[4], ok[4];
                        uint32_t addr[4];
                        int cj[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float4 c = tb[t + k - p];
                            const float tx = qx - c.x, ty = qy - c.y, tz = qz - c.z;
                            const float d2f = fmaf(tx, tx, fmaf(ty, ty, tz * tz));
                            ok[k] = sink.fast_bin(d2f, addr[k]);
                            slow[k] = d2f <= sp.d2_filter && !ok[k];
                            cj[k] = __float_as_int(c.w);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) sink.template fast_add<SELF>(ok[k], addr[k], qi, cj[k]);
                        if (slow[0] | slow[1] | slow[2] | slow[3]) {
#pragma unroll 1
                            for (int k = 0; k < 4; ++k) {
                                if (!((k == 0 && slow[0]) || (k == 1 && slow[1]) || (k == 2 && slow[2]) || (k == 3 && slow[3]))) continue;
                                row_exact<MIC, SELF, Sink>(sink, &box, sp.d2_max, sp.d2_min, q, c_atoms[ri.g[s] + (t + k - sb)],
                                                           1 - s, ri.ib, ri.ic, f);
                            }
                        }
                    }
                }
                for (; t < e; ++t) {
                    const float4 c = tb[t - p];
                    const float tx = qx - c.x, ty = qy - c.y, tz = qz - c.z;
                    const float d2f = fmaf(tx, tx, fmaf(ty, ty, tz * tz));
                    bool done = false;
                    if (Sink::FAST) done = sink.template try_fast_w<SELF>(qi, __float_as_int(c.w), d2f);
                    if (d2f <= sp.d2_filter && !done)
                        row_exact<MIC, SELF, Sink>(sink, &box, sp.d2_max, sp.d2_min, q, c_atoms[ri.g[s] + (t - sb)],
                                                   1 - s, ri.ib, ri.ic, f);
                }
                sb = se;
            }
            __syncthreads();                                       // buffer si & 1 may be refilled
            r = rn;
            p = pn;
            ++si;
        }
    }
    sink.end();
}

}  // namespace mdagpu
