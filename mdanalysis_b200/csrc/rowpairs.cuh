// rowpairs.cuh -- the cutoff pair search that RETURNS pairs (FastNS.search / self_search,
// lib/nsgrid.pyx:499-644), row-staged like rowtile.cuh but organised around the fact that
// here every accepted candidate needs the float64 distance.
//
// With ~10 % of the candidates inside the cutoff, a warp finds a hit in almost every step, so
// evaluating the float64 distance where the hit is found keeps 3 lanes busy and 29 waiting
// (measured: 11.7 ms for 1 M atoms / 107 M pairs with rowtile.cuh's per-pair path).  Instead:
//   * a warp walks ONE window (the union of its lanes' cell windows) so that the candidate
//     loop is warp-uniform; a lane only accepts candidates inside its OWN window [cell - W,
//     cell + W] -- for triclinic cells the reference's stencil can be incomplete and the
//     window is part of the result;
//   * the FP32 filter marks survivors, __ballot_sync compacts them (no atomics) into a
//     per-warp queue of (query slot, candidate slot, image);
//   * whenever 32 are queued, all 32 lanes evaluate one each with the reference's float64
//     arithmetic (the image predicted for the cell pair, geom.cuh) and append the accepted
//     ones through the sink (warp-aggregated global atomic, engine.cuh PairSink).
// Staging.  When all stencil rows of the CTA fit (RP_ALL candidates; 1 M atoms at 0.1 / A^3 and an
// 8 A cutoff: 14 rows x ~234), they are staged ONCE as x / y / z planes and every warp then
// runs through its rows on its own: with one __syncthreads per row the warps that had just
// drained their queue arrived late and everybody else waited (ncu: 4.2 barrier-stalled warps
// per issue, 39 % issue use).  The planes also feed FOUR candidates per trip (three LDS.128,
// four ballots, one branch).  Rows that do not fit are staged one at a time as float4
// {x, y, z, index}, two buffers, one __syncthreads per row.
#pragma once
#include "rowtile.cuh"

namespace mdagpu {

constexpr int RP_TILE = 1024;             // staged candidates per buffer (row-by-row path)
constexpr int RP_ALL = 3584;              // staged candidates of ALL rows (plane path): 3 x 14 kB
constexpr int RP_QUEUE = 96;              // queued survivors per warp (drained at >= 32)
constexpr int RP_NQ = 1;                  // quads of candidates per trip of the plane path (2: spills, 3.2 -> 3.5 ms)
constexpr int RP_PAD = 4 * RP_NQ;         // sentinels behind the last row (a trip may read past the end)
constexpr size_t RP_DYN_BYTES = 3 * (RP_ALL + RP_PAD) * sizeof(float);   // >= 2 * RP_TILE * sizeof(float4)
static_assert(RP_DYN_BYTES >= 2 * RP_TILE * sizeof(float4), "the row-by-row buffers live in the same allocation");

template <int MIC, bool SELF, class Sink>
__global__ void __launch_bounds__(Sink::THREADS)
row_pairs_kernel(const float4 *__restrict__ q_atoms, const uint32_t *__restrict__ q_start,
                 const float4 *__restrict__ c_atoms, const uint32_t *__restrict__ c_start,
                 const FrameGeom *__restrict__ fgs, int per_frame, RowParams rp, SearchParams sp, Sink sink)
{
    constexpr int T = Sink::THREADS, NW = T / 32;
    __shared__ RowBook bk;
    extern __shared__ __align__(16) unsigned char rp_dyn[];         // RP_DYN_BYTES
    float4 (*tile)[RP_TILE] = reinterpret_cast<float4 (*)[RP_TILE]>(rp_dyn);
    __shared__ uint32_t s_rowbase[ROW_MAXROWS + 1];
    __shared__ uint32_t s_next;                                     // next unclaimed work item (plane path)
    __shared__ uint32_t s_qq[NW][RP_QUEUE], s_qc[NW][RP_QUEUE];
    __shared__ uint8_t s_qi[NW][RP_QUEUE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t f = blockIdx.y;
    sink.begin();
    const RowJob job = row_prepare<SELF, T>(bk, q_start, c_start, fgs, per_frame, rp, f);
    if (job.empty) return;
    const int W = job.W, nh = job.nh, NC = job.NC, nrows = job.nrows;
    const BoxParams &box = bk.fg.box;
    float hx[3], hy[3], hz[3];                                      // lattice rows as FP32 vectors
    if (MIC == MIC_TRICLINIC) {
        hx[0] = box.h[0]; hx[1] = box.h[1]; hx[2] = box.h[2];
        hy[0] = box.h[3]; hy[1] = box.h[4]; hy[2] = box.h[5];
        hz[0] = box.h[6]; hz[1] = box.h[7]; hz[2] = box.h[8];
    } else {
        hx[0] = box.L[0]; hx[1] = 0.f; hx[2] = 0.f;
        hy[0] = 0.f; hy[1] = box.L[1]; hy[2] = 0.f;
        hz[0] = 0.f; hz[1] = 0.f; hz[2] = box.L[2];
    }
    const uint32_t qs0 = bk.qst[0], qs1 = bk.qst[nh];
    uint32_t *qq = s_qq[warp], *qc = s_qc[warp];
    uint8_t *qi = s_qi[warp];
    uint32_t nqueued = 0;                                           // warp-uniform

    // float64 evaluation of up to 32 queued survivors, one per lane
    auto evaluate = [&](uint32_t first, uint32_t count) {
        __syncwarp();
        if ((uint32_t)lane < count) {
            const uint32_t e = first + lane;
            const float4 q = q_atoms[qq[e]], c = c_atoms[qc[e]];
            const int code = qi[e];                                 // (1 - s) + 1 | (ib + 1) << 2 | (ic + 1) << 4
            double d2;
            if (MIC == MIC_TRICLINIC) {
                D3 d = sep_f32(q.x, q.y, q.z, c.x, c.y, c.z), img;
                d2 = tric_image_d2(d, box.h, (code & 3) - 1, -(((code >> 2) & 3) - 1), -(((code >> 4) & 3) - 1), img);
            } else {
                d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, c.x, c.y, c.z, box);
            }
            if (d2 <= sp.d2_max && d2 > sp.d2_min)
                sink.template hit_w<SELF>(__float_as_int(q.w), __float_as_int(c.w), d2, f);
        }
        __syncwarp();
    };

    // stage (row r, offset p .. p + RP_TILE) into buffer b: all threads, coalesced 16-byte loads
    auto stage = [&](int r, uint32_t p, int b) {
        const RowInfo &ri = bk.row[r];
        const uint32_t pe = min(p + (uint32_t)RP_TILE, ri.end[2]);
        for (uint32_t t = p + tid; t < pe; t += T) {
            const int s = t >= ri.end[1] ? 2 : (t >= ri.end[0] ? 1 : 0);
            tile[b][t - p] = c_atoms[ri.g[s] + (t - (s ? ri.end[s - 1] : 0u))];
        }
    };
    auto next_stage = [&](int &r, uint32_t &p) {
        p += RP_TILE;
        if (p >= bk.row[r].end[2]) {
            p = 0;
            do { ++r; } while (r < nrows && bk.row[r].end[2] == 0u);
        }
    };

    // ---- plane path: every row staged once ------------------------------------------------
    if (tid == 0) {
        uint32_t b0 = 0;
        for (int r = 0; r < nrows; ++r) { s_rowbase[r] = b0; b0 += bk.row[r].end[2]; }
        s_rowbase[nrows] = b0;
    }
    __syncthreads();
    if (s_rowbase[nrows] <= (uint32_t)RP_ALL) {
        float *px = reinterpret_cast<float *>(rp_dyn), *py = px + (RP_ALL + RP_PAD), *pz = py + (RP_ALL + RP_PAD);
        for (int r = 0; r < nrows; ++r) {
            const RowInfo &ri = bk.row[r];
            const uint32_t rb = s_rowbase[r];
            for (uint32_t t = tid; t < ri.end[2]; t += T) {
                const int s = t >= ri.end[1] ? 2 : (t >= ri.end[0] ? 1 : 0);
                const float4 c = c_atoms[ri.g[s] + (t - (s ? ri.end[s - 1] : 0u))];
                px[rb + t] = c.x; py[rb + t] = c.y; pz[rb + t] = c.z;
            }
        }
        if (tid < RP_PAD) {                                         // the last trip may read past the end
            const uint32_t e = s_rowbase[nrows] + tid;
            px[e] = py[e] = pz[e] = 1.0e18f;
        }
        if (tid == 0) s_next = (uint32_t)NW;
        __syncthreads();
        // Work items = (chunk of 32 queries, part of the rows), part-major; warp w starts with item w and
        // claims the next unclaimed one when it is done.  A home cell of ~82 queries is three chunks: with
        // one static chunk per warp the fourth warp of the CTA had nothing to do and the 51 kB of the CTA
        // stayed allocated until the slowest warp was through all 14 rows.
        const int nchunk = (int)((qs1 - qs0 + 31u) >> 5);
        const int nparts = nrows >= 4 ? 2 : 1;
        const int nitems = nchunk * nparts;
        for (int item = warp; item < nitems;) {
            const int part = item / nchunk, chunk = item - part * nchunk;
            const int ra = nrows * part / nparts, rb = nrows * (part + 1) / nparts;
            const uint32_t qslot = qs0 + 32u * (uint32_t)chunk + lane;
            const bool active = qslot < qs1;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            int cxl = 0;
            if (active) {
                q = q_atoms[qslot];
                while (cxl + 1 < nh && bk.qst[cxl + 1] <= qslot) ++cxl;
            }
            if (SELF && sp.include_self && active && ra == 0) {
                const double d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, q.x, q.y, q.z, box);
                if (d2 <= sp.d2_max && d2 > sp.d2_min) sink.template hit_w<false>(__float_as_int(q.w), __float_as_int(q.w), d2, f);
            }
            const int cmin = (int)__reduce_min_sync(0xffffffffu, active ? (unsigned)cxl : 0x7fffffffu);
            const int cmax = (int)__reduce_max_sync(0xffffffffu, active ? (unsigned)cxl : 0u);
            // float64 evaluation of up to 64 queued survivors, two per lane: the queue names the lane
            // that owns the query (a shuffle instead of a gather), both candidates are fetched before
            // either is used, and the accepted pairs of both rounds share one atomic
            auto drain = [&](uint32_t first, uint32_t count) {
                __syncwarp();
                const bool v0 = (uint32_t)lane < count, v1 = (uint32_t)lane + 32u < count;
                const uint32_t e0 = first + lane, e1 = e0 + 32u;
                const uint32_t l0 = v0 ? qq[e0] : 0u, l1 = v1 ? qq[e1] : 0u;
                const int code0 = v0 ? qi[e0] : 0, code1 = v1 ? qi[e1] : 0;
                float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
                if (v0) c0 = c_atoms[qc[e0]];
                if (v1) c1 = c_atoms[qc[e1]];
                float4 q0, q1;
                q0.x = __shfl_sync(0xffffffffu, q.x, l0); q0.y = __shfl_sync(0xffffffffu, q.y, l0);
                q0.z = __shfl_sync(0xffffffffu, q.z, l0); q0.w = __shfl_sync(0xffffffffu, q.w, l0);
                q1.x = __shfl_sync(0xffffffffu, q.x, l1); q1.y = __shfl_sync(0xffffffffu, q.y, l1);
                q1.z = __shfl_sync(0xffffffffu, q.z, l1); q1.w = __shfl_sync(0xffffffffu, q.w, l1);
                auto exact = [&](const float4 &qa, const float4 &ca, int code) {
                    if (MIC == MIC_TRICLINIC) {
                        D3 d = sep_f32(qa.x, qa.y, qa.z, ca.x, ca.y, ca.z), img;
                        return tric_image_d2(d, box.h, (code & 3) - 1, -(((code >> 2) & 3) - 1), -(((code >> 4) & 3) - 1), img);
                    }
                    return pair_d2_exact<MIC>(qa.x, qa.y, qa.z, ca.x, ca.y, ca.z, box);
                };
                const double d0 = exact(q0, c0, code0), d1 = exact(q1, c1, code1);
                sink.hit2(v0 && d0 <= sp.d2_max && d0 > sp.d2_min, __float_as_int(q0.w), __float_as_int(c0.w), d0,
                          v1 && d1 <= sp.d2_max && d1 > sp.d2_min, __float_as_int(q1.w), __float_as_int(c1.w), d1);
                __syncwarp();
            };
            for (int r = ra; r < rb; ++r) {
                const RowInfo &ri = bk.row[r];
                if (ri.end[2] == 0u) continue;
                const bool home = SELF && r == 0;
                const int ncell = home ? nh + W : NC;
                const int span = home ? W : 2 * W;
                uint32_t lo = 0, hi = 0;
                if (active) {
                    lo = bk.off[r][cxl];
                    hi = bk.off[r][min(cxl + span + 1, ncell)];
                    if (home) lo += (qslot - bk.qst[cxl]) + 1;
                }
                const uint32_t wlo = bk.off[r][cmin], whi = bk.off[r][min(cmax + span + 1, ncell)];
                const float qrx = q.x - ((float)ri.ib * hy[0] + (float)ri.ic * hz[0]);
                const float qry = q.y - ((float)ri.ib * hy[1] + (float)ri.ic * hz[1]);
                const float qrz = q.z - ((float)ri.ib * hy[2] + (float)ri.ic * hz[2]);
                const uint32_t rb = s_rowbase[r];
                uint32_t sb = 0;
#pragma unroll 1
                for (int s = 0; s < 3; ++s) {
                    const uint32_t se = ri.end[s];
                    const uint32_t a = max(wlo, sb), e = min(whi, se);      // warp-uniform
                    if (a < e) {
                        const float sgn = (float)(1 - s);
                        const float qx = qrx + sgn * hx[0], qy = qry + sgn * hx[1], qz = qrz + sgn * hx[2];
                        const uint32_t code = (uint32_t)((1 - s) + 1) | (uint32_t)(ri.ib + 1) << 2 | (uint32_t)(ri.ic + 1) << 4;
                        const uint32_t gbase = ri.g[s] - sb;
                        // the lane's own part of this run as (first, width) for one unsigned compare
                        const uint32_t lo2 = max(lo, a), hi2 = min(hi, e);
                        const uint32_t width = hi2 > lo2 ? hi2 - lo2 : 0u;
                        for (uint32_t qd = (rb + a) & ~3u; qd < rb + e; qd += 4u * RP_NQ) {
                            float cxs[4 * RP_NQ], cys[4 * RP_NQ], czs[4 * RP_NQ];
#pragma unroll
                            for (int g = 0; g < RP_NQ; ++g) {
                                const float4 X = *reinterpret_cast<const float4 *>(px + qd + 4 * g);
                                const float4 Y = *reinterpret_cast<const float4 *>(py + qd + 4 * g);
                                const float4 Z = *reinterpret_cast<const float4 *>(pz + qd + 4 * g);
                                cxs[4 * g] = X.x; cxs[4 * g + 1] = X.y; cxs[4 * g + 2] = X.z; cxs[4 * g + 3] = X.w;
                                cys[4 * g] = Y.x; cys[4 * g + 1] = Y.y; cys[4 * g + 2] = Y.z; cys[4 * g + 3] = Y.w;
                                czs[4 * g] = Z.x; czs[4 * g + 1] = Z.y; czs[4 * g + 2] = Z.z; czs[4 * g + 3] = Z.w;
                            }
                            const uint32_t rel = qd - rb - lo2;     // row coordinate of candidate 0, minus lo2 (mod 2^32)
                            uint32_t m[4 * RP_NQ], many = 0u;
                            bool keep[4 * RP_NQ];
#pragma unroll
                            for (int k = 0; k < 4 * RP_NQ; ++k) {
                                const float tx = qx - cxs[k], ty = qy - cys[k], tz = qz - czs[k];
                                const float d2f = fmaf(tx, tx, fmaf(ty, ty, tz * tz));
                                keep[k] = d2f <= sp.d2_filter && rel + (uint32_t)k < width;
                                m[k] = __ballot_sync(0xffffffffu, keep[k]);
                                many |= m[k];
                            }
                            if (many) {
#pragma unroll
                                for (int k = 0; k < 4 * RP_NQ; ++k) {
                                    if (m[k] == 0u) continue;
                                    if (keep[k]) {
                                        const uint32_t at = nqueued + __popc(m[k] & ((1u << lane) - 1u));
                                        qq[at] = (uint32_t)lane;
                                        qc[at] = gbase + (qd - rb) + (uint32_t)k;
                                        qi[at] = (uint8_t)code;
                                    }
                                    nqueued += __popc(m[k]);
                                    if (nqueued >= 64u) {
                                        drain(nqueued - 64u, 64u);
                                        nqueued -= 64u;
                                    }
                                }
                            }
                        }
                    }
                    sb = se;
                }
            }
            if (nqueued) drain(0u, nqueued);                         // < 64 left
            nqueued = 0u;
            uint32_t nxt = 0;
            if (lane == 0) nxt = atomicAdd(&s_next, 1u);
            item = (int)__shfl_sync(0xffffffffu, nxt, 0);
        }
        sink.end();
        return;
    }

    // ---- row-by-row path -------------------------------------------------------------------
    for (uint32_t qbase = qs0; qbase < qs1; qbase += T) {
        const uint32_t qslot = qbase + tid;
        const bool active = qslot < qs1;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        int cxl = 0;                                               // home cell of the query, relative to x0
        if (active) {
            q = q_atoms[qslot];
            while (cxl + 1 < nh && bk.qst[cxl + 1] <= qslot) ++cxl;
        }
        if (SELF && sp.include_self && active) {                   // capped_distance(x, x) reports (i, i)
            const double d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, q.x, q.y, q.z, box);
            if (d2 <= sp.d2_max && d2 > sp.d2_min) sink.template hit_w<false>(__float_as_int(q.w), __float_as_int(q.w), d2, f);
        }
        const int cmin = (int)__reduce_min_sync(0xffffffffu, active ? (unsigned)cxl : 0x7fffffffu);
        const int cmax = (int)__reduce_max_sync(0xffffffffu, active ? (unsigned)cxl : 0u);
        const bool warp_active = cmin != 0x7fffffff;

        int r = 0;
        uint32_t p = 0;
        while (r < nrows && bk.row[r].end[2] == 0u) ++r;
        int b = 0;
        if (r < nrows) stage(r, p, 0);
        __syncthreads();
        while (r < nrows) {
            int rn = r;
            uint32_t pn = p;
            next_stage(rn, pn);
            if (rn < nrows) stage(rn, pn, b ^ 1);
            if (warp_active) {
                const RowInfo &ri = bk.row[r];
                const bool home = SELF && r == 0;
                const int ncell = home ? nh + W : NC;
                const int span = home ? W : 2 * W;
                // own window (row coordinates) and the warp's union window
                uint32_t lo = 0, hi = 0;
                if (active) {
                    lo = bk.off[r][cxl];
                    hi = bk.off[r][min(cxl + span + 1, ncell)];
                    if (home) lo += (qslot - bk.qst[cxl]) + 1;     // later atoms of the own cell
                }
                const uint32_t pe = min(p + (uint32_t)RP_TILE, ri.end[2]);
                const uint32_t wlo = max(bk.off[r][cmin], p), whi = min(bk.off[r][min(cmax + span + 1, ncell)], pe);
                // the row's lattice translation goes onto the query: q - (c + t) = (q - t) - c
                const float qrx = q.x - ((float)ri.ib * hy[0] + (float)ri.ic * hz[0]);
                const float qry = q.y - ((float)ri.ib * hy[1] + (float)ri.ic * hz[1]);
                const float qrz = q.z - ((float)ri.ib * hy[2] + (float)ri.ic * hz[2]);
                const float4 *tb = tile[b];
                uint32_t sb = 0;
#pragma unroll 1
                for (int s = 0; s < 3; ++s) {
                    const uint32_t se = ri.end[s];
                    const uint32_t a = max(wlo, sb), e = min(whi, se);      // warp-uniform trip
                    const float sgn = (float)(1 - s);              // left run: image -h0, i.e. q + h0
                    const float qx = qrx + sgn * hx[0], qy = qry + sgn * hx[1], qz = qrz + sgn * hx[2];
                    const uint32_t code = (uint32_t)((1 - s) + 1) | (uint32_t)(ri.ib + 1) << 2 | (uint32_t)(ri.ic + 1) << 4;
                    const uint32_t gbase = ri.g[s] - sb;           // row coordinate -> global slot
                    for (uint32_t t = a; t < e; ++t) {
                        const float4 c = tb[t - p];
                        const float tx = qx - c.x, ty = qy - c.y, tz = qz - c.z;
                        const float d2f = fmaf(tx, tx, fmaf(ty, ty, tz * tz));
                        const bool keep = d2f <= sp.d2_filter && t >= lo && t < hi;
                        const uint32_t m = __ballot_sync(0xffffffffu, keep);
                        if (m) {
                            if (keep) {
                                const uint32_t at = nqueued + __popc(m & ((1u << lane) - 1u));
                                qq[at] = qslot;
                                qc[at] = gbase + t;
                                qi[at] = (uint8_t)code;
                            }
                            nqueued += __popc(m);
                            if (nqueued >= 64u) {                  // two full rounds, lanes converged
                                evaluate(nqueued - 32u, 32u);
                                evaluate(nqueued - 64u, 32u);
                                nqueued -= 64u;
                            }
                        }
                    }
                    sb = se;
                }
            }
            __syncthreads();                                       // buffer b may be refilled, b ^ 1 is complete
            r = rn;
            p = pn;
            b ^= 1;
        }
        // leftovers of this pass
        while (nqueued) {
            const uint32_t n = min(nqueued, 32u);
            evaluate(nqueued - n, n);
            nqueued -= n;
        }
    }
    sink.end();
}

}  // namespace mdagpu
