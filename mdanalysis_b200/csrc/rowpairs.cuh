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
// Staging: the candidates of a stencil row (float4 {x, y, z, index}) are copied to shared
// memory by all threads, two buffers, one __syncthreads per row.
#pragma once
#include "rowtile.cuh"

namespace mdagpu {

constexpr int RP_TILE = 1024;             // staged candidates per buffer
constexpr int RP_QUEUE = 96;              // queued survivors per warp (drained at >= 32)

template <int MIC, bool SELF, class Sink>
__global__ void __launch_bounds__(Sink::THREADS)
row_pairs_kernel(const float4 *__restrict__ q_atoms, const uint32_t *__restrict__ q_start,
                 const float4 *__restrict__ c_atoms, const uint32_t *__restrict__ c_start,
                 const FrameGeom *__restrict__ fgs, int per_frame, RowParams rp, SearchParams sp, Sink sink)
{
    constexpr int T = Sink::THREADS, NW = T / 32;
    __shared__ RowBook bk;
    __shared__ __align__(16) float4 tile[2][RP_TILE];
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
