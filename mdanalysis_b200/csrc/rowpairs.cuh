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
// r02 (plane path).  ncu of the ballot-compacting walk (profiles/r02_ncu_fastns_row_pairs_kernel.txt): 38
// thread-instructions per candidate, of which the compaction (ballot, popc, three stores, queue
// bookkeeping -- executed for nearly every candidate, a warp of 32 queries almost always has a
// survivor) cost as much as the filter.  Now every lane keeps its OWN survivors (a predicated
// 2-byte store and an add per candidate, [slot][lane] layout), the filter arithmetic is packed
// (FADD2 / FFMA2: two candidates per issue slot), and when a lane's queue is nearly full the
// warp evaluates the queued pairs two per lane and round -- each lane its own, so the query
// stays in registers (no shuffles) and the image code comes from the entry's row.
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
constexpr int RP_PAD = 8;                 // sentinels behind the last row (a trip may read past the end)
#ifndef RP_THREADS_
#define RP_THREADS_ 128
#endif
#ifndef RP_LQ_
#define RP_LQ_ 24
#endif
#ifndef RP_MINB_
#define RP_MINB_ 6
#endif
constexpr int RP_THREADS = RP_THREADS_;   // 4 warps claim the (chunk, row part) items of the CTA; 6 CTAs of 80 registers per SM
                                          // (profiles/microbench/fastns_variants_b200.txt: 192 / 160 / 256 threads are within 5 %)
constexpr int RP_LQ = RP_LQ_;                 // survivors a lane can hold between two drains
constexpr int RP_WQ = 128;                // compacted warp queue (evaluated 64 at a time)
constexpr int RP_PARTROWS = 16;           // rows per work item (4 bits of a queue entry)
// dynamic shared memory for a plane capacity of `cap` candidates (a multiple of 4, <= RP_ALL); the two
// buffers of the row-by-row path live in the same allocation (cap * 12 / 32 candidates each)
__host__ __device__ constexpr size_t rp_dyn_bytes(int cap) { return 3 * (size_t)(cap + RP_PAD) * sizeof(float); }
constexpr int RP_MINCAP = 512;

template <int MIC, bool SELF, class Sink, int T_ = RP_THREADS>
__global__ void __launch_bounds__(T_, RP_MINB_)
row_pairs_kernel(const float4 *__restrict__ q_atoms, const uint32_t *__restrict__ q_start,
                 const float4 *__restrict__ c_atoms, const uint32_t *__restrict__ c_start,
                 const FrameGeom *__restrict__ fgs, int per_frame, RowParams rp, SearchParams sp, Sink sink, int cap)
{
    constexpr int T = T_, NW = T / 32;
    __shared__ RowBook bk;
    extern __shared__ __align__(16) unsigned char rp_dyn[];         // rp_dyn_bytes(cap)
    const uint32_t rtile = (uint32_t)(rp_dyn_bytes(cap) / (2 * sizeof(float4)));   // candidates per buffer, row-by-row path
    float4 *tile0 = reinterpret_cast<float4 *>(rp_dyn);
    __shared__ uint32_t s_rowbase[ROW_MAXROWS + 1];
    // plane path: per row, where the runs end in the tile and what turns a tile position of a run into
    // a global slot; the row's image (ib, ic).  {end of run 0, end of run 1, offset 0, offset 1, offset 2, ib | ic << 8}
    __shared__ uint32_t s_runs[ROW_MAXROWS][6];
    __shared__ uint32_t s_next;                                     // next unclaimed work item (plane path)
    // plane path: per-lane survivor queues [warp][slot][lane]; row-by-row path: per-warp compacted queues
    constexpr int Q_PLANE = NW * (RP_LQ * 32 * 2 + RP_WQ * 4), Q_ROW = NW * RP_QUEUE * 9;
    __shared__ __align__(16) unsigned char s_queue[Q_PLANE > Q_ROW ? Q_PLANE : Q_ROW];
    uint32_t (*s_qq)[RP_QUEUE] = reinterpret_cast<uint32_t (*)[RP_QUEUE]>(s_queue);
    uint32_t (*s_qc)[RP_QUEUE] = reinterpret_cast<uint32_t (*)[RP_QUEUE]>(s_queue + NW * RP_QUEUE * 4);
    uint8_t (*s_qi)[RP_QUEUE] = reinterpret_cast<uint8_t (*)[RP_QUEUE]>(s_queue + NW * RP_QUEUE * 8);
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

    // stage (row r, offset p .. p + rtile) into buffer b: all threads, coalesced 16-byte loads
    auto stage = [&](int r, uint32_t p, int b) {
        const RowInfo &ri = bk.row[r];
        const uint32_t pe = min(p + rtile, ri.end[2]);
        for (uint32_t t = p + tid; t < pe; t += T) {
            const int s = t >= ri.end[1] ? 2 : (t >= ri.end[0] ? 1 : 0);
            tile0[(uint32_t)b * rtile + (t - p)] = c_atoms[ri.g[s] + (t - (s ? ri.end[s - 1] : 0u))];
        }
    };
    auto next_stage = [&](int &r, uint32_t &p) {
        p += rtile;
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
    if (s_rowbase[nrows] <= (uint32_t)cap) {
        float *px = reinterpret_cast<float *>(rp_dyn), *py = px + (cap + RP_PAD), *pz = py + (cap + RP_PAD);
        for (int r = 0; r < nrows; ++r) {
            const RowInfo &ri = bk.row[r];
            const uint32_t rb = s_rowbase[r];
            for (uint32_t t = tid; t < ri.end[2]; t += T) {
                const int s = t >= ri.end[1] ? 2 : (t >= ri.end[0] ? 1 : 0);
                const float4 c = c_atoms[ri.g[s] + (t - (s ? ri.end[s - 1] : 0u))];
                px[rb + t] = c.x; py[rb + t] = c.y; pz[rb + t] = c.z;
            }
        }
        for (int r = tid; r < nrows; r += T) {
            const RowInfo &ri = bk.row[r];
            const uint32_t rb = s_rowbase[r];
            s_runs[r][0] = rb + ri.end[0];
            s_runs[r][1] = rb + ri.end[1];
            s_runs[r][2] = ri.g[0] - rb;
            s_runs[r][3] = ri.g[1] - rb - ri.end[0];
            s_runs[r][4] = ri.g[2] - rb - ri.end[1];
            s_runs[r][5] = (uint32_t)(ri.ib + 1) | (uint32_t)(ri.ic + 1) << 8;
        }
        if (tid < RP_PAD) {                                         // the last trip may read past the end
            const uint32_t e = s_rowbase[nrows] + tid;
            px[e] = py[e] = pz[e] = 1.0e18f;
        }
        if (tid == 0) s_next = (uint32_t)NW;
        __syncthreads();
        // Work items = (chunk of 32 queries, part of the rows), part-major; warp w starts with item w and
        // claims the next unclaimed one when it is done.  A home cell of ~82 queries is three chunks: with
        // two row parts that is one item for each of the six warps.
        const int nchunk = (int)((qs1 - qs0 + 31u) >> 5);
        const int nparts = nrows >= 4 ? max(2, (nrows + RP_PARTROWS - 1) / RP_PARTROWS) : 1;
        const int nitems = nchunk * nparts;
        const uint32_t lq0 = smem_u32(s_queue) + (uint32_t)(warp * RP_LQ * 64 + lane * 2);
        const uint32_t px32 = smem_u32(px);
        const uint32_t PLANE = (uint32_t)(cap + RP_PAD) * 4u;         // bytes between the planes
        const uint32_t wq0 = smem_u32(s_queue) + (uint32_t)(NW * RP_LQ * 64 + warp * RP_WQ * 4);
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
            uint32_t lqp = lq0;                                     // where this lane's next survivor goes (64 bytes per entry)
            const uint32_t lqlim = lq0 + (uint32_t)(RP_LQ - 8) * 64u;
            uint32_t nwq = 0;                                       // entries in the compacted warp queue (warp-uniform)
            // float64 evaluation of 64 (or the last < 64) entries of the warp queue, two per lane: an entry
            // is (tile position | row within the part << 12 | owner lane << 16); the query comes by
            // shuffle from its lane, the run -- and with it the global slot and the image -- from the
            // row's descriptor; both candidate gathers are in flight together and the accepted pairs of
            // both rounds share one atomic.
            auto evaluate64 = [&](uint32_t first, uint32_t count) {
                __syncwarp();
                const bool v0 = (uint32_t)lane < count, v1 = (uint32_t)lane + 32u < count;
                uint32_t e0 = 0, e1 = 0;
                if (v0) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e0) : "r"(wq0 + 4u * (first + lane)));
                if (v1) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e1) : "r"(wq0 + 4u * (first + lane + 32u)));
                auto locate = [&](uint32_t e, int &ia, int &ib, int &ic) -> uint32_t {
                    const uint32_t *rr = s_runs[ra + (int)((e >> 12) & 15u)];
                    const uint32_t pos = e & 4095u;
                    const int sx = pos >= rr[1] ? 2 : (pos >= rr[0] ? 1 : 0);
                    const uint32_t img = rr[5];
                    ia = 1 - sx; ib = 1 - (int)(img & 255u); ic = 1 - (int)(img >> 8);
                    return pos + rr[2 + sx];
                };
                int a0 = 0, b0 = 0, g0 = 0, a1 = 0, b1 = 0, g1 = 0;
                const uint32_t s0 = locate(e0, a0, b0, g0), s1 = locate(e1, a1, b1, g1);
                float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
                if (v0) c0 = c_atoms[s0];
                if (v1) c1 = c_atoms[s1];
                const uint32_t l0 = e0 >> 16, l1 = e1 >> 16;
                float4 q0, q1;
                q0.x = __shfl_sync(0xffffffffu, q.x, l0); q0.y = __shfl_sync(0xffffffffu, q.y, l0);
                q0.z = __shfl_sync(0xffffffffu, q.z, l0); q0.w = __shfl_sync(0xffffffffu, q.w, l0);
                q1.x = __shfl_sync(0xffffffffu, q.x, l1); q1.y = __shfl_sync(0xffffffffu, q.y, l1);
                q1.z = __shfl_sync(0xffffffffu, q.z, l1); q1.w = __shfl_sync(0xffffffffu, q.w, l1);
                auto exact = [&](const float4 &qa, const float4 &ca, int ia, int ib, int ic) {
                    if (MIC == MIC_TRICLINIC) {
                        D3 d = sep_f32(qa.x, qa.y, qa.z, ca.x, ca.y, ca.z), img;
                        return tric_image_d2(d, box.h, ia, ib, ic, img);
                    }
                    return pair_d2_exact<MIC>(qa.x, qa.y, qa.z, ca.x, ca.y, ca.z, box);
                };
                const double d0 = exact(q0, c0, a0, b0, g0), d1 = exact(q1, c1, a1, b1, g1);
                sink.hit2(v0 && d0 <= sp.d2_max && d0 > sp.d2_min, __float_as_int(q0.w), __float_as_int(c0.w), d0,
                          v1 && d1 <= sp.d2_max && d1 > sp.d2_min, __float_as_int(q1.w), __float_as_int(c1.w), d1);
                __syncwarp();
            };
            // bulk compaction of the 32 per-lane queues into the warp queue: one ballot per slot row
            auto drain = [&]() {
                __syncwarp();
                const uint32_t cnt = (lqp - lq0) >> 6;
                const uint32_t rows = __reduce_max_sync(0xffffffffu, cnt);
                const uint32_t tag = (uint32_t)lane << 16, lt = (1u << lane) - 1u;
                for (uint32_t j = 0; j < rows; ++j) {
                    const bool v = j < cnt;
                    const uint32_t m = __ballot_sync(0xffffffffu, v);
                    if (v) {
                        uint32_t e;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(lq0 + 64u * j));
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(wq0 + 4u * (nwq + __popc(m & lt))), "r"(e | tag) : "memory");
                    }
                    nwq += __popc(m);
                    if (nwq >= 64u) {
                        evaluate64(nwq - 64u, 64u);
                        nwq -= 64u;
                    }
                }
                lqp = lq0;
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
                const uint32_t rbase = s_rowbase[r];
                const uint32_t rtag = (uint32_t)(r - ra) << 12;
                uint32_t sb = 0;
#pragma unroll 1
                for (int s = 0; s < 3; ++s) {
                    const uint32_t se = ri.end[s];
                    const uint32_t a = max(wlo, sb), e = min(whi, se);      // warp-uniform
                    if (a < e) {
                        const float sgn = (float)(1 - s);
                        // minus the translated query, twice: dx = c + (-q) has the same square as q - c
                        const float nx = -(qrx + sgn * hx[0]), ny = -(qry + sgn * hx[1]), nz = -(qrz + sgn * hx[2]);
                        const float2 nqx = make_float2(nx, nx), nqy = make_float2(ny, ny), nqz = make_float2(nz, nz);
                        // the lane's own part of this run as (first, width) for one unsigned compare
                        const uint32_t lo2 = max(lo, a), hi2 = min(hi, e);
                        const uint32_t width = hi2 > lo2 ? hi2 - lo2 : 0u;
                        const uint32_t qd0 = (rbase + a) & ~3u, qd1 = rbase + e;
                        uint32_t rel = qd0 - rbase - lo2;                   // row coordinate of candidate 0, minus lo2 (mod 2^32)
                        uint32_t val = qd0 | rtag;
                        // four candidates: three LDS.128, packed distance, a predicated 2-byte store per survivor
                        auto trip = [&](uint32_t ptr, uint32_t rel4, uint32_t val4) {
                            float4 X, Y, Z;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(X.x), "=f"(X.y), "=f"(X.z), "=f"(X.w) : "r"(ptr));
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(Y.x), "=f"(Y.y), "=f"(Y.z), "=f"(Y.w) : "r"(ptr + PLANE));
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(Z.x), "=f"(Z.y), "=f"(Z.z), "=f"(Z.w) : "r"(ptr + 2u * PLANE));
                            const float2 dxa = __fadd2_rn(make_float2(X.x, X.y), nqx), dxb = __fadd2_rn(make_float2(X.z, X.w), nqx);
                            const float2 dya = __fadd2_rn(make_float2(Y.x, Y.y), nqy), dyb = __fadd2_rn(make_float2(Y.z, Y.w), nqy);
                            const float2 dza = __fadd2_rn(make_float2(Z.x, Z.y), nqz), dzb = __fadd2_rn(make_float2(Z.z, Z.w), nqz);
                            const float2 da = __ffma2_rn(dxa, dxa, __ffma2_rn(dya, dya, __fmul2_rn(dza, dza)));
                            const float2 db = __ffma2_rn(dxb, dxb, __ffma2_rn(dyb, dyb, __fmul2_rn(dzb, dzb)));
                            const float d2f[4] = { da.x, da.y, db.x, db.y };
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (d2f[k] <= sp.d2_filter && rel4 + (uint32_t)k < width) {
                                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(lqp), "r"(val4 + (uint32_t)k) : "memory");
                                    lqp += 64u;
                                }
                            }
                        };
                        // two trips per queue check (a trip past the end of the run reads the next row or the
                        // sentinels: outside every lane's window)
                        for (uint32_t ptr = px32 + 4u * qd0; ptr < px32 + 4u * qd1; ptr += 32u, rel += 8u, val += 8u) {
                            trip(ptr, rel, val);
                            trip(ptr + 16u, rel + 4u, val + 4u);
                            if (__any_sync(0xffffffffu, lqp > lqlim)) drain();
                        }
                    }
                    sb = se;
                }
            }
            if (__any_sync(0xffffffffu, lqp != lq0)) drain();
            if (nwq) evaluate64(0u, nwq);                             // < 64 left
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
                const uint32_t pe = min(p + rtile, ri.end[2]);
                const uint32_t wlo = max(bk.off[r][cmin], p), whi = min(bk.off[r][min(cmax + span + 1, ncell)], pe);
                // the row's lattice translation goes onto the query: q - (c + t) = (q - t) - c
                const float qrx = q.x - ((float)ri.ib * hy[0] + (float)ri.ic * hz[0]);
                const float qry = q.y - ((float)ri.ib * hy[1] + (float)ri.ic * hz[1]);
                const float qrz = q.z - ((float)ri.ib * hy[2] + (float)ri.ic * hz[2]);
                const float4 *tb = tile0 + (uint32_t)b * rtile;
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
