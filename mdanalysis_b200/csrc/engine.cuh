// engine.cuh -- cell-list build and the cutoff pair traversal shared by FastNS,
// capped distances, InterRDF and contacts.
//
// Reference behaviour being reproduced (package/MDAnalysis/lib/nsgrid.pyx):
//   _prepare_box :294-387  grid geometry        -> grid_geometry_reference()
//   _pack_grid   :389-419  wrap + cell of atom  -> cell_key_kernel, counting sort
//   search       :499-580  27-cell stencil      -> pair_search_kernel<.., SELF=false>
//   self_search  :582-644  in-cell + 13 cells   -> pair_search_kernel<.., SELF=true>
//   calc_distsq  :483-497  float32 diff, double minimum image, d2 <= cutoff^2
//
// B200 design: instead of the reference's linked list (head_id/next_id) the
// atoms are counting-sorted by cell key (key = frame*ncell + cell) so that a
// cell is one contiguous run of float4 {x,y,z,index}; one thread owns one query
// (queries are processed in cell-sorted order, so the lanes of a warp read the
// same candidate runs and the 16-byte loads are L1 broadcasts).  Each candidate
// is first rejected in FP32 against a threshold widened by a rigorous error
// band; survivors are re-evaluated with the reference's exact float32->float64
// arithmetic (geom.cuh) and only that exact d2 takes the decision, which is
// what makes pair sets and histogram counts bit-identical.
#pragma once
#include "common.cuh"

namespace mdagpu {

// ---- geometry --------------------------------------------------------------------
// The reference's own grid (needed for parity on triclinic boxes, where its
// 27-cell stencil is not always complete).
void grid_geometry_reference(double cutoff, const float *box6, const float *h9, int pbc,
                             GridGeom *g, float *max_cutoff);
// Any complete orthorhombic grid with cells >= cutoff (used where the result is
// independent of the grid: orthorhombic / no-box searches).
void grid_geometry_ortho_free(double cutoff, const float *L3, const double *origin, int pbc,
                              GridGeom *g);
// a single cell holding everything (all-pairs / brute force)
void grid_geometry_single(GridGeom *g);
// bounding box of `total` points into lohi[6] = { min x, y, z, max x, y, z } (initialise with +-FLT_MAX; rdf.cu)
__global__ void bbox_kernel(const float *__restrict__ xyz, uint64_t total, float *__restrict__ lohi);

// Per frame-set description living in device memory (one entry per frame, or a
// single entry shared by all frames).
struct FrameGeom {
    GridGeom grid;
    BoxParams box;
    int64_t cell_base;     // offset of this frame's cells in the flattened key space
};

// ---- cell keys + counting sort -------------------------------------------------------
// keys[f*n + i] = cell_base[f] + cell(i); counts[key]++ .
// wrap: 0 = coordinates used as given; 1 = wrapped into the cell first (binned and
// stored wrapped, FastNS); 2 = binned by the wrapped position but stored as
// given (lib.distances ortho flavour works on unwrapped input).
__global__ void cell_key_kernel(const float *__restrict__ xyz, uint64_t n, uint64_t nframes,
                                const FrameGeom *__restrict__ fg, int per_frame, int wrap,
                                uint32_t *__restrict__ keys, uint32_t *__restrict__ counts,
                                uint32_t *__restrict__ ranks, unsigned *__restrict__ flag);
// exclusive scan of the per-cell counts into starts[], one CTA per frame (frame f starts at
// f * n_atoms); starts[total_cells] = total
__global__ void exclusive_scan_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ starts,
                                      const FrameGeom *__restrict__ fgs, int per_frame, uint64_t n_atoms,
                                      uint64_t nframes);
// sorted[starts[key] + ranks[t]] = {x,y,z,i} (ranks: the key kernel's atomicAdd results); wraps like cell_key_kernel
__global__ void cell_scatter_kernel(const float *__restrict__ xyz, uint64_t n, uint64_t nframes,
                                    const FrameGeom *__restrict__ fg, int per_frame, int wrap,
                                    const uint32_t *__restrict__ keys,
                                    const uint32_t *__restrict__ starts,
                                    const uint32_t *__restrict__ ranks, float4 *__restrict__ sorted,
                                    uint32_t *__restrict__ sorted_key);

struct CellList {
    DevBuf keys, counts, starts, cursor, sorted, sorted_key, flag;
    uint64_t n = 0, nframes = 0, total_cells = 0;
    // non-zero when some atom was folded into a cell it does not lie in (see cell_of_checked)
    const unsigned *inconsistent() const { return flag.as<unsigned>(); }
    const float4 *atoms() const { return sorted.as<float4>(); }
    const uint32_t *start() const { return starts.as<uint32_t>(); }
    const uint32_t *key_of_sorted() const { return sorted_key.as<uint32_t>(); }
};

// Builds the sorted cell list of `xyz` [nframes, n, 3] (device pointer).
int build_cell_list(const float *xyz, uint64_t n, uint64_t nframes, const FrameGeom *fg_dev,
                    int per_frame, uint64_t total_cells, int wrap, cudaStream_t st, CellList *out,
                    bool want_sorted_key = true);

// ---- FP32 pre-filter ----------------------------------------------------------------
// Returns an FP32 estimate of the minimum-image d2 whose relative error is far
// below the 1e-5 band the callers put on their thresholds.
template <int MIC>
__device__ __forceinline__ float pair_d2_filter(float tx, float ty, float tz, const BoxParams &b)
{
    if (MIC == MIC_ORTHO_LAZY) {
        tx = tx > b.half[0] ? tx - b.L[0] : (tx <= -b.half[0] ? tx + b.L[0] : tx);
        ty = ty > b.half[1] ? ty - b.L[1] : (ty <= -b.half[1] ? ty + b.L[1] : ty);
        tz = tz > b.half[2] ? tz - b.L[2] : (tz <= -b.half[2] ? tz + b.L[2] : tz);
        return fmaf(tx, tx, fmaf(ty, ty, tz * tz));
    } else if (MIC == MIC_TRICLINIC) {
        float best = FLT_MAX;
#pragma unroll
        for (int a = -1; a <= 1; ++a) {
            float xa = tx + b.h[0] * (float)a;
#pragma unroll
            for (int bb = -1; bb <= 1; ++bb) {
                float xab = xa + b.h[3] * (float)bb;
                float yb = ty + b.h[4] * (float)bb;
#pragma unroll
                for (int c = -1; c <= 1; ++c) {
                    float x = xab + b.h[6] * (float)c;
                    float y = yb + b.h[7] * (float)c;
                    float z = tz + b.h[8] * (float)c;
                    best = fminf(best, fmaf(x, x, fmaf(y, y, z * z)));
                }
            }
        }
        return best;
    } else {
        return fmaf(tx, tx, fmaf(ty, ty, tz * tz));
    }
}

// half-shell route of self_search (nsgrid.pyx:603-605), slot 0 = home cell;
// packed as (dx+1) | (dy+1)<<2 | (dz+1)<<4
__device__ __constant__ static const unsigned kSelfRoute[14] = { 21u, 22u, 26u, 25u, 24u, 6u, 10u,
                                                                 9u, 8u, 38u, 42u, 41u, 40u, 37u };

// Thresholds for the FP32 pre-filter.  `d_abs_err` is an absolute error bound on
// the FP32 distance estimate on top of its ~1e-6 relative error: zero except for
// the triclinic image sums, whose intermediate values are box-sized.
static inline double tric_abs_err(const float *h9)
{
    double s = 0.0;
    for (int k = 0; k < 9; ++k) s += fabs((double)h9[k]);
    return 2e-6 * s;
}
static inline float filter_threshold(double d2_max, double d_abs_err)
{
    double c = sqrt(d2_max > 0.0 ? d2_max : 0.0) + d_abs_err;
    double t = c * c * (1.0 + 1e-5) + 1e-30;
    float f = (float)t;
    if ((double)f < t) f = nextafterf(f, INFINITY);
    return f;
}
// below this FP32 value a pair is certainly inside d2_max
static inline float certain_threshold(double d2_max, double d_abs_err)
{
    double c = sqrt(d2_max > 0.0 ? d2_max : 0.0) - d_abs_err;
    if (c <= 0.0) return 0.0f;
    double t = c * c * (1.0 - 1e-5);
    float f = (float)t;
    if ((double)f >= t) f = nextafterf(f, 0.0f);
    return f;
}

// ---- traversal ---------------------------------------------------------------------
struct SearchParams {
    double d2_max;       // keep d2 <= d2_max (exact, double)
    double d2_min;       // and d2 > d2_min (set to -1 to disable)
    float d2_filter;     // FP32 rejection threshold (>= d2_max * (1 + band))
    int use_filter;      // 0 for flavours without a safe FP32 bound
    int stencil;         // 1: 27/13-cell stencil; 0: single cell (all pairs)
    int include_self;    // SELF traversals: also report the pair (i, i)  (capped_distance(x, x))
};

// One thread per query, in cell-sorted order.  `Sink` receives every accepted
// pair: sink.hit(query_index, candidate_index, d2, frame).
//   SELF = false : queries `q` (sorted list QL) against candidates CL, 27 cells
//   SELF = true  : QL == CL; in-cell pairs with a later slot + 13-cell half shell
template <int MIC, bool SELF, class Sink>
__global__ void __launch_bounds__(Sink::THREADS)
pair_search_kernel(const float4 *__restrict__ q_atoms, const uint32_t *__restrict__ q_keys,
                   uint64_t nq_per_frame, const float4 *__restrict__ c_atoms,
                   const uint32_t *__restrict__ c_start, const FrameGeom *__restrict__ fgs,
                   int per_frame, SearchParams sp, Sink sink)
{
    __shared__ FrameGeom sfg;
    const uint64_t f = blockIdx.y;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(fgs + (per_frame ? f : 0));
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sfg);
        for (int w = threadIdx.x; w < (int)(sizeof(FrameGeom) / 4); w += blockDim.x) dst[w] = src[w];
    }
    sink.begin();
    __syncthreads();
    const FrameGeom &fg = sfg;
    const BoxParams box = sfg.box;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nq_per_frame) {
        const float4 q = q_atoms[f * nq_per_frame + t];
        const int qi = __float_as_int(q.w);
        const int64_t base = per_frame ? fg.cell_base : (int64_t)f * fg.grid.ncell_total;
        int cx = 0, cy = 0, cz = 0;
        if (sp.stencil) {
            int64_t cell = (int64_t)q_keys[f * nq_per_frame + t] - base;
            cx = (int)(cell % fg.grid.ncells[0]);
            int64_t r = cell / fg.grid.ncells[0];
            cy = (int)(r % fg.grid.ncells[1]);
            cz = (int)(r / fg.grid.ncells[1]);
        }
        // SELF: slot 0 is the home cell (later atoms only), then the 13-cell half
        // shell in the reference's route order (nsgrid.pyx:603-605).
        const int nsten = sp.stencil ? (SELF ? 14 : 27) : 1;
        for (int s = 0; s < nsten; ++s) {
            int64_t cell;
            if (!sp.stencil) {
                cell = 0;
            } else if (SELF) {
                unsigned r = kSelfRoute[s];
                cell = cell_neighbor(fg.grid, cx + (int)(r & 3u) - 1, cy + (int)((r >> 2) & 3u) - 1,
                                     cz + (int)((r >> 4) & 3u) - 1);
            } else {
                cell = cell_neighbor(fg.grid, cx + s / 9 - 1, cy + (s / 3) % 3 - 1, cz + s % 3 - 1);
            }
            if (cell < 0) continue;
            uint32_t j0 = c_start[base + cell], j1 = c_start[base + cell + 1];
            if (SELF && s == 0) j0 = (uint32_t)(f * nq_per_frame + t) + 1;   // later slots of the home cell
            for (uint32_t j = j0; j < j1; ++j) {
                const float4 c = c_atoms[j];
                const float tx = q.x - c.x, ty = q.y - c.y, tz = q.z - c.z;   // a - b (nsgrid.pyx:486-488)
                if (sp.use_filter) {
                    float d2f = pair_d2_filter<MIC>(tx, ty, tz, box);
                    if (d2f > sp.d2_filter) continue;
                    if (Sink::FAST && sink.try_fast(qi, __float_as_int(c.w), d2f)) continue;
                }
                const double d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, c.x, c.y, c.z, box);
                if (d2 <= sp.d2_max && d2 > sp.d2_min) sink.hit(qi, __float_as_int(c.w), d2, f);
            }
        }
    }
    sink.end();
}

// ---- tiled traversal: one CTA per home cell --------------------------------------------
// For dense cells (tens of atoms or more, e.g. the RDF workload) the candidates of a
// neighbour cell are staged once per CTA in shared memory with coalesced 16-byte
// loads and every thread (one query of the home cell each) walks the tile with
// warp-broadcast LDS.128 reads: no per-lane global-load latency, uniform trip
// counts, no divergence.  blockIdx.x = cell * chunks + chunk, blockIdx.y = frame.
//   SELF = true : the home cell contributes only later slots and the 13-cell half
//                 shell is visited, i.e. every unordered pair is seen once; the sink
//                 is told via hit_sym()/try_fast_sym() that the pair stands for both
//                 orders (InterRDF of a group with itself, self_search, contacts).
constexpr int TILE_ATOMS = 256;

template <int MIC, bool SELF, class Sink>
__global__ void __launch_bounds__(Sink::THREADS)
cell_tile_kernel(const float4 *__restrict__ q_atoms, const uint32_t *__restrict__ q_start,
                 const float4 *__restrict__ c_atoms, const uint32_t *__restrict__ c_start,
                 const FrameGeom *__restrict__ fgs, int per_frame, int chunks, SearchParams sp,
                 Sink sink)
{
    constexpr int T = Sink::THREADS;
    __shared__ FrameGeom sfg;
    __shared__ float4 tile[TILE_ATOMS];
    const uint64_t f = blockIdx.y;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(fgs + (per_frame ? f : 0));
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sfg);
        for (int w = threadIdx.x; w < (int)(sizeof(FrameGeom) / 4); w += T) dst[w] = src[w];
    }
    sink.begin();
    __syncthreads();
    const int64_t cell = blockIdx.x / chunks;
    if (cell >= sfg.grid.ncell_total) return;                 // uniform: frames may have fewer cells
    const int chunk0 = blockIdx.x % chunks;
    const BoxParams box = sfg.box;
    const int64_t base = per_frame ? sfg.cell_base : (int64_t)f * sfg.grid.ncell_total;
    int cx = 0, cy = 0, cz = 0;
    if (sp.stencil) {
        cx = (int)(cell % sfg.grid.ncells[0]);
        const int64_t r = cell / sfg.grid.ncells[0];
        cy = (int)(r % sfg.grid.ncells[1]);
        cz = (int)(r / sfg.grid.ncells[1]);
    }
    const uint32_t q0 = q_start[base + cell], q1 = q_start[base + cell + 1];
    const int nsten = sp.stencil ? (SELF ? 14 : 27) : 1;
    for (uint32_t qbase = q0 + (uint32_t)chunk0 * T; qbase < q1; qbase += (uint32_t)chunks * T) {
        const uint32_t qslot = qbase + threadIdx.x;
        const bool active = qslot < q1;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) q = q_atoms[qslot];
        const int qi = __float_as_int(q.w);
        if (SELF && sp.include_self && active) {
            const double d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, q.x, q.y, q.z, box);
            if (d2 <= sp.d2_max && d2 > sp.d2_min) sink.template hit_w<SELF>(qi, qi, d2, f);
        }
        for (int s = 0; s < nsten; ++s) {
            int64_t ncell;
            if (!sp.stencil) ncell = 0;
            else if (SELF) {
                const unsigned r = kSelfRoute[s];
                ncell = cell_neighbor(sfg.grid, cx + (int)(r & 3u) - 1, cy + (int)((r >> 2) & 3u) - 1,
                                      cz + (int)((r >> 4) & 3u) - 1);
            } else {
                ncell = cell_neighbor(sfg.grid, cx + s / 9 - 1, cy + (s / 3) % 3 - 1, cz + s % 3 - 1);
            }
            if (ncell < 0) continue;
            uint32_t j0 = c_start[base + ncell];
            const uint32_t j1 = c_start[base + ncell + 1];
            // home cell of a SELF search: only slots after the first query of this chunk matter
            if (SELF && s == 0 && j0 < qbase + 1) j0 = qbase + 1;
            for (uint32_t tb = j0; tb < j1; tb += TILE_ATOMS) {
                const int tn = (int)min((uint32_t)TILE_ATOMS, j1 - tb);
                __syncthreads();
                for (int k = threadIdx.x; k < tn; k += T) tile[k] = c_atoms[tb + k];
                __syncthreads();
                if (!active) continue;
                int k = 0;
                if (SELF && s == 0 && qslot + 1 > tb) k = (int)(qslot + 1 - tb);   // later slots only
#pragma unroll 4
                for (; k < tn; ++k) {
                    const float4 c = tile[k];
                    const float tx = q.x - c.x, ty = q.y - c.y, tz = q.z - c.z;     // a - b (nsgrid.pyx:486-488)
                    if (sp.use_filter) {
                        const float d2f = pair_d2_filter<MIC>(tx, ty, tz, box);
                        if (d2f > sp.d2_filter) continue;
                        if (Sink::FAST && sink.template try_fast_w<SELF>(qi, __float_as_int(c.w), d2f)) continue;
                    }
                    const double d2 = pair_d2_exact<MIC>(q.x, q.y, q.z, c.x, c.y, c.z, box);
                    if (d2 <= sp.d2_max && d2 > sp.d2_min) sink.template hit_w<SELF>(qi, __float_as_int(c.w), d2, f);
                }
            }
        }
    }
    sink.end();
}

// ---- sinks -----------------------------------------------------------------------------
// Appends (i, j, sqrt(d2)) with one atomic per converged group of lanes.
struct PairSink {
    static constexpr int THREADS = 128;
    long long *pairs;                 // [capacity][2]
    double *dist;                     // [capacity]
    unsigned long long *counter;      // total hits (keeps counting past capacity)
    unsigned long long capacity;
    int orient;                       // 1: emit (min(i,j), max(i,j))  (brute-force self variant)
    static constexpr bool FAST = false;
    __device__ __forceinline__ bool try_fast(int, int, float) { return false; }
    template <bool SYM> __device__ __forceinline__ bool try_fast_w(int, int, float) { return false; }
    template <bool SYM> __device__ __forceinline__ void hit_w(int i, int j, double d2, uint64_t f) { hit(i, j, d2, f); }
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void end() {}
    __device__ __forceinline__ void put(unsigned long long slot, int i, int j, double d2)
    {
        if (slot < capacity) {
            longlong2 p;
            p.x = i; p.y = j;
            if (orient && i > j) { p.x = j; p.y = i; }
            *reinterpret_cast<longlong2 *>(pairs + 2 * slot) = p;
            dist[slot] = sqrt(d2);
        }
    }
    // two candidates per lane, all 32 lanes converged: one atomic for both rounds
    __device__ __forceinline__ void hit2(bool h0, int i0, int j0, double d0, bool h1, int i1, int j1, double d1)
    {
        const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
        const int n0 = __popc(m0), n1 = __popc(m1);
        if (n0 + n1 == 0) return;
        const int lane = threadIdx.x & 31;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)(n0 + n1));
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned lt = (1u << lane) - 1u;
        if (h0) put(base + __popc(m0 & lt), i0, j0, d0);
        if (h1) put(base + n0 + __popc(m1 & lt), i1, j1, d1);
    }
    __device__ __forceinline__ void hit(int i, int j, double d2, uint64_t)
    {
        const unsigned m = __activemask();
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
        base = __shfl_sync(m, base, leader);
        const unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < capacity) {
            longlong2 p;
            p.x = i; p.y = j;
            if (orient && i > j) { p.x = j; p.y = i; }
            *reinterpret_cast<longlong2 *>(pairs + 2 * slot) = p;
            dist[slot] = sqrt(d2);
        }
    }
};

}  // namespace mdagpu
