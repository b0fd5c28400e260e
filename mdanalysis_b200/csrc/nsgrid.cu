// nsgrid.cu -- FastNS on the GPU: grid geometry, counting-sort cell list, cutoff
// pair search with warp-aggregated compaction, brute-force capped distances.
// Reference: package/MDAnalysis/lib/nsgrid.pyx (cited per function) and
// package/MDAnalysis/lib/distances.py:589-673, 1050-1130 (brute force).
#include "rowpairs.cuh"

#include <math.h>
#include <new>

namespace mdagpu {

// ---------------------------------------------------------------------------
// host: grid geometry
// ---------------------------------------------------------------------------
static float sin_deg_f32(float deg)
{
    // nsgrid.pyx:119-122: float32 angle in radians, libm double sine, float32 result
    deg = (float)((double)deg * (M_PI / 180.));
    return (float)sin((double)deg);
}

void grid_geometry_reference(double cutoff_in, const float *box6, const float *h9, int pbc,
                             GridGeom *g, float *max_cutoff)
{
    // nsgrid.pyx:294-387
    float cutoff = (float)(cutoff_in > 1.0 ? cutoff_in : 1.0);
    g->triclinic = (h9[3] != 0.0f || h9[6] != 0.0f || h9[7] != 0.0f);
    g->pbc = pbc;
    const float sa = sin_deg_f32(box6[3]), sb = sin_deg_f32(box6[4]), sg = sin_deg_f32(box6[5]);
    float mc = h9[4] * sa;
    mc = fminf(mc, h9[8] * sa);
    mc = fminf(mc, h9[0] * sb);
    mc = fminf(mc, h9[8] * sb);
    mc = fminf(mc, h9[0] * sg);
    mc = fminf(mc, h9[4] * sg);
    mc /= 2;
    if (max_cutoff) *max_cutoff = mc;

    float min_cell = cutoff;
    if (g->triclinic) {
        const float s[3] = { sa, sb, sg };
        for (int k = 0; k < 3; ++k) {
            float c = cutoff / s[k];
            if (c > min_cell) min_cell = c;
        }
    }
    min_cell = (float)((double)min_cell + 0.001);
    const float edge[3] = { h9[0], h9[4], h9[8] };
    for (int k = 0; k < 3; ++k) {
        double nc = floor((double)(edge[k] / min_cell));
        g->ncells[k] = (int)(nc < 1290.0 ? nc : 1290.0);
        if (pbc) {
            if (g->ncells[k] <= 3) { g->ncells[k] = 1; g->periodic[k] = 0; }
            else g->periodic[k] = 1;
        } else {
            if (g->ncells[k] <= 2) g->ncells[k] = 1;
            g->periodic[k] = 0;
        }
    }
    for (int k = 0; k < 9; ++k) g->cs[k] = 0.0;
    g->origin[0] = g->origin[1] = g->origin[2] = 0.0;
    g->cs[0] = (double)h9[0] / (double)g->ncells[0];
    g->cs[3] = (double)(h9[3] / h9[4]);
    g->cs[4] = (double)h9[4] / (double)g->ncells[1];
    g->cs[6] = (double)(h9[6] / h9[8]);
    g->cs[7] = (double)(h9[7] / h9[8]);
    g->cs[8] = (double)h9[8] / (double)g->ncells[2];
    g->ncell_total = (int64_t)g->ncells[0] * g->ncells[1] * g->ncells[2];
}

void grid_geometry_single(GridGeom *g)
{
    for (int k = 0; k < 3; ++k) { g->ncells[k] = 1; g->periodic[k] = 0; }
    for (int k = 0; k < 9; ++k) g->cs[k] = 0.0;
    g->origin[0] = g->origin[1] = g->origin[2] = 0.0;
    g->cs[0] = g->cs[4] = g->cs[8] = INFINITY;   // every coordinate maps to cell 0
    g->triclinic = 0;
    g->pbc = 0;
    g->ncell_total = 1;
}

// ---------------------------------------------------------------------------
// device: cell keys, scan, scatter
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cell_key_kernel(const float *__restrict__ xyz, uint64_t n, uint64_t nframes,
                const FrameGeom *__restrict__ fgs, int per_frame, int wrap,
                uint32_t *__restrict__ keys, uint32_t *__restrict__ counts, uint32_t *__restrict__ ranks,
                unsigned *__restrict__ flag)
{
    // (n * nframes < 2^32, checked by build_cell_list: 32-bit index arithmetic, no 64-bit division)
    const uint32_t total = (uint32_t)(n * nframes), n32 = (uint32_t)n;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t64 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t64 < total; t64 += stride) {
        const uint32_t t = (uint32_t)t64;
        const uint32_t f = t / n32;
        const FrameGeom &fg = fgs[per_frame ? f : 0];
        float x = xyz[3 * (uint64_t)t], y = xyz[3 * (uint64_t)t + 1], z = xyz[3 * (uint64_t)t + 2];
        if (wrap) wrap_any(x, y, z, fg.box);
        int64_t cell = 0;
        if (fg.grid.ncell_total > 1) {
            int c[3];
            if (!cell_of_checked(fg.grid, x, y, z, c)) atomicOr(flag, 1u);
            // the reference's C remainder can only leave [0,ncells) for positions
            // outside the cell; clamp so that a bad input cannot corrupt memory
            for (int k = 0; k < 3; ++k) c[k] = c[k] < 0 ? 0 : c[k];
            cell = cell_linear(fg.grid, c);
        }
        const int64_t base = per_frame ? fg.cell_base : (int64_t)f * fg.grid.ncell_total;
        const uint32_t key = (uint32_t)(base + cell);
        keys[t] = key;
        // the count before this atom is its rank in the cell: the scatter needs no second atomic
        ranks[t] = atomicAdd(&counts[key], 1u);
    }
}

// One CTA per frame (1024 threads, chunked with a running carry): every frame holds
// exactly n atoms, so the prefix of frame f starts at f * n and frames are independent.
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ starts,
                      const FrameGeom *__restrict__ fgs, int per_frame, uint64_t n_atoms, uint64_t nframes)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t f = blockIdx.x;
    const FrameGeom &fg = fgs[per_frame ? f : 0];
    const uint64_t n = (uint64_t)fg.grid.ncell_total;
    const uint64_t cell0 = per_frame ? (uint64_t)fg.cell_base : f * n;
    counts += cell0;
    starts += cell0;
    if (threadIdx.x == 0) carry = (uint32_t)(f * n_atoms);
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t v = i < n ? counts[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += u;
            }
            warp_sums[lane] = wi - w;   // exclusive prefix of the warp totals
        }
        __syncthreads();
        const uint32_t c = carry;
        const uint32_t excl = c + warp_sums[warp] + incl - v;
        if (i < n) starts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && f + 1 == nframes) starts[n] = carry;
}

__global__ void __launch_bounds__(256)
cell_scatter_kernel(const float *__restrict__ xyz, uint64_t n, uint64_t nframes,
                    const FrameGeom *__restrict__ fgs, int per_frame, int wrap,
                    const uint32_t *__restrict__ keys, const uint32_t *__restrict__ starts,
                    const uint32_t *__restrict__ ranks, float4 *__restrict__ sorted,
                    uint32_t *__restrict__ sorted_key)
{
    const uint32_t total = (uint32_t)(n * nframes), n32 = (uint32_t)n;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t64 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t64 < total; t64 += stride) {
        const uint32_t t = (uint32_t)t64;
        const uint32_t f = t / n32;
        const FrameGeom &fg = fgs[per_frame ? f : 0];
        float x = xyz[3 * (uint64_t)t], y = xyz[3 * (uint64_t)t + 1], z = xyz[3 * (uint64_t)t + 2];
        if (wrap == 1) wrap_any(x, y, z, fg.box);   // wrap == 2: binned wrapped, stored as given
        const uint32_t key = keys[t];
        const uint32_t slot = starts[key] + ranks[t];
        sorted[slot] = make_float4(x, y, z, __int_as_float((int)(t - f * n32)));
        if (sorted_key) sorted_key[slot] = key;     // only the thread-per-query kernels read it
    }
}

int build_cell_list(const float *xyz, uint64_t n, uint64_t nframes, const FrameGeom *fg_dev,
                    int per_frame, uint64_t total_cells, int wrap, cudaStream_t st, CellList *cl, bool want_sorted_key)
{
    const uint64_t total = n * nframes;
    MDA_REQUIRE(total < 0xffffffffull && total_cells < 0xfffffff0ull,
                "cell list too large for 32-bit keys (%llu atoms, %llu cells)",
                (unsigned long long)total, (unsigned long long)total_cells);
    cl->n = n; cl->nframes = nframes; cl->total_cells = total_cells;
    MDA_CUDA(cl->keys.alloc(total * sizeof(uint32_t), st));
    MDA_CUDA(cl->counts.alloc(total_cells * sizeof(uint32_t), st));
    MDA_CUDA(cl->starts.alloc((total_cells + 1) * sizeof(uint32_t), st));
    MDA_CUDA(cl->cursor.alloc((total ? total : 1) * sizeof(uint32_t), st));       // rank of every atom in its cell
    MDA_CUDA(cl->sorted.alloc(total * sizeof(float4), st));
    if (want_sorted_key) MDA_CUDA(cl->sorted_key.alloc(total * sizeof(uint32_t), st));
    MDA_CUDA(cl->flag.alloc(sizeof(unsigned), st));
    MDA_CUDA(cudaMemsetAsync(cl->flag.p, 0, sizeof(unsigned), st));
    MDA_CUDA(cudaMemsetAsync(cl->counts.p, 0, total_cells * sizeof(uint32_t), st));
    if (total == 0) {
        MDA_CUDA(cudaMemsetAsync(cl->starts.p, 0, (total_cells + 1) * sizeof(uint32_t), st));
        return MDAGPU_OK;
    }
    unsigned grid = cdiv(total, 256);
    const unsigned cap = (unsigned)num_sms() * 16;
    if (grid > cap) grid = cap;
    MDA_LAUNCH(cell_key_kernel, grid, 256, 0, st, xyz, n, nframes, fg_dev, per_frame, wrap,
               cl->keys.as<uint32_t>(), cl->counts.as<uint32_t>(), cl->cursor.as<uint32_t>(), cl->flag.as<unsigned>());
    MDA_CHECK_LAUNCH();
    MDA_LAUNCH(exclusive_scan_kernel, (unsigned)nframes, 1024, 0, st, cl->counts.as<uint32_t>(),
               cl->starts.as<uint32_t>(), fg_dev, per_frame, n, nframes);
    MDA_CHECK_LAUNCH();
    MDA_LAUNCH(cell_scatter_kernel, grid, 256, 0, st, xyz, n, nframes, fg_dev, per_frame, wrap,
               cl->keys.as<uint32_t>(), cl->starts.as<uint32_t>(), cl->cursor.as<uint32_t>(),
               cl->sorted.as<float4>(), want_sorted_key ? cl->sorted_key.as<uint32_t>() : nullptr);
    MDA_CHECK_LAUNCH();
    return MDAGPU_OK;
}

// ---------------------------------------------------------------------------
// pair list container
// ---------------------------------------------------------------------------
}  // namespace mdagpu

// Result buffers come from the stream-ordered pool (kept cached between calls, common.cuh):
// a fresh cudaMalloc of the GB-sized lists of a 1 M-atom search costs more than the search.
struct mdagpu_pairs {
    long long *pairs = nullptr;     // device [count][2]
    double *dist = nullptr;         // device [count]
    unsigned long long count = 0, capacity = 0;
    cudaStream_t stream = nullptr;
    void release()
    {
        if (pairs) cudaFreeAsync(pairs, stream);
        if (dist) cudaFreeAsync(dist, stream);
        pairs = nullptr;
        dist = nullptr;
    }
    ~mdagpu_pairs() { release(); }
};

struct mdagpu_grid {
    mdagpu::FrameGeom fg;
    mdagpu::FrameGeom *fg_dev = nullptr;
    mdagpu::CellList cl;
    double cutoff = 0.0;
    float max_cutoff = 0.0f;
    int mic = 0;
    uint64_t n = 0;
    ~mdagpu_grid()
    {
        if (fg_dev) cudaFree(fg_dev);
    }
};

namespace mdagpu {

// How the search walks the grid: the row-staged kernel (rowpairs.cuh) when the per-cell-pair
// image prediction holds for this geometry, else one thread per query with global loads.
struct PairTraversal {
    int row = 0;
    RowParams rp = { 1, 1 };
    uint64_t grid_x = 0;
    int cap = RP_ALL;          // plane capacity of the row kernel (candidates): sizes its shared memory
};

static PairTraversal choose_traversal(int mic, const FrameGeom &fg, double cutoff, uint64_t n_queries, bool consistent)
{
    PairTraversal tv;
    if (mic == MIC_ORTHO_ROUND || !consistent || fg.grid.ncell_total < 8) return tv;
    if (!rowtile_supported(mic, fg.grid, fg.box, cutoff, 1)) return tv;
    const double occupancy = (double)n_queries / (double)fg.grid.ncell_total;
    const int nx = fg.grid.ncells[0];
    int R = (int)(128.0 / fmax(occupancy, 1e-9));
    R = R < 1 ? 1 : (R > ROW_MAXR ? ROW_MAXR : R);
    if (R > nx) R = nx;
    const int runs = (nx + R - 1) / R;
    tv.rp.W = 1;
    tv.rp.R = (nx + runs - 1) / runs;
    tv.grid_x = (uint64_t)runs * (uint64_t)fg.grid.ncells[1] * (uint64_t)fg.grid.ncells[2];
    tv.row = 1;
    // Tile capacity: the CTA stages (2W + 1)^2 rows (the half shell of a self search: the home row and
    // W + (2W + 1) W more) of R + 2W cells; 1.5 x the mean leaves room for the spread of the cell
    // occupancies (a CTA whose rows do not fit stages them one at a time).  The resident CTAs per SM
    // follow from this number: 1 M atoms at 0.1 / A^3, 8 A: 1856 candidates = 22 kB (the fixed 43 kB
    // allowed 4 CTAs of 4 warps).
    const double rows = (double)((2 * tv.rp.W + 1) * (2 * tv.rp.W + 1));
    const double mean = rows * (double)(tv.rp.R + 2 * tv.rp.W) * occupancy;
    int cap = (int)(1.5 * mean) + 64;
    cap = (cap + 63) & ~63;
    tv.cap = cap < RP_MINCAP ? RP_MINCAP : (cap > RP_ALL ? RP_ALL : cap);
    return tv;
}

template <int MIC, bool SELF>
static int run_pair_search(const CellList &ql, uint64_t nq, const CellList &cl,
                           const FrameGeom *fg_dev, const SearchParams &sp, uint64_t guess,
                           int orient, const PairTraversal &tv, cudaStream_t st, mdagpu_pairs *res)
{
    DevBuf counter;
    MDA_CUDA(counter.alloc(sizeof(unsigned long long), st));
    unsigned long long cap = guess < 1024 ? 1024 : guess;
    for (int attempt = 0; attempt < 2; ++attempt) {
        res->release();
        res->stream = st;
        pool_init();
        MDA_CUDA(cudaMallocAsync(&res->pairs, cap * 2 * sizeof(long long), st));
        MDA_CUDA(cudaMallocAsync(&res->dist, cap * sizeof(double), st));
        res->capacity = cap;
        MDA_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), st));
        PairSink sink = { res->pairs, res->dist, counter.as<unsigned long long>(), cap, orient };
        if constexpr (MIC != MIC_ORTHO_ROUND) if (tv.row) {
            dim3 grid((unsigned)tv.grid_x, 1);
            // (a self search stages the half shell, about half of what tv.cap was sized for: the same
            // number is used for both so that one attribute setting serves the kernel)
            const int cap = SELF ? (((tv.cap * 5 / 9 + 63) & ~63) < RP_MINCAP ? RP_MINCAP : ((tv.cap * 5 / 9 + 63) & ~63)) : tv.cap;
            MDA_CUDA(cudaFuncSetAttribute(row_pairs_kernel<MIC, SELF, PairSink>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rp_dyn_bytes(RP_ALL)));
            MDA_LAUNCH_AS(SELF ? "pair_self_search" : "pair_search", (row_pairs_kernel<MIC, SELF, PairSink>), grid, RP_THREADS, rp_dyn_bytes(cap), st,
                       ql.atoms(), ql.start(), cl.atoms(), cl.start(), fg_dev, 0, tv.rp, sp, sink, cap);
        }
        if (!tv.row || MIC == MIC_ORTHO_ROUND) {
            dim3 grid(cdiv(nq, PairSink::THREADS), 1);
            MDA_LAUNCH_AS(SELF ? "pair_self_search" : "pair_search", (pair_search_kernel<MIC, SELF, PairSink>), grid, PairSink::THREADS, 0, st,
                       ql.atoms(), ql.key_of_sorted(), nq, cl.atoms(), cl.start(), fg_dev, 0, sp, sink);
        }
        MDA_CHECK_LAUNCH();
        unsigned long long found = 0;
        MDA_CUDA(cudaMemcpyAsync(&found, counter.p, sizeof(found), cudaMemcpyDeviceToHost, st));
        MDA_CUDA(cudaStreamSynchronize(st));
        res->count = found;
        if (found <= cap) return MDAGPU_OK;
        cap = found;    // exact size known now: second pass cannot overflow
    }
    return set_error(MDAGPU_ERR_CUDA, "pair search overflowed twice");
}

template <bool SELF>
static int dispatch_pair_search(int mic, const CellList &ql, uint64_t nq, const CellList &cl,
                                const FrameGeom *fg_dev, const SearchParams &sp, uint64_t guess,
                                int orient, const PairTraversal &tv, cudaStream_t st, mdagpu_pairs *res)
{
    switch (mic) {
    case MIC_NONE: return run_pair_search<MIC_NONE, SELF>(ql, nq, cl, fg_dev, sp, guess, orient, tv, st, res);
    case MIC_ORTHO_ROUND: return run_pair_search<MIC_ORTHO_ROUND, SELF>(ql, nq, cl, fg_dev, sp, guess, orient, tv, st, res);
    case MIC_ORTHO_LAZY: return run_pair_search<MIC_ORTHO_LAZY, SELF>(ql, nq, cl, fg_dev, sp, guess, orient, tv, st, res);
    case MIC_TRICLINIC: return run_pair_search<MIC_TRICLINIC, SELF>(ql, nq, cl, fg_dev, sp, guess, orient, tv, st, res);
    }
    return set_error(MDAGPU_ERR_INVALID, "unknown minimum-image flavour %d", mic);
}

// expected number of hits per query from the mean density, for the first
// allocation (overflow is handled by a second, exactly sized pass)
static uint64_t guess_pairs(uint64_t nq, uint64_t nc, double cutoff, double volume)
{
    double per_query = nc;
    if (volume > 0.0) {
        double v = 4.0 / 3.0 * M_PI * cutoff * cutoff * cutoff / volume * (double)nc;
        if (v < per_query) per_query = v;
    }
    double g = (double)nq * per_query * 1.25 + 4096.0;
    if (g > 4.0e9) g = 4.0e9;
    return (uint64_t)g;
}

}  // namespace mdagpu

using namespace mdagpu;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" int mdagpu_grid_geometry(double cutoff, const float *box6, const float *h9, int pbc,
                                    int *ncells3, double *cellsize9, int *periodic3,
                                    int *triclinic, float *max_cutoff)
{
    MDA_REQUIRE(box6 && h9, "mdagpu_grid_geometry: box is NULL");
    GridGeom g;
    float mc = 0.0f;
    grid_geometry_reference(cutoff, box6, h9, pbc, &g, &mc);
    if (ncells3) memcpy(ncells3, g.ncells, sizeof g.ncells);
    if (cellsize9) memcpy(cellsize9, g.cs, sizeof g.cs);
    if (periodic3) memcpy(periodic3, g.periodic, sizeof g.periodic);
    if (triclinic) *triclinic = g.triclinic;
    if (max_cutoff) *max_cutoff = mc;
    return MDAGPU_OK;
}

extern "C" int mdagpu_grid_create(double cutoff, const float *coords, uint64_t n, const float *box6,
                                  const float *h9, int pbc, int space, mdagpu_stream_t stream,
                                  mdagpu_grid **out)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(out != nullptr, "mdagpu_grid_create: out is NULL");
    *out = nullptr;
    MDA_REQUIRE(box6 && h9, "mdagpu_grid_create: box is NULL");
    // nsgrid.pyx:284-291
    MDA_REQUIRE(!(box6[0] == 0.0f || box6[1] == 0.0f || box6[2] == 0.0f), "Any of the box dimensions cannot be 0");
    MDA_REQUIRE(cutoff >= 0.0, "Cutoff must be positive");
    mdagpu_grid *g = new (std::nothrow) mdagpu_grid;
    if (!g) return set_error(MDAGPU_ERR_NOMEM, "out of host memory");
    g->cutoff = cutoff;
    g->n = n;
    grid_geometry_reference(cutoff, box6, h9, pbc, &g->fg.grid, &g->max_cutoff);
    if (cutoff > (double)g->max_cutoff) {
        float mc = g->max_cutoff;
        delete g;
        return set_error(MDAGPU_ERR_INVALID, "Cutoff %g too large for box (max %g)", cutoff, (double)mc);
    }
    const int tri = g->fg.grid.triclinic;
    // wrapping always happens (nsgrid.pyx:406-414); the minimum image only with pbc
    g->fg.box = tri ? make_box_params(BOX_TRICLINIC, h9) : make_box_params(BOX_ORTHO, box6);
    g->fg.cell_base = 0;
    g->mic = pbc ? (tri ? MIC_TRICLINIC : MIC_ORTHO_LAZY) : MIC_NONE;
    if (g->fg.grid.ncell_total > (int64_t)1 << 28) {
        const long long nc = (long long)g->fg.grid.ncell_total;
        delete g;
        return set_error(MDAGPU_ERR_NOMEM, "grid of %lld cells is too large", nc);
    }
    cudaError_t e = cudaMalloc(&g->fg_dev, sizeof(FrameGeom));
    if (e == cudaSuccess) e = cudaMemcpyAsync(g->fg_dev, &g->fg, sizeof(FrameGeom), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { delete g; MDA_CUDA(e); }
    StagedIn in;
    e = in.init(coords, n * 3 * sizeof(float), space, st);
    if (e != cudaSuccess) { delete g; MDA_CUDA(e); }
    int rc = build_cell_list(in.as<float>(), n, 1, g->fg_dev, 0, (uint64_t)g->fg.grid.ncell_total, 1, st, &g->cl);
    if (rc == MDAGPU_OK) {
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = set_error(MDAGPU_ERR_CUDA, "grid build failed: %s", cudaGetErrorString(e));
    }
    if (rc != MDAGPU_OK) { delete g; return rc; }
    *out = g;
    return MDAGPU_OK;
}

extern "C" int mdagpu_grid_info(const mdagpu_grid *g, int *ncells3, double *cellsize9,
                                int *periodic3, int *triclinic, float *max_cutoff)
{
    MDA_REQUIRE(g != nullptr, "mdagpu_grid_info: grid is NULL");
    if (ncells3) memcpy(ncells3, g->fg.grid.ncells, sizeof g->fg.grid.ncells);
    if (cellsize9) memcpy(cellsize9, g->fg.grid.cs, sizeof g->fg.grid.cs);
    if (periodic3) memcpy(periodic3, g->fg.grid.periodic, sizeof g->fg.grid.periodic);
    if (triclinic) *triclinic = g->fg.grid.triclinic;
    if (max_cutoff) *max_cutoff = g->max_cutoff;
    return MDAGPU_OK;
}

extern "C" void mdagpu_grid_free(mdagpu_grid *g) { delete g; }

// Triclinic cells: an atom that the float64 cell arithmetic folded into a cell it does not lie
// in (flagged by cell_key_kernel) breaks the per-cell image prediction of the row kernel.
static int cells_consistent(int mic, const CellList &cl, const CellList *ql, cudaStream_t st, bool *ok)
{
    *ok = true;
    if (mic != MIC_TRICLINIC) return MDAGPU_OK;
    unsigned bad[2] = { 0u, 0u };
    MDA_CUDA(cudaMemcpyAsync(&bad[0], cl.inconsistent(), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    if (ql) MDA_CUDA(cudaMemcpyAsync(&bad[1], ql->inconsistent(), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    MDA_CUDA(cudaStreamSynchronize(st));
    *ok = !(bad[0] | bad[1]);
    return MDAGPU_OK;
}

static double grid_volume(const mdagpu_grid *g)
{
    const GridGeom &gg = g->fg.grid;
    return gg.cs[0] * gg.ncells[0] * gg.cs[4] * gg.ncells[1] * gg.cs[8] * gg.ncells[2];
}

extern "C" int mdagpu_grid_search(mdagpu_grid *g, const float *queries, uint64_t nq, int space,
                                  mdagpu_stream_t stream, mdagpu_pairs **out)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(g && out, "mdagpu_grid_search: NULL argument");
    *out = nullptr;
    mdagpu_pairs *res = new (std::nothrow) mdagpu_pairs;
    if (!res) return set_error(MDAGPU_ERR_NOMEM, "out of host memory");
    if (nq == 0 || g->n == 0) { *out = res; return MDAGPU_OK; }
    StagedIn in;
    cudaError_t e = in.init(queries, nq * 3 * sizeof(float), space, st);
    if (e != cudaSuccess) { delete res; MDA_CUDA(e); }
    CellList ql;   // queries are wrapped and binned exactly like grid atoms (nsgrid.pyx:551-560)
    int rc = build_cell_list(in.as<float>(), nq, 1, g->fg_dev, 0, (uint64_t)g->fg.grid.ncell_total, 1, st, &ql);
    if (rc == MDAGPU_OK) {
        SearchParams sp;
        sp.d2_max = g->cutoff * g->cutoff;   // nsgrid.pyx:540
        sp.d2_min = -1.0;
        sp.d2_filter = filter_threshold(sp.d2_max, g->mic == MIC_TRICLINIC ? tric_abs_err(g->fg.box.h) : 0.0);
        sp.use_filter = 1;
        sp.stencil = 1;
        sp.include_self = 0;
        bool consistent = true;
        rc = cells_consistent(g->mic, g->cl, &ql, st, &consistent);
        if (rc == MDAGPU_OK) {
            // the row kernel filters with its own, tighter error bound (rowtile_abs_err)
            const PairTraversal tv = choose_traversal(g->mic, g->fg, g->cutoff, nq, consistent);
            if (tv.row) sp.d2_filter = filter_threshold(sp.d2_max, rowtile_abs_err(g->mic, g->fg.box, g->cutoff));
            rc = dispatch_pair_search<false>(g->mic, ql, nq, g->cl, g->fg_dev, sp,
                                             guess_pairs(nq, g->n, g->cutoff, grid_volume(g)), 0, tv, st, res);
        }
    }
    if (rc != MDAGPU_OK) { delete res; return rc; }
    *out = res;
    return MDAGPU_OK;
}

extern "C" int mdagpu_grid_self_search(mdagpu_grid *g, mdagpu_stream_t stream, mdagpu_pairs **out)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(g && out, "mdagpu_grid_self_search: NULL argument");
    *out = nullptr;
    mdagpu_pairs *res = new (std::nothrow) mdagpu_pairs;
    if (!res) return set_error(MDAGPU_ERR_NOMEM, "out of host memory");
    if (g->n < 2) { *out = res; return MDAGPU_OK; }
    SearchParams sp;
    sp.d2_max = g->cutoff * g->cutoff;       // nsgrid.pyx:601
    sp.d2_min = -1.0;
    sp.d2_filter = filter_threshold(sp.d2_max, g->mic == MIC_TRICLINIC ? tric_abs_err(g->fg.box.h) : 0.0);
    sp.use_filter = 1;
    sp.stencil = 1;
    sp.include_self = 0;
    bool consistent = true;
    int rc = cells_consistent(g->mic, g->cl, nullptr, st, &consistent);
    if (rc == MDAGPU_OK) {
        const PairTraversal tv = choose_traversal(g->mic, g->fg, g->cutoff, g->n, consistent);
        if (tv.row) sp.d2_filter = filter_threshold(sp.d2_max, rowtile_abs_err(g->mic, g->fg.box, g->cutoff));
        rc = dispatch_pair_search<true>(g->mic, g->cl, g->n, g->cl, g->fg_dev, sp,
                                        guess_pairs(g->n, g->n, g->cutoff, grid_volume(g)) / 2 + 4096, 0, tv, st, res);
    }
    if (rc != MDAGPU_OK) { delete res; return rc; }
    *out = res;
    return MDAGPU_OK;
}

extern "C" uint64_t mdagpu_pairs_count(const mdagpu_pairs *p) { return p ? p->count : 0; }

extern "C" int mdagpu_pairs_export(const mdagpu_pairs *p, int64_t *pairs, double *distances,
                                   int space, mdagpu_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(p != nullptr, "mdagpu_pairs_export: NULL pairs");
    if (p->count == 0) return MDAGPU_OK;
    if (space == MDAGPU_DEVICE) {
        if (pairs) MDA_CUDA(cudaMemcpyAsync(pairs, p->pairs, p->count * 2 * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        if (distances) MDA_CUDA(cudaMemcpyAsync(distances, p->dist, p->count * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return MDAGPU_OK;
    }
    // the host arrays are NumPy's (pageable): staged through the pinned ring (hostcopy.cu)
    if (pairs) MDA_CUDA(device_to_host(pairs, p->pairs, p->count * 2 * sizeof(int64_t), st));
    if (distances) MDA_CUDA(device_to_host(distances, p->dist, p->count * sizeof(double), st));
    return MDAGPU_OK;
}

extern "C" void mdagpu_pairs_free(mdagpu_pairs *p) { delete p; }

// d <= max_cutoff on the rounded float64 distance  <=>  d2 <= largest double whose
// correctly rounded sqrt is <= max_cutoff (sqrt is monotone).
static double d2_threshold_le(double c)
{
    if (!(c >= 0.0)) return -1.0;
    if (isinf(c)) return c;
    double x = c * c;
    while (sqrt(x) > c) x = nextafter(x, -INFINITY);
    while (sqrt(nextafter(x, INFINITY)) <= c) x = nextafter(x, INFINITY);
    return x;
}

extern "C" int mdagpu_capped_bruteforce(const float *ref, uint64_t n, const float *conf, uint64_t m,
                                        double max_cutoff, double min_cutoff, int box_mode,
                                        const float *box, int space, mdagpu_stream_t stream,
                                        mdagpu_pairs **out)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(out != nullptr, "mdagpu_capped_bruteforce: out is NULL");
    *out = nullptr;
    MDA_REQUIRE(box_mode >= 0 && box_mode <= 2, "bad box_mode %d", box_mode);
    MDA_REQUIRE(box_mode == BOX_NONE || box != nullptr, "box is NULL");
    const bool self = conf == nullptr;
    mdagpu_pairs *res = new (std::nothrow) mdagpu_pairs;
    if (!res) return set_error(MDAGPU_ERR_NOMEM, "out of host memory");
    if (n == 0 || (!self && m == 0) || (self && n < 2)) { *out = res; return MDAGPU_OK; }

    FrameGeom fg;
    grid_geometry_single(&fg.grid);
    fg.box = make_box_params(box_mode, box);
    fg.cell_base = 0;
    StagedIn r, c;
    cudaError_t e = r.init(ref, n * 3 * sizeof(float), space, st);
    if (e == cudaSuccess && !self) e = c.init(conf, m * 3 * sizeof(float), space, st);
    // the triclinic distance kernels work on wrapped copies (calc_distances.h:429-430);
    // every atom carries its original index, so the slot order does not matter.
    int wrap = box_mode == BOX_TRICLINIC;
    // Large inputs (the reference picks 'pkdtree' for them: thousands of atoms, cutoff a few per cent
    // of the box): all-pairs would be O(n m).  Same float64 arithmetic, but the atoms are binned on a
    // grid of cells >= max_cutoff and only the 27-cell stencil is visited -- orthorhombic boxes
    // (binned by the wrapped position, evaluated on the coordinates as given) and no box (grid over
    // the bounding box of both groups).  Triclinic cells keep the single cell.
    int stencil = 0;
    const uint64_t biggest = self ? n : (n > m ? n : m);
    if (e == cudaSuccess && box_mode != BOX_TRICLINIC && biggest >= 8192 && max_cutoff > 0.0 && !getenv("MDAGPU_NO_BRUTE_GRID")) {
        GridGeom g;
        bool ok = true;
        if (box_mode == BOX_ORTHO) {
            grid_geometry_ortho_free(max_cutoff, box, nullptr, 1, &g);
            wrap = 2;
        } else {
            DevBuf lohi;
            e = lohi.alloc(6 * sizeof(float), st);
            const float init[6] = { FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX };
            float h[6];
            if (e == cudaSuccess) e = cudaMemcpyAsync(lohi.p, init, sizeof init, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) {
                unsigned blocks = cdiv(n, 256);
                if (blocks > (unsigned)num_sms() * 8) blocks = (unsigned)num_sms() * 8;
                MDA_LAUNCH(bbox_kernel, blocks, 256, 0, st, r.as<float>(), n, lohi.as<float>());
                if (!self) {
                    blocks = cdiv(m, 256);
                    if (blocks > (unsigned)num_sms() * 8) blocks = (unsigned)num_sms() * 8;
                    MDA_LAUNCH(bbox_kernel, blocks, 256, 0, st, c.as<float>(), m, lohi.as<float>());
                }
                e = cudaMemcpyAsync(h, lohi.p, sizeof h, cudaMemcpyDeviceToHost, st);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e == cudaSuccess) {
                float L[3];
                double origin[3];
                for (int k = 0; k < 3; ++k) {
                    ok = ok && isfinite(h[k]) && isfinite(h[3 + k]);
                    const double pad = 1e-3 + 1e-5 * fmax(fabs((double)h[k]), fabs((double)h[3 + k]));
                    origin[k] = (double)h[k] - pad;
                    L[k] = (float)((double)h[3 + k] - (double)h[k] + 2.0 * pad + max_cutoff * 1e-3);
                }
                if (ok) grid_geometry_ortho_free(max_cutoff, L, origin, 0, &g);
            }
        }
        if (e == cudaSuccess && ok && g.ncell_total >= 27) { fg.grid = g; stencil = 1; }
        else if (box_mode == BOX_ORTHO) wrap = 0;
    }
    DevBuf fg_dev;
    if (e == cudaSuccess) e = fg_dev.alloc(sizeof(FrameGeom), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(fg_dev.p, &fg, sizeof(FrameGeom), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { delete res; MDA_CUDA(e); }
    CellList rl, cl;
    const uint64_t ncell = (uint64_t)fg.grid.ncell_total;
    int rc = build_cell_list(r.as<float>(), n, 1, fg_dev.as<FrameGeom>(), 0, ncell, wrap, st, &rl);
    if (rc == MDAGPU_OK && !self)
        rc = build_cell_list(c.as<float>(), m, 1, fg_dev.as<FrameGeom>(), 0, ncell, wrap, st, &cl);
    if (rc == MDAGPU_OK) {
        SearchParams sp;
        sp.d2_max = d2_threshold_le(max_cutoff);
        sp.d2_min = min_cutoff >= 0.0 ? d2_threshold_le(min_cutoff) : -1.0;
        sp.d2_filter = filter_threshold(sp.d2_max, box_mode == BOX_TRICLINIC ? tric_abs_err(fg.box.h) : 0.0);
        sp.use_filter = box_mode != BOX_ORTHO;   // no safe FP32 bound for unwrapped ortho input
        sp.stencil = stencil;
        sp.include_self = 0;
        const int mic = box_mode == BOX_ORTHO ? MIC_ORTHO_ROUND : (box_mode == BOX_TRICLINIC ? MIC_TRICLINIC : MIC_NONE);
        const uint64_t guess = (self ? n * 32 : n * 64) + 4096;
        if (self) rc = dispatch_pair_search<true>(mic, rl, n, rl, fg_dev.as<FrameGeom>(), sp, guess, 1, PairTraversal(), st, res);
        else rc = dispatch_pair_search<false>(mic, rl, n, cl, fg_dev.as<FrameGeom>(), sp, guess, 0, PairTraversal(), st, res);
    }
    if (rc != MDAGPU_OK) { delete res; return rc; }
    *out = res;
    return MDAGPU_OK;
}
