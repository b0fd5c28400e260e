// rdf.cu -- InterRDF pair histogram (K9) and fused contacts (K2c) on top of the
// pair engine.
// Reference: package/MDAnalysis/analysis/rdf.py:268-291 (capped_distance +
// np.histogram per frame), analysis/contacts.py:315-338 and
// analysis/distances.py:102-108 (contact predicate d <= radius).
//
// Histogram design: every warp owns a private 32-bit histogram in shared memory
// (nbins words per warp) that is updated with shared-memory atomics (ATOMS.ADD, one
// instruction per candidate step, predicated on the lanes that hit); lanes of one
// warp only collide when they fall into the same bin in the same step.  At the end
// of the CTA the warp histograms are summed and merged into the global 64-bit
// histogram with one atomic per non-empty bin.  The FP32 binning uses no
// conversion-unit instruction: floor() is a magic-number add (the XU pipe runs at
// 16 lanes/clk/SM on B200, profiles/microbench/pipes_b200.txt, and is kept for
// the one MUFU.RSQ per step).
#include "rowfine.cuh"

#include <math.h>
#include <stdlib.h>
#include <type_traits>
#include <vector>

namespace mdagpu {

void grid_geometry_ortho_free(double cutoff, const float *L3, const double *origin, int pbc,
                              GridGeom *g)
{
    const double cs_min = cutoff * (1.0 + 1e-4) + 1e-3;
    g->triclinic = 0;
    g->pbc = pbc;
    for (int k = 0; k < 9; ++k) g->cs[k] = 0.0;
    int64_t total = 1;
    for (int k = 0; k < 3; ++k) {
        g->origin[k] = origin ? origin[k] : 0.0;
        double nc = floor((double)L3[k] / cs_min);
        if (!(nc >= 1.0)) nc = 1.0;
        if (nc > 256.0) nc = 256.0;              // keeps the key space small; cells only get larger
        int n = (int)nc;
        if (pbc) {
            if (n < 3) { n = 1; g->periodic[k] = 0; }
            else g->periodic[k] = 1;
        } else {
            g->periodic[k] = 0;
        }
        g->ncells[k] = n;
        g->cs[4 * k] = (double)L3[k] / (double)n;
        total *= n;
    }
    g->ncell_total = total;
}

// ---- triclinic cells: is a (2W + 1)^3 stencil on grid `g` COMPLETE for this cutoff? ----
// cell_of() (geom.cuh, the reference's coord2cellxyz, nsgrid.pyx:435-446) maps a position to the
// cell coordinates  z' = z / cs_zz,  y' = (y - t_yz z) / cs_yy,  x' = (x - t_xy y - t_xz z) / cs_xx
// with the tilt ratios t_xy = b_x / b_y, t_xz = c_x / c_z, t_yz = c_y / c_z.  Two positions within the
// cutoff differ by at most  r_c / cs_zz,  r_c sqrt(1 + t_yz^2) / cs_yy,  r_c sqrt(1 + t_xy^2 + t_xz^2) / cs_xx
// (the maximum of a linear form over a sphere).  A shift by the lattice vectors a, b changes the cell
// coordinates by whole grids; a shift by c moves x' by -(b_x c_y / b_y) / cs_xx as well (the reference
// shears x by t_xy y, not by the fractional coordinate): a pair across the periodic z boundary is that
// much further apart in x'.  The stencil is complete when W cells cover these reaches.  This is what
// decides whether the reference's own search (W = 1 on its grid) returns EVERY pair within the
// cutoff (for [40, 35, 50, 70, 80, 100] it does not) -- only then may a finer
// grid be used for a triclinic InterRDF.
static bool tric_stencil_complete(const GridGeom &g, const float *h9, double cutoff, int W)
{
    const double rc = cutoff * (1.0 + 1e-4) + 1e-4;
    const double fx = sqrt(1.0 + g.cs[3] * g.cs[3] + g.cs[6] * g.cs[6]), fy = sqrt(1.0 + g.cs[7] * g.cs[7]);
    const double dx = fabs((double)h9[3] * (double)h9[7] / (double)h9[4]);
    return (double)W * g.cs[0] >= rc * fx + dx && (double)W * g.cs[4] >= rc * fy && (double)W * g.cs[8] >= rc;
}

// Triclinic grid of the library's own: the reference's cell coordinates (same tilt ratios, so
// cell_of() applies unchanged) with cells just wide enough for a complete (2W + 1)^3 stencil.
static void grid_geometry_triclinic_free(double cutoff, int W, const float *h9, GridGeom *g)
{
    g->triclinic = 1;
    g->pbc = 1;
    for (int k = 0; k < 9; ++k) g->cs[k] = 0.0;
    g->origin[0] = g->origin[1] = g->origin[2] = 0.0;
    g->cs[3] = (double)(h9[3] / h9[4]);
    g->cs[6] = (double)(h9[6] / h9[8]);
    g->cs[7] = (double)(h9[7] / h9[8]);
    const double rc = cutoff * (1.0 + 2e-4) + 1e-3;
    const double fx = sqrt(1.0 + g->cs[3] * g->cs[3] + g->cs[6] * g->cs[6]), fy = sqrt(1.0 + g->cs[7] * g->cs[7]);
    const double dx = fabs((double)h9[3] * (double)h9[7] / (double)h9[4]);
    const double need[3] = { (rc * fx + dx) / W, rc * fy / W, rc / W };
    const double edge[3] = { (double)h9[0], (double)h9[4], (double)h9[8] };
    int64_t total = 1;
    for (int k = 0; k < 3; ++k) {
        double nc = floor(edge[k] / need[k]);
        if (!(nc >= 1.0)) nc = 1.0;
        if (nc > 256.0) nc = 256.0;
        const int n = (int)nc;
        g->ncells[k] = n;
        g->periodic[k] = n >= 3 ? 1 : 0;
        g->cs[4 * k] = edge[k] / (double)n;
        total *= n;
    }
    g->ncell_total = total;
}

// Orthorhombic grid with a different minimum cell width per axis (rowfine.cuh: fine cells
// along x, cutoff / 2 along y and z).
static void grid_geometry_ortho_axes(const double *cutoff3, const float *L3, GridGeom *g)
{
    g->triclinic = 0;
    g->pbc = 1;
    for (int k = 0; k < 9; ++k) g->cs[k] = 0.0;
    int64_t total = 1;
    for (int k = 0; k < 3; ++k) {
        const double cs_min = cutoff3[k] * (1.0 + 1e-4) + 1e-3;
        g->origin[k] = 0.0;
        double nc = floor((double)L3[k] / cs_min);
        if (!(nc >= 1.0)) nc = 1.0;
        if (nc > 256.0) nc = 256.0;
        const int n = (int)nc;
        g->periodic[k] = n >= 3 ? 1 : 0;
        g->ncells[k] = n;
        g->cs[4 * k] = (double)L3[k] / (double)n;
        total *= n;
    }
    g->ncell_total = total;
}

// ---------------------------------------------------------------------------
// histogram sink
// ---------------------------------------------------------------------------
// EXCL: an exclusion block is active (rdf.py:274-278), the weight of a pair then
// depends on both atom indices.
template <int T, bool EXCL>
struct HistSink {
    static constexpr int THREADS = T;
    static constexpr bool FAST = true;
    const double *thr;            // [nbins+1] exact thresholds on d2 (device)
    unsigned long long *counts;   // [nbins] global histogram
    int nbins, excl_a, excl_b;
    float c0, inv_dr, band_lim, d2_fast_max;     // c0 = -lo*inv_dr - 0.5, band_lim = 0.5 - band
    uint32_t hb;                  // shared address of this warp's histogram, biased (see begin())
    const int *ids_a, *ids_b;     // exclude_same: per-atom group ids (device) or NULL

    __device__ __forceinline__ uint32_t *mine() const
    {
        extern __shared__ uint32_t s_hist[];
        return s_hist + (threadIdx.x >> 5) * nbins;
    }
    __device__ __forceinline__ void begin()
    {
        extern __shared__ uint32_t s_hist[];
        for (int w = threadIdx.x; w < nbins * (T / 32); w += T) s_hist[w] = 0u;
        // bin k of an accepted pair arrives as the bit pattern 0x4B400000 + k (see fast_bin)
        hb = (uint32_t)__cvta_generic_to_shared(mine()) - 4u * 0x4B400000u;
    }
    __device__ __forceinline__ bool excluded(int i, int j) const
    {
        if (!EXCL) return false;
        if (ids_a) return ids_a[i] == ids_b[j];          // exclude_same, rdf.py:280-285
        return (i / excl_a) == (j / excl_b);             // exclusion_block, rdf.py:274-278
    }
    // an unordered pair seen once stands for (i,j) and (j,i)
    template <bool SYM> __device__ __forceinline__ uint32_t weight_of(int i, int j) const
    {
        if (!EXCL) return (SYM && i != j) ? 2u : 1u;
        if (!SYM) return excluded(i, j) ? 0u : 1u;
        if (i == j) return excluded(i, i) ? 0u : 1u;
        return (excluded(i, j) ? 0u : 1u) + (excluded(j, i) ? 0u : 1u);
    }
    // FP32 binning, straight-line: the pair is counted here only when the estimate is
    // safely inside a bin and safely inside the cutoff (everything near a decision
    // boundary is left to hit_w()).  floor(u) = round_to_nearest(u - 0.5) by the magic
    // number trick: t = (u - 0.5) + 1.5 * 2^23 (unit spacing for |u| < 2^22, so also for the
    // slightly negative values of d ~ 0) has the bin number in its low mantissa bits and
    // fr = (u - 0.5) - k in [-0.5, 0.5];  |fr| < 0.5 - band  <=>  frac(u) in (band, 1 - band).
    // The host guarantees 0 <= k < nbins for every d2f < d2_fast_max (else band_lim < 0).
    // d2f == 0 sits on an edge, NaN compares false: both fall through.
    // Split in two so that a caller can run the arithmetic of several candidates before
    // any of their (order-preserving, volatile) shared-memory updates.
    __device__ __forceinline__ bool fast_bin(float d2f, uint32_t &addr) const
    {
        float d;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(d2f));
        const float u = fmaf(d, inv_dr, c0);
        const float t = u + 12582912.0f;
        const float fr = u - (t - 12582912.0f);
        addr = hb + 4u * __float_as_uint(t);
        return d2f < d2_fast_max && fabsf(fr) < band_lim;
    }
    template <bool SYM> __device__ __forceinline__ void fast_add(bool ok, uint32_t addr, int i, int j) const
    {
        if (!ok) return;
        uint32_t wgt = SYM ? 2u : 1u;
        if (EXCL) {
            wgt = weight_of<SYM>(i, j);
            if (!wgt) return;
        }
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(wgt));
    }
    template <bool SYM> __device__ __forceinline__ bool try_fast_w(int i, int j, float d2f)
    {
        uint32_t addr;
        const bool ok = fast_bin(d2f, addr);
        fast_add<SYM>(ok, addr, i, j);
        return ok;
    }
    __device__ __forceinline__ bool try_fast(int i, int j, float d2f) { return try_fast_w<false>(i, j, d2f); }
    template <bool SYM> __device__ __forceinline__ void hit_w(int i, int j, double d2, uint64_t)
    {
        const uint32_t wgt = weight_of<SYM>(i, j);
        if (!wgt) return;
        if (!(d2 >= thr[0]) || !(d2 <= thr[nbins])) return;     // np.histogram keeps lo <= d <= hi
        int k = (int)(fmaf(sqrtf((float)d2), inv_dr, c0) + 0.5f);
        k = k < 0 ? 0 : (k > nbins - 1 ? nbins - 1 : k);
        while (k > 0 && d2 < thr[k]) --k;
        while (k < nbins - 1 && d2 >= thr[k + 1]) ++k;
        atomicAdd(mine() + k, wgt);
    }
    __device__ __forceinline__ void hit(int i, int j, double d2, uint64_t f) { hit_w<false>(i, j, d2, f); }
    __device__ __forceinline__ void end()
    {
        extern __shared__ uint32_t s_hist[];
        __syncthreads();
        for (int k = threadIdx.x; k < nbins; k += T) {
            unsigned long long sum = 0;
#pragma unroll
            for (int w = 0; w < T / 32; ++w) sum += s_hist[w * nbins + k];
            if (sum) atomicAdd(&counts[k], sum);
        }
    }
};

// Histogram of rowfine.cuh: ONE histogram per CTA, bin-major with 1 << ls_shift lane columns
// per bin (word = (bin << ls_shift) + (lane & mask)); with 32 columns every warp-wide RED is
// conflict free.  Two scratch rows (nbins, nbins + 1) take the misses.
template <int T>
struct HistSinkFine {
    static constexpr int THREADS = T;
    static constexpr bool FAST = true;
    const double *thr;            // [nbins+1] exact thresholds on d2 (device)
    unsigned long long *counts;   // [nbins] global histogram
    int nbins, ls_shift;
    float c0, inv_dr, band_lim;

    __device__ __forceinline__ int hist_words() const { return (nbins + 2) << ls_shift; }
    __device__ __forceinline__ uint32_t hist_addr() const
    {
        extern __shared__ uint32_t s_hist[];
        return (uint32_t)__cvta_generic_to_shared(s_hist);
    }
    __device__ __forceinline__ void begin()
    {
        extern __shared__ uint32_t s_hist[];
        for (int w = threadIdx.x; w < hist_words(); w += T) s_hist[w] = 0u;
    }
    template <bool SYM> __device__ __forceinline__ void hit_w(int i, int j, double d2, uint64_t)
    {
        extern __shared__ uint32_t s_hist[];
        const uint32_t wgt = (SYM && i != j) ? 2u : 1u;
        if (!(d2 >= thr[0]) || !(d2 <= thr[nbins])) return;     // np.histogram keeps lo <= d <= hi
        int k = (int)(fmaf(sqrtf((float)d2), inv_dr, c0) + 0.5f);
        k = k < 0 ? 0 : (k > nbins - 1 ? nbins - 1 : k);
        while (k > 0 && d2 < thr[k]) --k;
        while (k < nbins - 1 && d2 >= thr[k + 1]) ++k;
        atomicAdd(s_hist + (k << ls_shift) + (threadIdx.x & ((1 << ls_shift) - 1)), wgt);
    }
    __device__ __forceinline__ void end()
    {
        extern __shared__ uint32_t s_hist[];
        __syncthreads();
        const int cols = 1 << ls_shift;
        for (int k = threadIdx.x; k < nbins; k += T) {
            uint32_t sum = 0;                                  // wraps like the columns do (transient un-counts)
            for (int c = 0; c < cols; ++c) sum += s_hist[(k << ls_shift) + ((c + threadIdx.x) & (cols - 1))];
            if (sum) atomicAdd(&counts[k], (unsigned long long)sum);
        }
    }
};

// fallback for very many bins: straight global atomics
struct HistSinkGlobal {
    static constexpr int THREADS = 128;
    static constexpr bool FAST = false;
    const double *thr;
    unsigned long long *counts;
    int nbins, excl_a, excl_b;
    float lo, inv_dr;
    const int *ids_a, *ids_b;
    __device__ __forceinline__ bool try_fast(int, int, float) { return false; }
    template <bool SYM> __device__ __forceinline__ bool try_fast_w(int, int, float) { return false; }
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void end() {}
    __device__ __forceinline__ bool excluded(int i, int j) const
    {
        if (ids_a) return ids_a[i] == ids_b[j];
        return excl_a > 0 && (i / excl_a) == (j / excl_b);
    }
    template <bool SYM> __device__ __forceinline__ void hit_w(int i, int j, double d2, uint64_t)
    {
        unsigned long long wgt = excluded(i, j) ? 0ull : 1ull;
        if (SYM && i != j) wgt += excluded(j, i) ? 0ull : 1ull;
        if (!wgt) return;
        if (!(d2 >= thr[0]) || !(d2 <= thr[nbins])) return;
        int k = (int)((sqrtf((float)d2) - lo) * inv_dr);
        k = k < 0 ? 0 : (k > nbins - 1 ? nbins - 1 : k);
        while (k > 0 && d2 < thr[k]) --k;
        while (k < nbins - 1 && d2 >= thr[k + 1]) ++k;
        atomicAdd(&counts[k], wgt);
    }
    __device__ __forceinline__ void hit(int i, int j, double d2, uint64_t f) { hit_w<false>(i, j, d2, f); }
};

// InterRDF_s (rdf.py:585-597): one histogram per (site of A, site of B) pair,
// counts[(i * nb + j) * nbins + k]; an unordered pair of a self search stands for (i, j)
// and (j, i).
struct HistSinkSites {
    static constexpr int THREADS = 128;
    static constexpr bool FAST = false;
    const double *thr;
    unsigned long long *counts;
    int nbins;
    uint64_t nb;
    float lo, inv_dr;
    __device__ __forceinline__ bool try_fast(int, int, float) { return false; }
    template <bool SYM> __device__ __forceinline__ bool try_fast_w(int, int, float) { return false; }
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void end() {}
    template <bool SYM> __device__ __forceinline__ void hit_w(int i, int j, double d2, uint64_t)
    {
        if (!(d2 >= thr[0]) || !(d2 <= thr[nbins])) return;
        int k = (int)((sqrtf((float)d2) - lo) * inv_dr);
        k = k < 0 ? 0 : (k > nbins - 1 ? nbins - 1 : k);
        while (k > 0 && d2 < thr[k]) --k;
        while (k < nbins - 1 && d2 >= thr[k + 1]) ++k;
        atomicAdd(&counts[((uint64_t)i * nb + (uint64_t)j) * (uint64_t)nbins + k], 1ull);
        if (SYM && i != j) atomicAdd(&counts[((uint64_t)j * nb + (uint64_t)i) * (uint64_t)nbins + k], 1ull);
    }
    __device__ __forceinline__ void hit(int i, int j, double d2, uint64_t f) { hit_w<false>(i, j, d2, f); }
};

// contacts: pair count per frame and optional per-atom degree
struct CountSink {
    static constexpr int THREADS = 128;
    static constexpr bool FAST = false;
    unsigned long long *counts;   // [nframes]
    int *degree;                  // [nframes*n] or NULL
    uint64_t n;
    unsigned local;
    __device__ __forceinline__ bool try_fast(int, int, float) { return false; }
    template <bool SYM> __device__ __forceinline__ bool try_fast_w(int, int, float) { return false; }
    template <bool SYM> __device__ __forceinline__ void hit_w(int i, int j, double d2, uint64_t f) { hit(i, j, d2, f); }
    __device__ __forceinline__ void begin() { local = 0; }
    __device__ __forceinline__ void hit(int i, int j, double, uint64_t f)
    {
        ++local;
        if (degree) {
            atomicAdd(&degree[f * n + i], 1);
            atomicAdd(&degree[f * n + j], 1);
        }
    }
    __device__ __forceinline__ void end()
    {
        unsigned v = local;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counts[blockIdx.y], (unsigned long long)v);
    }
};

// How one chunk of frames is traversed.
struct Traversal {
    int row;            // 1: row_tile_kernel (rowtile.cuh), 0: cell_tile_kernel (engine.cuh)
    int quad;           // with row: row_quad_kernel (rowquad.cuh) when the sink and geometry allow it
    RowParams rp;
    uint64_t grid_x;
    int chunks;
};

template <int MIC, bool SELF, class Sink>
static int launch_hist(const CellList &ql, const CellList &cl, const FrameGeom *fg_dev, int per_frame,
                       uint64_t nframes, const Traversal &tv, const SearchParams &sp,
                       const Sink &sink, size_t smem, cudaStream_t st)
{
    dim3 grid((unsigned)tv.grid_x, (unsigned)nframes);
    if constexpr (Sink::THREADS == 224) {
        // the 224-thread sink exists for the packed kernel only (a whole row of home cells per CTA)
        if constexpr (MIC == MIC_ORTHO_LAZY || MIC == MIC_TRICLINIC) {
            if (tv.row && tv.quad) {
                auto kern = row_quad_kernel<SELF, Sink, WQ_CAP_224, 4, MIC>;
                const size_t total = ((smem + 15) & ~(size_t)15) + 3 * (size_t)WQ_CAP_224 * sizeof(float);
                MDA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
                MDA_LAUNCH_AS("rdf_pair_histogram", kern, grid, Sink::THREADS, total, st, ql.atoms(), ql.start(),
                              cl.atoms(), cl.start(), fg_dev, per_frame, tv.rp, sp, sink);
                MDA_CHECK_LAUNCH();
                return MDAGPU_OK;
            }
        }
        return set_error(MDAGPU_ERR_INVALID, "internal: 224-thread histogram sink without the packed kernel");
    } else {
    if constexpr ((MIC == MIC_ORTHO_LAZY || MIC == MIC_TRICLINIC) && std::is_same<Sink, HistSink<128, false>>::value) {
        if (tv.row && tv.quad) {
            auto kern = row_quad_kernel<SELF, Sink, WQ_CAP_128, 7, MIC>;
            const size_t total = ((smem + 15) & ~(size_t)15) + 3 * (size_t)WQ_CAP_128 * sizeof(float);
            if (total > 24 * 1024)
                MDA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
            MDA_LAUNCH_AS("rdf_pair_histogram", kern, grid, Sink::THREADS, total, st, ql.atoms(), ql.start(),
                          cl.atoms(), cl.start(), fg_dev, per_frame, tv.rp, sp, sink);
            MDA_CHECK_LAUNCH();
            return MDAGPU_OK;
        }
    }
    if (tv.row) {
        auto kern = row_tile_kernel<MIC, SELF, Sink>;
        if (smem > 24 * 1024)
            MDA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MDA_LAUNCH_AS("rdf_pair_histogram", kern, grid, Sink::THREADS, smem, st, ql.atoms(), ql.start(),
                      cl.atoms(), cl.start(), fg_dev, per_frame, tv.rp, sp, sink);
    } else {
        auto kern = cell_tile_kernel<MIC, SELF, Sink>;
        if (smem > 24 * 1024)
            MDA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MDA_LAUNCH_AS("rdf_pair_histogram", kern, grid, Sink::THREADS, smem, st, ql.atoms(), ql.start(),
                      cl.atoms(), cl.start(), fg_dev, per_frame, tv.chunks, sp, sink);
    }
    MDA_CHECK_LAUNCH();
    return MDAGPU_OK;
    }
}

template <class Sink>
static int dispatch_hist(int mic, bool self, const CellList &ql, const CellList &cl,
                         const FrameGeom *fg_dev, int per_frame, uint64_t nframes, const Traversal &tv,
                         const SearchParams &sp, const Sink &sink, size_t smem, cudaStream_t st)
{
#define HIST_CASE(M)                                                                               \
    case M:                                                                                        \
        return self ? launch_hist<M, true>(ql, cl, fg_dev, per_frame, nframes, tv, sp, sink, smem, st) \
                    : launch_hist<M, false>(ql, cl, fg_dev, per_frame, nframes, tv, sp, sink, smem, st)
    switch (mic) {
        HIST_CASE(MIC_NONE);
        HIST_CASE(MIC_ORTHO_ROUND);
        HIST_CASE(MIC_ORTHO_LAZY);
        HIST_CASE(MIC_TRICLINIC);
    }
#undef HIST_CASE
    return set_error(MDAGPU_ERR_INVALID, "unknown minimum-image flavour %d", mic);
}

// bounding box of [nframes, n, 3] coordinates (for no-box grids)
__global__ void __launch_bounds__(256)
bbox_kernel(const float *__restrict__ xyz, uint64_t total, float *__restrict__ lohi)
{
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (uint64_t)gridDim.x * blockDim.x)
        for (int k = 0; k < 3; ++k) {
            float v = xyz[3 * t + k];
            lo[k] = fminf(lo[k], v);
            hi[k] = fmaxf(hi[k], v);
        }
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            // ordered-int trick for float atomics (values are finite)
            int l = __float_as_int(lo[k]), h = __float_as_int(hi[k]);
            if (l >= 0) atomicMin((int *)&lohi[k], l); else atomicMax((unsigned *)&lohi[k], (unsigned)l);
            if (h >= 0) atomicMax((int *)&lohi[3 + k], h); else atomicMin((unsigned *)&lohi[3 + k], (unsigned)h);
        }
    }
}

// The three "this piece cannot go through the planner / walker pair" flags of a piece in one word
// triple that the host reads one piece later (no host synchronisation inside a piece).
__global__ void rdf_flags_kernel(const unsigned *over, const unsigned *bad0, const unsigned *bad1, unsigned *out3)
{
    out3[0] = over ? *over : 0u;
    out3[1] = bad0 ? *bad0 : 0u;
    out3[2] = bad1 ? *bad1 : 0u;
}

// pinned landing slots for those flags (per thread, reused by every call)
struct FlagSlots {
    static constexpr int N = 64;
    unsigned *host = nullptr;
    cudaEvent_t ev[N] = { nullptr };
    cudaError_t init()
    {
        if (host) return cudaSuccess;
        cudaError_t e = cudaHostAlloc((void **)&host, N * 4 * sizeof(unsigned), cudaHostAllocDefault);
        for (int k = 0; k < N && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
        return e;
    }
};
static FlagSlots &flag_slots()
{
    static thread_local FlagSlots fs[8];          // recursion depth of rdf_accumulate_impl is 2; per device below
    int dev = 0;
    cudaGetDevice(&dev);
    return fs[dev & 7];
}

// Upload ring for host-resident frames: piece k is copied on a private stream
// into buffer k % NB (after the kernels that read the buffer NB pieces ago have finished)
// and the compute stream waits for the copy before it touches the buffer.
struct HostPipe {
    const float *A = nullptr, *B = nullptr;
    uint64_t na = 0, nb = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    static constexpr int NB = 3;                      // buffers: uploads run up to two pieces ahead of the kernels
    cudaEvent_t ready[NB] = { nullptr }, freed[NB] = { nullptr };
    DevBuf abuf[NB], bbuf[NB];
    ~HostPipe()
    {
        for (int k = 0; k < NB; ++k) {
            if (ready[k]) cudaEventDestroy(ready[k]);
            if (freed[k]) cudaEventDestroy(freed[k]);
        }
        if (copy) {
            cudaStreamSynchronize(copy);
            cudaStreamDestroy(copy);
        }
    }
    cudaError_t init(const float *a, const float *b, uint64_t na_, uint64_t nb_, uint64_t piece, cudaStream_t st)
    {
        A = a; B = b; na = na_; nb = nb_; compute = st;
        cudaError_t e = cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking);
        for (int k = 0; k < NB && e == cudaSuccess; ++k) {
            e = cudaEventCreateWithFlags(&ready[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&freed[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = abuf[k].alloc(piece * na * 3 * sizeof(float), st);
            if (e == cudaSuccess && B) e = bbuf[k].alloc(piece * nb * 3 * sizeof(float), st);
            // the buffers become usable in stream order of `st`: the first copies wait for that
            if (e == cudaSuccess) e = cudaEventRecord(freed[k], st);
        }
        return e;
    }
    cudaError_t upload(uint64_t piece_no, uint64_t f0, uint64_t fc)
    {
        const int k = (int)(piece_no % NB);
        cudaError_t e = done[k] ? cudaSuccess : cudaStreamWaitEvent(copy, freed[k], 0);
        done[k] = false;
        if (e == cudaSuccess)
            e = host_to_device(abuf[k].p, A + f0 * na * 3, fc * na * 3 * sizeof(float), copy);
        if (e == cudaSuccess && B)
            e = host_to_device(bbuf[k].p, B + f0 * nb * 3, fc * nb * 3 * sizeof(float), copy);
        if (e == cudaSuccess) e = cudaEventRecord(ready[k], copy);
        return e;
    }
    cudaError_t acquire(uint64_t piece_no, const float **a, const float **b)
    {
        const int k = (int)(piece_no % NB);
        *a = abuf[k].as<float>();
        *b = B ? bbuf[k].as<float>() : nullptr;
        return cudaStreamWaitEvent(compute, ready[k], 0);
    }
    cudaError_t release(uint64_t piece_no) { return cudaEventRecord(freed[(int)(piece_no % NB)], compute); }
    // the host has SEEN the piece finish (it synchronised on something behind its kernels): no event --
    // one recorded now would sit behind the kernels of the pieces queued since
    bool done[NB] = { false };
    void release_completed(uint64_t piece_no) { done[(int)(piece_no % NB)] = true; }
};

}  // namespace mdagpu

using namespace mdagpu;

extern "C" int mdagpu_grid_stencil_complete(double cutoff, const float *box6, const float *h9, int W, int *complete)
{
    MDA_REQUIRE(box6 && h9 && complete, "mdagpu_grid_stencil_complete: NULL argument");
    MDA_REQUIRE(W >= 1, "mdagpu_grid_stencil_complete: W must be >= 1");
    GridGeom g;
    float mc = 0.0f;
    grid_geometry_reference(cutoff, box6, h9, 1, &g, &mc);
    bool ok = true;
    for (int k = 0; k < 3; ++k) ok = ok && g.periodic[k];          // (a non-periodic axis is one cell: complete, but not this test's business)
    *complete = (ok && (g.triclinic ? tric_stencil_complete(g, h9, cutoff, W)
                                     : (W * g.cs[0] >= cutoff && W * g.cs[4] >= cutoff && W * g.cs[8] >= cutoff))) ? 1 : 0;
    return MDAGPU_OK;
}


static int rdf_accumulate_impl(const float *A, uint64_t na, const float *B, uint64_t nb,
                               uint64_t nframes, const float *box6, const float *h9,
                               int boxes_per_frame, int mic, int wrap, double cutoff,
                               double d2_max, int use_grid, const double *thresholds,
                               int nbins, double r_lo, double r_hi, int excl_a, int excl_b,
                               const int32_t *ids_a_host, const int32_t *ids_b_host,
                               unsigned long long *counts, int per_site, int space, mdagpu_stream_t stream,
                               int force_basic = 0)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(A && thresholds && counts, "mdagpu_rdf_accumulate: NULL argument");
    MDA_REQUIRE(nbins > 0, "mdagpu_rdf_accumulate: nbins must be positive");
    MDA_REQUIRE(mic >= 0 && mic <= 3, "mdagpu_rdf_accumulate: bad mic %d", mic);
    MDA_REQUIRE(mic == MIC_NONE || (box6 && h9), "mdagpu_rdf_accumulate: box is NULL");
    MDA_REQUIRE((excl_a > 0) == (excl_b > 0), "mdagpu_rdf_accumulate: exclusion block needs two sizes");
    const bool same = (B == nullptr || B == A);
    if (same) nb = na;
    if (nframes == 0 || na == 0 || nb == 0) return MDAGPU_OK;
    // a self RDF with exclusion_block=(1, 1) only drops the pairs (i, i)
    int report_self = same ? 1 : 0;
    if (same && excl_a == 1 && excl_b == 1) { excl_a = excl_b = 0; report_self = 0; }

    // ---- per-frame geometry on the host ------------------------------------------------
    // FastNS semantics (use_grid): the reference's own grid decides which pairs a
    // triclinic search can see, so it is reproduced; for an orthorhombic cell every
    // complete grid yields the same pair set and a finer one (cells >= cutoff/2, 5-cell
    // window) is used when the box has room for it: fewer candidates per query.
    const uint64_t nbox = boxes_per_frame ? nframes : 1;
    std::vector<FrameGeom> fgs(nbox);
    int W = 0;                                  // 0: cell_tile_kernel on the reference grid
    // triclinic cells: a finer grid (W = 2) only when the reference's own stencil is complete for
    // every frame's cell (the pair set is then "every pair within the cutoff" on any complete grid)
    bool tric_fine = mic == MIC_TRICLINIC && use_grid && box6 && h9 && !getenv("MDAGPU_RDF_NO_TRIC_FINE");
    for (uint64_t f = 0; f < nbox && tric_fine; ++f) {
        GridGeom gref;
        float mc = 0.0f;
        grid_geometry_reference(cutoff, box6 + 6 * f, h9 + 9 * f, 1, &gref, &mc);
        if (!gref.triclinic || !tric_stencil_complete(gref, h9 + 9 * f, cutoff, 1)) tric_fine = false;
    }
    const int tiled = (!use_grid || force_basic == 1) ? 0 : (mic == MIC_ORTHO_LAZY ? 2 : (mic == MIC_TRICLINIC ? (tric_fine ? 2 : 1) : 0));
    // the packed kernels (rowfine.cuh, rowquad.cuh) fold the cutoff into the top histogram edge and
    // have no exclusion blocks
    const bool by_ids = ids_a_host != nullptr;
    const bool packed_any = tiled == 2 && !per_site && !by_ids && excl_a == 0 && r_hi > 0.0 && r_lo <= 0.0
                            && fabs(d2_max - r_hi * r_hi) <= 1e-12 * d2_max && !getenv("MDAGPU_RDF_NO_QUAD");
    const bool packed_ok = packed_any && mic == MIC_ORTHO_LAZY;
    // triclinic cells: the packed kernel on the library's own grid (W = 2) only -- on the reference's
    // grid a lane's cell window is part of the result
    const bool packed_tric = packed_any && mic == MIC_TRICLINIC && !getenv("MDAGPU_RDF_NO_TRIC_QUAD");
    // ---- rowfine.cuh: fine cells along x (cutoff / WX), cutoff / 2 along y and z ----
    int fineWX = 0;
    float fine_reach = 0.0f;
    if (packed_ok && force_basic != 2 && !getenv("MDAGPU_RDF_NO_FINE")) {
        const int choices[2] = { 8, 4 };
        for (int ci = 0; ci < 2 && !fineWX; ++ci) {
            const int wx = choices[ci];
            bool ok = true;
            int64_t base = 0;
            double reach = 0.0;
            for (uint64_t f = 0; f < nbox && ok; ++f) {
                FrameGeom &fg = fgs[f];
                const float *b6 = box6 + 6 * f;
                fg.box = make_box_params(BOX_ORTHO, b6);
                float mc = 0.0f;
                grid_geometry_reference(cutoff, b6, h9 + 9 * f, 1, &fg.grid, &mc);
                MDA_REQUIRE(!(cutoff > (double)mc), "Cutoff %g too large for box (max %g)", cutoff, (double)mc);
                const double c3[3] = { cutoff / wx, cutoff / 2.0, cutoff / 2.0 };
                grid_geometry_ortho_axes(c3, b6, &fg.grid);
                const GridGeom &g = fg.grid;
                const double lmax = fmax((double)b6[0], fmax((double)b6[1], (double)b6[2]));
                // margin of the per-lane windows: FP32 evaluation of the gaps (box-sized values)
                const double rch = cutoff * (1.0 + 2e-5) + 1e-4 + 4e-6 * lmax;
                reach = fmax(reach, rch);
                const int w3[3] = { wx, 2, 2 };
                for (int k = 0; k < 3; ++k) {
                    if (!g.periodic[k] || g.ncells[k] < 2 * w3[k] + 2) ok = false;
                    if (!(rch <= w3[k] * g.cs[4 * k])) ok = false;                 // the stencil covers the reach
                    if (!((double)b6[k] > 2.0 * rch + 2.0 * g.cs[4 * k])) ok = false;   // one image per window
                }
                // the FP32 path needs a usable band (see derive() below) and a histogram that fits
                {
                    const double inv = (double)nbins / (r_hi - r_lo);
                    const double u_err = (4e-7 * fmax(fabs(r_hi), fabs(r_lo)) + rowquad_abs_err(fg.box, cutoff)) * inv
                                         + 2e-7 * (fabs(-r_lo * inv - 0.5) + (double)nbins);
                    if (!(fmax(5e-4, 2.0 * u_err) < 0.2)) ok = false;
                    if ((size_t)(nbins + 2) * sizeof(uint32_t) > 20 * 1024) ok = false;
                }
                // thin cells make the staged rows mostly bookkeeping
                if ((double)nb / (double)g.ncell_total < (wx == 8 ? 1.5 : 0.75)) ok = false;
                fg.cell_base = base;
                base += g.ncell_total;
            }
            if (ok) { fineWX = wx; fine_reach = (float)reach; W = 2; }
        }
    }
    // ---- rowfine.cuh for triclinic cells: the same traversal in the cell coordinates of cell_of() ----
    // Only where the reference's own stencil is complete (tric_fine: the pair set is then "every pair
    // within the cutoff") and the lattice images are whole grids in cell coordinates (t_xy c_y = 0).
    bool fine_tric = false;
    if (!fineWX && packed_tric && tric_fine && force_basic != 2 && !getenv("MDAGPU_RDF_NO_FINE")
        && !getenv("MDAGPU_RDF_NO_TRIC_FINE_X")) {
        const int choices[2] = { 8, 4 };
        for (int ci = 0; ci < 2 && !fineWX; ++ci) {
            const int wx = choices[ci];
            bool ok = true;
            int64_t base = 0;
            double reach = 0.0;
            for (uint64_t f = 0; f < nbox && ok; ++f) {
                FrameGeom &fg = fgs[f];
                const float *hh = h9 + 9 * f;
                fg.box = make_box_params(BOX_TRICLINIC, hh);
                float mc = 0.0f;
                grid_geometry_reference(cutoff, box6 + 6 * f, hh, 1, &fg.grid, &mc);
                MDA_REQUIRE(!(cutoff > (double)mc), "Cutoff %g too large for box (max %g)", cutoff, (double)mc);
                GridGeom &g = fg.grid;
                g.triclinic = 1;
                g.pbc = 1;
                for (int k = 0; k < 9; ++k) g.cs[k] = 0.0;
                g.origin[0] = g.origin[1] = g.origin[2] = 0.0;
                g.cs[3] = (double)(hh[3] / hh[4]);
                g.cs[6] = (double)(hh[6] / hh[8]);
                g.cs[7] = (double)(hh[7] / hh[8]);
                if (fabs(g.cs[3] * (double)hh[7]) > 1e-7 * (double)hh[0]) ok = false;
                double lmax = 0.0;
                for (int k = 0; k < 3; ++k)
                    lmax = fmax(lmax, fabs((double)hh[k]) + fabs((double)hh[3 + k]) + fabs((double)hh[6 + k]));
                // margin of the per-lane windows: cell coordinates and interval ends evaluated in FP32
                // from box-sized values
                const double rch = cutoff * (1.0 + 2e-5) + 1e-4 + 1.2e-5 * lmax;
                reach = fmax(reach, rch);
                const double fxx = sqrt(1.0 + g.cs[3] * g.cs[3] + g.cs[6] * g.cs[6]), fyy = sqrt(1.0 + g.cs[7] * g.cs[7]);
                const double need[3] = { rch * fxx / wx, rch * fyy / 2.0, rch / 2.0 };
                const double edge[3] = { (double)hh[0], (double)hh[4], (double)hh[8] };
                const int w3[3] = { wx, 2, 2 };
                int64_t total = 1;
                for (int k = 0; k < 3; ++k) {
                    double nc = floor(edge[k] / (need[k] * (1.0 + 1e-4) + 1e-3));
                    if (!(nc >= 1.0)) nc = 1.0;
                    if (nc > 256.0) nc = 256.0;
                    g.ncells[k] = (int)nc;
                    g.periodic[k] = g.ncells[k] >= 3 ? 1 : 0;
                    g.cs[4 * k] = edge[k] / (double)g.ncells[k];
                    total *= g.ncells[k];
                    if (!g.periodic[k] || g.ncells[k] < 2 * w3[k] + 2) ok = false;
                }
                g.ncell_total = total;
                // one image per window: perpendicular widths as in rowtile_supported(), with the
                // stencil half-width of each axis
                if (ok) {
                    const double h0[3] = { hh[0], hh[1], hh[2] }, h1[3] = { hh[3], hh[4], hh[5] }, h2[3] = { hh[6], hh[7], hh[8] };
                    auto cross = [](const double *u, const double *v, double *o) {
                        o[0] = u[1] * v[2] - u[2] * v[1]; o[1] = u[2] * v[0] - u[0] * v[2]; o[2] = u[0] * v[1] - u[1] * v[0];
                    };
                    auto norm = [](const double *u) { return sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]); };
                    double c12[3], c20[3], c01[3];
                    cross(h1, h2, c12); cross(h2, h0, c20); cross(h0, h1, c01);
                    const double vol = fabs(h0[0] * c12[0] + h0[1] * c12[1] + h0[2] * c12[2]);
                    const double width[3] = { vol / norm(c12), vol / norm(c20), vol / norm(c01) };
                    for (int k = 0; k < 3; ++k) {
                        if (!(width[k] > 2.0 * rch * (1.0 + 1e-5))) ok = false;
                        if (!((1.0 - (double)(w3[k] + 1) / g.ncells[k]) * width[k] > rch * (1.0 + 1e-5))) ok = false;
                    }
                    // a window reaches at most wx cells into the halo that is staged on either side
                    if (!((double)wx * g.cs[0] >= rch * fxx)) ok = false;
                }
                {
                    const double inv = (double)nbins / (r_hi - r_lo);
                    const double u_err = (4e-7 * fmax(fabs(r_hi), fabs(r_lo)) + rowquad_abs_err(fg.box, cutoff)) * inv
                                         + 2e-7 * (fabs(-r_lo * inv - 0.5) + (double)nbins);
                    if (!(fmax(5e-4, 2.0 * u_err) < 0.2)) ok = false;
                    if ((size_t)(nbins + 2) * sizeof(uint32_t) > 20 * 1024) ok = false;
                }
                if ((double)nb / (double)g.ncell_total < (wx == 8 ? 1.5 : 0.75)) ok = false;
                fg.cell_base = base;
                base += g.ncell_total;
            }
            if (ok) { fineWX = wx; fine_reach = (float)reach; W = 2; fine_tric = true; }
        }
    }
    for (int tryW = tiled; tryW >= 0 && W == 0; --tryW) {
        int64_t base = 0;
        bool ok = tryW > 0;
        for (uint64_t f = 0; f < nbox; ++f) {
            FrameGeom &fg = fgs[f];
            const float *b6 = box6 ? box6 + 6 * f : nullptr;
            const float *hh = h9 ? h9 + 9 * f : nullptr;
            const bool tri = hh && (hh[3] != 0.0f || hh[6] != 0.0f || hh[7] != 0.0f);
            if (mic == MIC_NONE && !b6) fg.box = make_box_params(BOX_NONE, nullptr);
            else fg.box = tri ? make_box_params(BOX_TRICLINIC, hh) : make_box_params(BOX_ORTHO, b6);
            if (use_grid) {
                float mc = 0.0f;
                grid_geometry_reference(cutoff, b6, hh, 1, &fg.grid, &mc);
                // FastNS refuses such a cutoff (nsgrid.pyx:288-291)
                MDA_REQUIRE(!(cutoff > (double)mc), "Cutoff %g too large for box (max %g)", cutoff, (double)mc);
                if (tryW > 0 && mic == MIC_ORTHO_LAZY)
                    grid_geometry_ortho_free(cutoff / tryW, b6, nullptr, 1, &fg.grid);
                if (tryW > 1 && mic == MIC_TRICLINIC) {
                    grid_geometry_triclinic_free(cutoff, tryW, hh, &fg.grid);
                    if (!tric_stencil_complete(fg.grid, hh, cutoff, tryW)) ok = false;
                }
            } else {
                grid_geometry_single(&fg.grid);
            }
            if (tryW > 0) {
                ok = ok && rowtile_supported(mic, fg.grid, fg.box, cutoff, tryW);
                // very thin cells make the staged rows mostly bookkeeping
                if (tryW > 1 && (double)nb / (double)fg.grid.ncell_total < 2.0) ok = false;
            }
            fg.cell_base = base;
            base += fg.grid.ncell_total;
        }
        if (ok) W = tryW;
    }
    if (getenv("MDAGPU_DEBUG"))
        fprintf(stderr, "mdagpu: rdf traversal W = %d, fine x = %d, triclinic fine grid allowed = %d, grid %d x %d x %d\n", W, fineWX,
                (int)tric_fine, fgs[0].grid.ncells[0], fgs[0].grid.ncells[1], fgs[0].grid.ncells[2]);
    const int64_t cells_per_frame = boxes_per_frame ? 0 : fgs[0].grid.ncell_total;
    // rowquad.cuh keeps one 32-bit histogram per warp of a 128-thread CTA
    const bool quad = ((packed_ok && W > 0) || (packed_tric && W == 2)) && (size_t)nbins * 4 * sizeof(uint32_t) <= 160 * 1024;
    // error of the FP32 distance estimate for the row traversals (per-cell-pair image prediction)
    // and for the per-pair traversals the code falls back to (27-image FP32 scan for triclinic cells)
    double err_row = 0.0, err_cell = 0.0;
    for (uint64_t f = 0; f < nbox; ++f) {
        if (W) err_row = fmax(err_row, quad ? rowquad_abs_err(fgs[f].box, cutoff) : rowtile_abs_err(mic, fgs[f].box, cutoff));
        err_cell = fmax(err_cell, mic == MIC_TRICLINIC ? tric_abs_err(fgs[f].box.h) : 0.0);
    }

    StagedIn thr_in, ida_in, idb_in;
    StagedOut cnt;
    if (by_ids) {      // exclude_same: group ids of the atoms of A and of B (host arrays)
        MDA_CUDA(ida_in.init(ids_a_host, na * sizeof(int32_t), MDAGPU_HOST, st));
        MDA_CUDA(idb_in.init(ids_b_host ? ids_b_host : ids_a_host, nb * sizeof(int32_t), MDAGPU_HOST, st));
    }
    const int *ida = by_ids ? ida_in.as<int>() : nullptr, *idb = by_ids ? idb_in.as<int>() : nullptr;
    MDA_CUDA(thr_in.init(thresholds, (size_t)(nbins + 1) * sizeof(double), MDAGPU_HOST, st));   // always a host array
    const size_t n_hist = per_site ? (size_t)na * (size_t)nb : 1;          // histograms in `counts`
    MDA_CUDA(cnt.init(counts, n_hist * (size_t)nbins * sizeof(unsigned long long), space, st, true));

    SearchParams sp;
    sp.d2_max = d2_max;
    sp.d2_min = -1.0;
    sp.d2_filter = 0.0f;                         // per traversal, see derive()
    sp.use_filter = mic != MIC_ORTHO_ROUND;
    sp.stencil = use_grid ? 1 : 0;

    // FP32 bin coordinate u = (d - lo) * inv_dr - 0.5: its error is the error of d (4e-7
    // relative from the FP32 evaluation and the approximate sqrt, d_abs_err absolute)
    // times inv_dr; pairs closer than twice that to a bin edge are decided in float64.
    // A band >= 0.5 disables the FP32 path.
    const float inv_dr = (float)((double)nbins / (r_hi - r_lo));
    const float c0 = (float)(-r_lo * (double)inv_dr - 0.5);
    struct Derived { float d2_filter, band_lim, d2_fast_max; };
    auto derive = [&](double d_abs_err) {
        Derived dv;
        dv.d2_filter = filter_threshold(d2_max, d_abs_err);
        const double u_err = (4e-7 * fmax(fabs(r_hi), fabs(r_lo)) + d_abs_err) * (double)nbins / (r_hi - r_lo)
                             + 2e-7 * (fabs((double)c0) + (double)nbins);
        const float band = (float)fmax(5e-4, 2.0 * u_err);
        // the straight-line FP32 path does not range-check the bin: it is enabled only when
        // every pair it accepts (d2f < d2_fast_max) provably lands in [0, nbins)
        double fast_d2 = d2_max;
        if (r_hi > 0.0 && r_hi * r_hi < fast_d2) fast_d2 = r_hi * r_hi;
        const bool fast_ok = r_lo <= 0.0 && r_hi > 0.0 && band < 0.25f;
        dv.band_lim = fast_ok ? 0.5f - band : -1.0f;
        dv.d2_fast_max = certain_threshold(fast_d2, d_abs_err);
        return dv;
    };

    // frames are processed in chunks that keep keys 32-bit and the grid within limits
    uint64_t chunk = nframes;
    const uint64_t per_frame_atoms = na > nb ? na : nb;
    const uint64_t max_atoms = 1ull << 27;                 // 128 Mi atoms (2 GiB of float4) per chunk
    if (chunk * per_frame_atoms > max_atoms) chunk = max_atoms / per_frame_atoms ? max_atoms / per_frame_atoms : 1;
    if (chunk > 65535) chunk = 65535;
    if (fineWX) {
        // the planner's record lists: ~11 kB per (y, z) row and x segment; at most 1 GiB per chunk
        const uint64_t rows = (uint64_t)fgs[0].grid.ncells[1] * (uint64_t)fgs[0].grid.ncells[2];
        uint64_t by_lists = (2ull << 30) / (rows * 4ull * 4096ull * 2ull);
        if (by_lists < 1) by_lists = 1;
        if (chunk > by_lists) chunk = by_lists;
    }
    if (!boxes_per_frame && cells_per_frame > 0) {
        uint64_t by_cells = (uint64_t)((1ull << 30) / (uint64_t)cells_per_frame);
        if (by_cells < 1) by_cells = 1;
        if (chunk > by_cells) chunk = by_cells;
    }

    // Host input is streamed: frames go up in pieces on a second stream into a ring of
    // buffers while the pieces before are being binned (the reference-facing call hands over
    // HOST arrays; 400 MB per 1000 frames of C2 would otherwise sit in front of the kernels).
    // A piece is a quarter of the call (measured: 1/3 ... 1/16 make no difference, profiles/r02_notes.md):
    // staging a piece (pageable frames: ~36 GB/s through the pinned ring, on this thread) takes half the
    // time its kernels do and uploads run up to two pieces ahead; the first pieces are smaller --
    // their upload is the one nothing hides.
    HostPipe pipe;
    if (space == MDAGPU_HOST) {
        static const int div_env = getenv("MDAGPU_RDF_PIECE_DIV") ? atoi(getenv("MDAGPU_RDF_PIECE_DIV")) : 0;
        const uint64_t div = div_env > 0 ? (uint64_t)div_env : 4;
        uint64_t piece = (nframes + div - 1) / div;
        if (piece < 16) piece = 16;
        if (chunk > piece) chunk = piece;
        MDA_CUDA(pipe.init(A, same ? nullptr : B, na, nb, chunk, st));
    }

    // piece schedule: device input in equal chunks; host input starts at an eighth of a piece and doubles
    std::vector<uint64_t> piece_start;
    {
        uint64_t size = chunk;
        if (space == MDAGPU_HOST) {
            static const int ramp_env = getenv("MDAGPU_RDF_RAMP_DIV") ? atoi(getenv("MDAGPU_RDF_RAMP_DIV")) : 0;
            size = chunk / (ramp_env > 0 ? (uint64_t)ramp_env : 8);
            if (size < 16) size = 16;
            if (size > chunk) size = chunk;
        }
        for (uint64_t f0 = 0; f0 < nframes;) {
            piece_start.push_back(f0);
            f0 += size;
            size = size * 2 > chunk ? chunk : size * 2;
        }
        piece_start.push_back(nframes);
    }

    DevBuf fg_dev;
    // Upload ring bookkeeping: piece p may go up once piece p - NB has been released.
    const uint64_t n_pieces = piece_start.size() - 1;
    uint64_t next_upload = 0, n_released = 0;
    auto pump = [&]() -> cudaError_t {
        while (space == MDAGPU_HOST && next_upload < n_pieces && next_upload < n_released + HostPipe::NB) {
            cudaError_t e = pipe.upload(next_upload, piece_start[next_upload], piece_start[next_upload + 1] - piece_start[next_upload]);
            if (e != cudaSuccess) return e;
            ++next_upload;
        }
        return cudaSuccess;
    };
    auto release_piece = [&](uint64_t p, bool seen_complete = false) -> cudaError_t {
        if (space != MDAGPU_HOST) return cudaSuccess;
        cudaError_t e = cudaSuccess;
        if (seen_complete) pipe.release_completed(p);
        else e = pipe.release(p);
        ++n_released;
        return e == cudaSuccess ? pump() : e;
    };
    // Pieces of the planner / walker path whose flags the host has not looked at yet.  The walker
    // skips a flagged piece on the device; the host reads the flags ONE PIECE LATER -- by then the
    // next piece is queued behind it, so the device never waits for the host -- and sends a
    // flagged piece through the other kernels.  Only then is the piece's upload buffer released.
    struct Pending { uint64_t piece, f0, fc; const float *a, *b; int slot; };
    std::vector<Pending> pending;
    int slots_used = 0;
    auto settle = [&](size_t keep) -> int {
        while (pending.size() > keep) {
            const Pending pd = pending.front();
            pending.erase(pending.begin());
            FlagSlots &fs = flag_slots();
            MDA_CUDA(cudaEventSynchronize(fs.ev[pd.slot]));
            const unsigned *fl = fs.host + 4 * pd.slot;
            if (fl[0] | fl[1] | fl[2]) {
                const int how = (fl[1] | fl[2]) ? 1 : 2;           // folded atoms: reference grid; planner declined: no planner
                if (getenv("MDAGPU_DEBUG"))
                    fprintf(stderr, "mdagpu: rdf piece of %llu frames redone without the planner (flags %u %u %u)\n",
                            (unsigned long long)pd.fc, fl[0], fl[1], fl[2]);
                int rc2 = rdf_accumulate_impl(pd.a, na, same ? nullptr : pd.b, nb, pd.fc, boxes_per_frame ? box6 + 6 * pd.f0 : box6,
                                              boxes_per_frame ? h9 + 9 * pd.f0 : h9, boxes_per_frame, mic, wrap, cutoff, d2_max,
                                              use_grid, thresholds, nbins, r_lo, r_hi, (same && !report_self) ? 1 : excl_a,
                                              (same && !report_self) ? 1 : excl_b, ids_a_host, ids_b_host,
                                              cnt.as<unsigned long long>(), per_site, MDAGPU_DEVICE, stream, how);
                if (rc2) return rc2;
            }
            // (a redone piece queued new kernels that read the buffer: then the event is needed)
            MDA_CUDA(release_piece(pd.piece, !(fl[0] | fl[1] | fl[2])));
        }
        return MDAGPU_OK;
    };
    if (space == MDAGPU_HOST) {                           // only piece 0 goes up before the first kernels are queued
        MDA_CUDA(pipe.upload(0, piece_start[0], piece_start[1] - piece_start[0]));
        next_upload = 1;
    }
    for (uint64_t piece_no = 0; piece_no + 1 < piece_start.size(); ++piece_no) {
        const uint64_t f0 = piece_start[piece_no], fc = piece_start[piece_no + 1] - f0;
        const float *a_dev = A + f0 * na * 3, *b_dev = same ? nullptr : B + f0 * nb * 3;
        if (space == MDAGPU_HOST) MDA_CUDA(pipe.acquire(piece_no, &a_dev, &b_dev));
        uint64_t total_cells;
        if (boxes_per_frame) {
            // re-base this chunk's cells at zero
            std::vector<FrameGeom> part(fgs.begin() + f0, fgs.begin() + f0 + fc);
            const int64_t b0 = part[0].cell_base;
            for (auto &p : part) p.cell_base -= b0;
            total_cells = (uint64_t)(part.back().cell_base + part.back().grid.ncell_total);
            MDA_CUDA(fg_dev.alloc(fc * sizeof(FrameGeom), st));
            MDA_CUDA(cudaMemcpyAsync(fg_dev.p, part.data(), fc * sizeof(FrameGeom), cudaMemcpyHostToDevice, st));
            MDA_CUDA(cudaStreamSynchronize(st));   // `part` dies at the end of this scope
        } else {
            total_cells = (uint64_t)cells_per_frame * fc;
            if (f0 == 0) {
                MDA_CUDA(fg_dev.alloc(sizeof(FrameGeom), st));
                MDA_CUDA(cudaMemcpyAsync(fg_dev.p, fgs.data(), sizeof(FrameGeom), cudaMemcpyHostToDevice, st));
            }
        }
        const FrameGeom *fgd = fg_dev.as<FrameGeom>();
        CellList cl, ql;
        // (none of the RDF traversals reads the cell key of a sorted atom: one scattered store less per atom)
        int rc = build_cell_list(same ? a_dev : b_dev, nb, fc, fgd, boxes_per_frame, total_cells, wrap, st, &cl, false);
        if (rc) return rc;
        if (!same) {
            rc = build_cell_list(a_dev, na, fc, fgd, boxes_per_frame,
                                 total_cells, wrap, st, &ql, false);
            if (rc) return rc;
        }
        const CellList &qq = same ? cl : ql;
        const double occupancy = use_grid ? (double)na * (double)fc / (double)total_cells : (double)na;

        Traversal tv;
        tv.row = W > 0;
        tv.quad = quad;
        tv.chunks = 1;
        const bool lagged = fineWX && tv.row && flag_slots().init() == cudaSuccess;
        if (!lagged) { const int rs = settle(0); if (rs) return rs; }       // these paths synchronise themselves
        if (tv.row && !lagged) {
            // an atom that the float64 cell arithmetic folded into a cell it does not lie
            // in (a coordinate the reference's wrap left ON the upper face of the cell, e.g.
            // -1e-7 + L == L in float32) breaks the per-cell image prediction: the per-pair
            // kernel takes over -- for orthorhombic cells on the reference's own grid
            unsigned bad[2] = { 0u, 0u };
            MDA_CUDA(cudaMemcpyAsync(&bad[0], cl.inconsistent(), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            if (!same) MDA_CUDA(cudaMemcpyAsync(&bad[1], ql.inconsistent(), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            MDA_CUDA(cudaStreamSynchronize(st));
            if (bad[0] | bad[1]) {
                if (getenv("MDAGPU_DEBUG")) fprintf(stderr, "mdagpu: rdf cell list has folded atoms, piece of %llu frames\n", (unsigned long long)fc);
                // (triclinic cells on the reference's own grid: its 27-cell per-pair kernel takes over in
                // place; on a grid of the library's own -- orthorhombic, or triclinic with W = 2 -- the
                // piece is recomputed on the reference's grid)
                if (mic == MIC_TRICLINIC && W < 2) tv.row = 0;
                else {
                    rc = rdf_accumulate_impl(a_dev, na, same ? nullptr : b_dev, nb, fc, boxes_per_frame ? box6 + 6 * f0 : box6,
                                             boxes_per_frame ? h9 + 9 * f0 : h9, boxes_per_frame, mic, wrap, cutoff, d2_max,
                                             use_grid, thresholds, nbins, r_lo, r_hi, (same && !report_self) ? 1 : excl_a,
                                         (same && !report_self) ? 1 : excl_b, ids_a_host, ids_b_host,
                                             cnt.as<unsigned long long>(), per_site, MDAGPU_DEVICE, stream, 1);
                    if (rc) return rc;
                    MDA_CUDA(release_piece(piece_no));
                    continue;
                }
            }
        }
        if (lagged) {
            // rowfine.cuh: a CTA owns R fine home cells of a (y, z) row (the whole row when its queries fit
            // 224 threads and the staged rows fit the tile); a planner kernel writes row descriptors and
            // sorted window records, the walker stages and bins.  Densities the two cannot take (tile,
            // record list or thread count exceeded somewhere) go to rowquad.cuh on its own grid.
            int nx_max = 0, rows_max = 0;
            for (uint64_t f = boxes_per_frame ? f0 : 0; f < (boxes_per_frame ? f0 + fc : 1); ++f) {
                const GridGeom &g = fgs[f].grid;
                nx_max = g.ncells[0] > nx_max ? g.ncells[0] : nx_max;
                rows_max = g.ncells[1] * g.ncells[2] > rows_max ? g.ncells[1] * g.ncells[2] : rows_max;
            }
            const double occ_q = (double)na * (double)fc / (double)total_cells;
            const double occ_c = (double)nb * (double)fc / (double)total_cells;
            const int nrows = same ? 13 : 25;
            const int list_cap = same ? 1664 : 3072;            // window records per unit (one per pair of queries and row)
            double q_target = 208.0;
            if (1.7 * list_cap / nrows < q_target) q_target = 1.7 * list_cap / nrows;
            int R = (int)(q_target / fmax(occ_q, 1e-9));
            const int r_tile = (int)(0.9 * RF_CAP / (nrows * fmax(occ_c, 1e-9))) - 2 * fineWX;
            if (R > r_tile) R = r_tile;
            if (R > RF_MAXEXT - 2 * fineWX) R = RF_MAXEXT - 2 * fineWX;
            if (R > nx_max) R = nx_max;
            if (R < 1) R = 1;
            const int runs = (nx_max + R - 1) / R;
            R = (nx_max + runs - 1) / runs;              // even split of the row
            FineParams fp;
            fp.W = 2; fp.WX = fineWX; fp.R = R; fp.reach = fine_reach;
            fp.debug = getenv("MDAGPU_FINE_DEBUG") ? atoi(getenv("MDAGPU_FINE_DEBUG")) : 0;
            const Derived dv = derive(err_row);
            SearchParams spc = sp;
            spc.d2_filter = dv.d2_filter;
            spc.include_self = report_self;
            const uint64_t units_x = (uint64_t)runs * (uint64_t)rows_max, n_units = units_x * fc;
            const size_t unit_bytes = same ? sizeof(FineUnit<13>) : sizeof(FineUnit<25>);
            DevBuf units, lists, oflag;
            MDA_CUDA(units.alloc(n_units * unit_bytes, st));
            MDA_CUDA(lists.alloc(n_units * (size_t)list_cap * sizeof(uint32_t), st));
            MDA_CUDA(oflag.alloc(sizeof(unsigned), st));
            MDA_CUDA(cudaMemsetAsync(oflag.p, 0, sizeof(unsigned), st));
            dim3 grid((unsigned)units_x, (unsigned)fc);
            const size_t plan_smem = (size_t)nrows * (size_t)(R + 2 * fineWX + 1) * sizeof(uint32_t);
#define FINE_PLAN(SELF_, MIC_, QL_, NR_, BAD1_)                                                                  \
    MDA_LAUNCH_AS("rdf_window_planner", (fine_plan_kernel<SELF_, 224, MIC_>), grid, 224, plan_smem, st, QL_.atoms(), QL_.start(), \
                  cl.start(), fgd, boxes_per_frame, fgs[boxes_per_frame ? f0 : 0], fp, RF_CAP, list_cap,          \
                  units.as<FineUnit<NR_>>(), lists.as<uint32_t>(), oflag.as<unsigned>(), cl.inconsistent(), BAD1_)
            if (same) {
                if (fine_tric) FINE_PLAN(true, MIC_TRICLINIC, cl, 13, (const unsigned *)nullptr);
                else FINE_PLAN(true, MIC_ORTHO_LAZY, cl, 13, (const unsigned *)nullptr);
            } else {
                if (fine_tric) FINE_PLAN(false, MIC_TRICLINIC, ql, 25, ql.inconsistent());
                else FINE_PLAN(false, MIC_ORTHO_LAZY, ql, 25, ql.inconsistent());
            }
#undef FINE_PLAN
            MDA_CHECK_LAUNCH();
            // lane columns of the histogram: 16 keep four CTAs on an SM (55 kB each)
            int ls = 4;
            while (ls > 0 && ((size_t)(nbins + 2) << ls) * sizeof(uint32_t) > 10 * 1024) --ls;
            if (getenv("MDAGPU_RDF_LS")) ls = atoi(getenv("MDAGPU_RDF_LS"));
            HistSinkFine<224> sink = { thr_in.as<double>(), cnt.as<unsigned long long>(), nbins, ls, c0, inv_dr, dv.band_lim };
            const size_t hist = ((((size_t)(nbins + 2) << ls) * sizeof(uint32_t)) + 15) & ~(size_t)15;
            const size_t total = hist + 3 * (size_t)RF_CAP * sizeof(float);
            // (per device and cheap: set on every call, like the other kernels of this file)
            const BoxParams bx = fgs[boxes_per_frame ? f0 : 0].box;
            // (per-frame boxes: the walker only needs the box for the rare float64 evaluations; frames of
            // one call that differ in their box are binned frame by frame)
#define FINE_WALK(SELF_, MIC_, QL_, NR_, BAD2_)                                                                  \
    do {                                                                                                        \
        auto kern = fine_walk_kernel<SELF_, HistSinkFine<224>, RF_CAP, 4, MIC_>;                                 \
        MDA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));           \
        if (boxes_per_frame) {                                                                                  \
            for (uint64_t ff = 0; ff < fc; ++ff) {                                                              \
                dim3 g1((unsigned)units_x, 1u);                                                                 \
                const BoxParams bxf = fgs[f0 + ff].box;                                                         \
                MDA_LAUNCH_AS("rdf_pair_histogram", kern, g1, 224, total, st, QL_.atoms(), cl.atoms(),          \
                              units.as<FineUnit<NR_>>() + ff * units_x, lists.as<uint32_t>() + ff * units_x * list_cap, \
                              list_cap, bxf, spc, sink, oflag.as<unsigned>(), cl.inconsistent(), BAD2_);         \
            }                                                                                                   \
        } else {                                                                                                \
            MDA_LAUNCH_AS("rdf_pair_histogram", kern, grid, 224, total, st, QL_.atoms(), cl.atoms(),            \
                          units.as<FineUnit<NR_>>(), lists.as<uint32_t>(), list_cap, bx, spc, sink,              \
                          oflag.as<unsigned>(), cl.inconsistent(), BAD2_);                                       \
        }                                                                                                       \
    } while (0)
            if (same) {
                if (fine_tric) FINE_WALK(true, MIC_TRICLINIC, cl, 13, (const unsigned *)nullptr);
                else FINE_WALK(true, MIC_ORTHO_LAZY, cl, 13, (const unsigned *)nullptr);
            } else {
                if (fine_tric) FINE_WALK(false, MIC_TRICLINIC, ql, 25, ql.inconsistent());
                else FINE_WALK(false, MIC_ORTHO_LAZY, ql, 25, ql.inconsistent());
            }
#undef FINE_WALK
            MDA_CHECK_LAUNCH();
            {
                // flags of this piece -> pinned slot; looked at after the NEXT piece has been queued
                FlagSlots &fs = flag_slots();
                const int slot = slots_used++ % FlagSlots::N;
                DevBuf f3;
                MDA_CUDA(f3.alloc(3 * sizeof(unsigned), st));
                MDA_LAUNCH(rdf_flags_kernel, 1, 1, 0, st, oflag.as<unsigned>(), cl.inconsistent(),
                           same ? (const unsigned *)nullptr : ql.inconsistent(), f3.as<unsigned>());
                MDA_CUDA(cudaMemcpyAsync(fs.host + 4 * slot, f3.p, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
                MDA_CUDA(cudaEventRecord(fs.ev[slot], st));
                pending.push_back({ piece_no, f0, fc, a_dev, b_dev, slot });
                const int rs = settle(1);
                if (rs) return rs;
                MDA_CUDA(pump());                        // (the first iterations: nothing released yet, buffers still free)
            }
            continue;
        }
        int threads;
        if (tv.row) {
            threads = 128;
            int nx_min = INT32_MAX, rows_max = 0;
            for (uint64_t f = boxes_per_frame ? f0 : 0; f < (boxes_per_frame ? f0 + fc : 1); ++f) {
                const GridGeom &g = fgs[f].grid;
                nx_min = g.ncells[0] < nx_min ? g.ncells[0] : nx_min;
                rows_max = g.ncells[1] * g.ncells[2] > rows_max ? g.ncells[1] * g.ncells[2] : rows_max;
            }
            int nx_max = 0;
            for (uint64_t f = boxes_per_frame ? f0 : 0; f < (boxes_per_frame ? f0 + fc : 1); ++f)
                nx_max = fgs[f].grid.ncells[0] > nx_max ? fgs[f].grid.ncells[0] : nx_max;
            // packed kernel: a 224-thread CTA takes the whole row of home cells when its queries fit
            // (no idle fourth warp in the short runs, one staging group, fewer staged cells per query)
            if (quad && nx_max <= ROW_MAXR && (double)nx_max * occupancy <= 208.0
                && (size_t)nbins * 7 * sizeof(uint32_t) <= 64 * 1024) threads = 224;
            int R = (int)((double)threads / fmax(occupancy, 1e-9));
            R = R < 1 ? 1 : (R > ROW_MAXR ? ROW_MAXR : R);
            if (R > nx_max) R = nx_max;
            const int runs = (nx_max + R - 1) / R;
            R = (nx_max + runs - 1) / runs;              // even split of the row
            tv.rp.W = W;
            tv.rp.R = R;
            tv.grid_x = (uint64_t)runs * (uint64_t)rows_max;
            (void)nx_min;
        } else {
            // threads per CTA ~ atoms per home cell (one query per thread)
            const uint64_t max_cells = boxes_per_frame ? total_cells / fc + 1 : (uint64_t)cells_per_frame;
            threads = occupancy > 128.0 ? 192 : (occupancy > 64.0 ? 128 : 64);
            tv.grid_x = max_cells;
            if (boxes_per_frame) {
                tv.grid_x = 0;
                for (uint64_t f = f0; f < f0 + fc; ++f)
                    tv.grid_x = (uint64_t)fgs[f].grid.ncell_total > tv.grid_x ? (uint64_t)fgs[f].grid.ncell_total : tv.grid_x;
            }
        }
        // one 32-bit histogram per warp in shared memory
        const size_t hist_bytes = (size_t)nbins * (threads / 32) * sizeof(uint32_t);
        const bool fits = !per_site && hist_bytes <= 160 * 1024;
        if (!fits) threads = HistSinkGlobal::THREADS;
        if (!tv.row && !use_grid) {                  // single cell: split the queries over CTAs
            tv.chunks = (int)((na + threads - 1) / threads);
            tv.grid_x = (uint64_t)tv.chunks;
        }
        const Derived dv = derive(tv.row ? err_row : err_cell);
        const float band_lim = dv.band_lim, d2_fast_max = dv.d2_fast_max;
        SearchParams spc = sp;
        spc.d2_filter = dv.d2_filter;
        spc.include_self = report_self;
#define COMMA ,
#define RUN_HIST(SINK)                                                                              \
    do {                                                                                           \
        SINK sink = { thr_in.as<double>(), cnt.as<unsigned long long>(), nbins, excl_a, excl_b, c0, inv_dr, band_lim, d2_fast_max, 0u, ida, idb }; \
        rc = dispatch_hist(mic, same, qq, cl, fgd, boxes_per_frame, fc, tv, spc, sink, hist_bytes, st); \
    } while (0)
        if (per_site) {
            HistSinkSites sink = { thr_in.as<double>(), cnt.as<unsigned long long>(), nbins, nb, (float)r_lo, inv_dr };
            rc = dispatch_hist(mic, same, qq, cl, fgd, boxes_per_frame, fc, tv, spc, sink, 0, st);
        }
        else if (!fits) {
            HistSinkGlobal sink = { thr_in.as<double>(), cnt.as<unsigned long long>(), nbins, excl_a, excl_b, (float)r_lo, inv_dr, ida, idb };
            rc = dispatch_hist(mic, same, qq, cl, fgd, boxes_per_frame, fc, tv, spc, sink, 0, st);
        }
        else if (excl_a > 0 || by_ids) {
            if (threads == 192) RUN_HIST(HistSink<192 COMMA true>);
            else if (threads == 128) RUN_HIST(HistSink<128 COMMA true>);
            else RUN_HIST(HistSink<64 COMMA true>);
        } else {
            if (threads == 224) RUN_HIST(HistSink<224 COMMA false>);
            else if (threads == 192) RUN_HIST(HistSink<192 COMMA false>);
            else if (threads == 128) RUN_HIST(HistSink<128 COMMA false>);
            else RUN_HIST(HistSink<64 COMMA false>);
        }
#undef RUN_HIST
#undef COMMA
        if (rc) return rc;
        MDA_CUDA(release_piece(piece_no));
    }
    { const int rs = settle(0); if (rs) return rs; }
    MDA_CUDA(cnt.finish());
    if (space == MDAGPU_HOST) MDA_CUDA(cudaStreamSynchronize(st));
    return MDAGPU_OK;
}

extern "C" int mdagpu_rdf_accumulate(const float *A, uint64_t na, const float *B, uint64_t nb,
                                     uint64_t nframes, const float *box6, const float *h9,
                                     int boxes_per_frame, int mic, int wrap, double cutoff,
                                     double d2_max, int use_grid, const double *thresholds,
                                     int nbins, double r_lo, double r_hi, int excl_a, int excl_b,
                                     unsigned long long *counts, int space, mdagpu_stream_t stream)
{
    return rdf_accumulate_impl(A, na, B, nb, nframes, box6, h9, boxes_per_frame, mic, wrap, cutoff, d2_max,
                               use_grid, thresholds, nbins, r_lo, r_hi, excl_a, excl_b, nullptr, nullptr, counts, 0,
                               space, stream);
}

extern "C" int mdagpu_rdf_accumulate_ids(const float *A, uint64_t na, const float *B, uint64_t nb,
                                         uint64_t nframes, const float *box6, const float *h9,
                                         int boxes_per_frame, int mic, int wrap, double cutoff,
                                         double d2_max, int use_grid, const double *thresholds,
                                         int nbins, double r_lo, double r_hi, const int32_t *ids_a,
                                         const int32_t *ids_b, unsigned long long *counts, int space,
                                         mdagpu_stream_t stream)
{
    MDA_REQUIRE(ids_a != nullptr, "mdagpu_rdf_accumulate_ids: ids_a is NULL");
    MDA_REQUIRE(ids_b != nullptr || B == nullptr || B == A, "mdagpu_rdf_accumulate_ids: ids_b is NULL");
    return rdf_accumulate_impl(A, na, B, nb, nframes, box6, h9, boxes_per_frame, mic, wrap, cutoff, d2_max,
                               use_grid, thresholds, nbins, r_lo, r_hi, 0, 0, ids_a, ids_b, counts, 0, space, stream);
}

extern "C" int mdagpu_rdf_sites_accumulate(const float *A, uint64_t na, const float *B, uint64_t nb,
                                           uint64_t nframes, const float *box6, const float *h9,
                                           int boxes_per_frame, int mic, int wrap, double cutoff,
                                           double d2_max, int use_grid, const double *thresholds,
                                           int nbins, double r_lo, double r_hi,
                                           unsigned long long *counts, int space, mdagpu_stream_t stream)
{
    MDA_REQUIRE(na * nb <= (1ull << 26), "mdagpu_rdf_sites_accumulate: %llu x %llu site pairs is too many "
                "(one histogram each)", (unsigned long long)na, (unsigned long long)(B && B != A ? nb : na));
    return rdf_accumulate_impl(A, na, B, nb, nframes, box6, h9, boxes_per_frame, mic, wrap, cutoff, d2_max,
                               use_grid, thresholds, nbins, r_lo, r_hi, 0, 0, nullptr, nullptr, counts, 1, space, stream);
}

extern "C" int mdagpu_contacts_self(const float *xyz, uint64_t n, uint64_t nframes, int box_mode,
                                    const float *box, double d2_max, double cutoff,
                                    unsigned long long *counts, int32_t *degree, int space,
                                    mdagpu_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(xyz && counts, "mdagpu_contacts_self: NULL argument");
    MDA_REQUIRE(box_mode >= 0 && box_mode <= 2, "bad box_mode %d", box_mode);
    MDA_REQUIRE(box_mode == BOX_NONE || box != nullptr, "box is NULL");
    MDA_REQUIRE(nframes <= 65535, "mdagpu_contacts_self: at most 65535 frames per call");
    if (nframes == 0) return MDAGPU_OK;

    StagedIn in;
    StagedOut cnt, deg;
    MDA_CUDA(in.init(xyz, nframes * n * 3 * sizeof(float), space, st));
    MDA_CUDA(cnt.init(counts, nframes * sizeof(unsigned long long), space, st));
    MDA_CUDA(cudaMemsetAsync(cnt.dev, 0, nframes * sizeof(unsigned long long), st));
    if (degree) {
        MDA_CUDA(deg.init(degree, nframes * n * sizeof(int32_t), space, st));
        MDA_CUDA(cudaMemsetAsync(deg.dev, 0, nframes * n * sizeof(int32_t), st));
    }
    if (n >= 2) {
        FrameGeom fg;
        fg.cell_base = 0;
        fg.box = make_box_params(box_mode, box);
        int mic = MIC_NONE, wrap = 0, stencil = 1;
        if (box_mode == BOX_ORTHO) {
            // lib.distances flavour on unwrapped input: bin by the wrapped position,
            // evaluate on the coordinates as given
            mic = MIC_ORTHO_ROUND; wrap = 2;
            grid_geometry_ortho_free(cutoff, box, nullptr, 1, &fg.grid);
        } else if (box_mode == BOX_TRICLINIC) {
            mic = MIC_TRICLINIC; wrap = 1; stencil = 0;
            grid_geometry_single(&fg.grid);
        } else {
            DevBuf lohi;
            MDA_CUDA(lohi.alloc(6 * sizeof(float), st));
            const float init[6] = { FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX };
            MDA_CUDA(cudaMemcpyAsync(lohi.p, init, sizeof init, cudaMemcpyHostToDevice, st));
            unsigned grid = cdiv(nframes * n, 256);
            if (grid > (unsigned)num_sms() * 8) grid = (unsigned)num_sms() * 8;
            MDA_LAUNCH(bbox_kernel, grid, 256, 0, st, in.as<float>(), nframes * n, lohi.as<float>());
            MDA_CHECK_LAUNCH();
            float h[6];
            MDA_CUDA(cudaMemcpyAsync(h, lohi.p, sizeof h, cudaMemcpyDeviceToHost, st));
            MDA_CUDA(cudaStreamSynchronize(st));
            float L[3];
            double origin[3];
            for (int k = 0; k < 3; ++k) {
                const double pad = 1e-3 + 1e-5 * fmax(fabs((double)h[k]), fabs((double)h[3 + k]));
                origin[k] = (double)h[k] - pad;
                L[k] = (float)((double)h[3 + k] - (double)h[k] + 2.0 * pad + cutoff * 1e-3);
            }
            grid_geometry_ortho_free(cutoff, L, origin, 0, &fg.grid);
        }
        DevBuf fg_dev;
        MDA_CUDA(fg_dev.alloc(sizeof(FrameGeom), st));
        MDA_CUDA(cudaMemcpyAsync(fg_dev.p, &fg, sizeof(FrameGeom), cudaMemcpyHostToDevice, st));
        CellList cl;
        int rc = build_cell_list(in.as<float>(), n, nframes, fg_dev.as<FrameGeom>(), 0,
                                 (uint64_t)fg.grid.ncell_total * nframes, wrap, st, &cl);
        if (rc) return rc;
        SearchParams sp;
        sp.d2_max = d2_max;
        sp.d2_min = -1.0;
        sp.d2_filter = filter_threshold(d2_max, mic == MIC_TRICLINIC ? tric_abs_err(fg.box.h) : 0.0);
        sp.use_filter = mic != MIC_ORTHO_ROUND;
        sp.stencil = stencil;
        CountSink sink = { cnt.as<unsigned long long>(), degree ? deg.as<int>() : nullptr, n, 0u };
        dim3 grid(cdiv(n, CountSink::THREADS), (unsigned)nframes);
        const FrameGeom *fgd = fg_dev.as<FrameGeom>();
        switch (mic) {
        case MIC_NONE:
            MDA_LAUNCH_AS("contacts_pair_count", (pair_search_kernel<MIC_NONE, true, CountSink>), grid, CountSink::THREADS, 0, st, cl.atoms(), cl.key_of_sorted(), n, cl.atoms(), cl.start(), fgd, 0, sp, sink);
            break;
        case MIC_ORTHO_ROUND:
            MDA_LAUNCH_AS("contacts_pair_count", (pair_search_kernel<MIC_ORTHO_ROUND, true, CountSink>), grid, CountSink::THREADS, 0, st, cl.atoms(), cl.key_of_sorted(), n, cl.atoms(), cl.start(), fgd, 0, sp, sink);
            break;
        default:
            MDA_LAUNCH_AS("contacts_pair_count", (pair_search_kernel<MIC_TRICLINIC, true, CountSink>), grid, CountSink::THREADS, 0, st, cl.atoms(), cl.key_of_sorted(), n, cl.atoms(), cl.start(), fgd, 0, sp, sink);
            break;
        }
        MDA_CHECK_LAUNCH();
        MDA_CUDA(cudaStreamSynchronize(st));   // scratch (cell list) is released on return
    }
    MDA_CUDA(cnt.finish());
    if (degree) MDA_CUDA(deg.finish());
    return MDAGPU_OK;
}

