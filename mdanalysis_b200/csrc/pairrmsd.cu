// pairrmsd.cu -- all-pairs RMSD matrix of a trajectory for sm_100a.
//
// Reference: analysis/diffusionmap.py:238-268 (DistanceMatrix._single_frame: a Python loop
// over frame pairs, one rms.rmsd() call each) with the metric rms.rmsd, analysis/rms.py:178-283:
//   superposition=False (the DistanceMatrix default):  sqrt(sum(w (a - b)^2) / N), float64
//   superposition=True:  both frames centred (weighted mean), QCP (qcprot.pyx:326-543)
//
// One CTA owns a 32 x 32 tile of frame pairs (tiles below the diagonal are not launched: the
// matrix is symmetric), a thread a 2 x 2 block of it.  The selected atoms are streamed through
// shared memory in chunks of KA atoms, widened to float64 once at staging time (centred for the
// superposition flavour, scaled by sqrt(w_k) so that the weight costs nothing in the inner loop),
// laid out coordinate-major so that the two frames of a thread are one 16-byte load: the row
// operand is a broadcast, the column operand is conflict-free.  Inner loop per coordinate:
//   no fit:  d = a - b; acc = fma(d, d, acc)                  (2 FP64 ops per pair)
//   QCP:     nine inner products a_c1 * b_c2                  (9 FP64 ops per pair and atom)
// then one qcp_solve() per pair in the epilogue.  FP64-pipe bound; the trajectory is re-read
// F/32 times from L2.
#include "common.cuh"

#include <vector>

namespace mdagpu {

constexpr int PR_T = 32;          // frames per tile side
constexpr int PR_KA = 16;         // atoms per staged chunk
constexpr int PR_LD = PR_T + 8;   // shared row pitch in doubles (see pair_rmsd_stage)
constexpr int PR_THREADS = 256;   // 16 x 16 threads, 2 x 2 pairs each

struct FrameCentre { double c[3]; double g; };   // weighted centroid, G = sum w |x - c|^2

// per-frame weighted centroid and inner product (np.average(a, axis=0, weights=w), rms.py:262-266)
__global__ void __launch_bounds__(256)
pair_rmsd_centre_kernel(const float *__restrict__ traj, uint64_t natoms, const int32_t *__restrict__ sel,
                        uint64_t nsel, const double *__restrict__ w, FrameCentre *__restrict__ out)
{
    const float *x = traj + (uint64_t)blockIdx.x * natoms * 3;
    __shared__ double red[4][8];
    __shared__ double centre[3];
    double s[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (uint64_t k = threadIdx.x; k < nsel; k += blockDim.x) {
        const uint64_t a = sel ? (uint64_t)sel[k] : k;
        const double wk = w ? w[k] : 1.0;
        s[0] += wk * (double)x[3 * a]; s[1] += wk * (double)x[3 * a + 1]; s[2] += wk * (double)x[3 * a + 2];
        s[3] += wk;
    }
    for (int q = 0; q < 4; ++q) {
        for (int o = 16; o; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = s[q];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0, tw = 0.0;
        for (int i = 0; i < 8; ++i) { t += red[threadIdx.x][i]; tw += red[3][i]; }
        centre[threadIdx.x] = t / tw;
    }
    __syncthreads();
    const double cx = centre[0], cy = centre[1], cz = centre[2];
    double g = 0.0;
    for (uint64_t k = threadIdx.x; k < nsel; k += blockDim.x) {
        const uint64_t a = sel ? (uint64_t)sel[k] : k;
        const double wk = w ? w[k] : 1.0;
        const double dx = (double)x[3 * a] - cx, dy = (double)x[3 * a + 1] - cy, dz = (double)x[3 * a + 2] - cz;
        g += wk * (dx * dx + dy * dy + dz * dz);
    }
    for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = g;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[0][i];
        out[blockIdx.x].c[0] = cx; out[blockIdx.x].c[1] = cy; out[blockIdx.x].c[2] = cz;
        out[blockIdx.x].g = t;
    }
}

// Staging.  A thread always serves the same frame of the tile and the same six coordinate slots
// of a chunk (slot = kc0 + 8 r): a warp step covers 8 frames x 4 consecutive coordinates (16-byte
// runs of global memory) and, with rows of PR_LD doubles, its 8-byte shared stores are
// conflict-free per half warp.  The values of the NEXT chunk are fetched into registers before
// the current chunk is computed, so the global-load latency is covered by the arithmetic.
// TF frames per tile side, KC coordinates per chunk, LD shared row pitch in doubles.
template <int TF, int KC, int LD>
struct PairStager {
    static constexpr int NG = TF / 8;                    // warps side by side along the frames
    static constexpr int STEP = 4 * (PR_THREADS / 32 / NG);   // coordinates covered per step of the CTA
    static constexpr int SLOTS = KC / STEP;
    static_assert(KC % STEP == 0 && TF % 8 == 0, "tile shape");
    const float *row;       // first coordinate of the thread's frame (nullptr: frame outside the trajectory)
    int fr, kc0;
    const double *c;        // centroid of the frame (superposition flavour), read when needed
    __device__ __forceinline__ void init(const float *traj, uint64_t nframes, uint64_t natoms, uint64_t f0,
                                         const FrameCentre *centres)
    {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        fr = (w % NG) * 8 + (lane & 7);
        kc0 = (w / NG) * 4 + (lane >> 3);
        const uint64_t f = f0 + fr;
        row = f < nframes ? traj + f * natoms * 3 : nullptr;
        c = (centres && row) ? centres[f].c : nullptr;
    }
    // raw float32 values of chunk k0 (0 outside the selection / trajectory); k0 counts atoms
    __device__ __forceinline__ void fetch(float (&v)[SLOTS], const int32_t *__restrict__ sel, uint64_t nsel,
                                          uint64_t k0) const
    {
#pragma unroll
        for (int r = 0; r < SLOTS; ++r) {
            const int kc = kc0 + STEP * r;
            const uint64_t k = k0 + kc / 3;
            v[r] = 0.0f;
            if (row && k < nsel) v[r] = sel ? row[(uint64_t)sel[k] * 3 + kc % 3] : row[k0 * 3 + kc];
        }
    }
    // s[(atom * 3 + comp) * LD + frame] = sqrt(w) * (x - centre)
    template <bool FIT>
    __device__ __forceinline__ void store(double *s, const float (&v)[SLOTS], const double *__restrict__ sqw,
                                          uint64_t nsel, uint64_t k0) const
    {
#pragma unroll
        for (int r = 0; r < SLOTS; ++r) {
            const int kc = kc0 + STEP * r;
            const uint64_t k = k0 + kc / 3;
            double x = 0.0;
            if (row && k < nsel) {
                x = (double)v[r];
                if (FIT) x -= c[kc % 3];
                if (sqw) x *= sqw[k];
            }
            s[kc * LD + fr] = x;
        }
    }
};
using Stager32 = PairStager<PR_T, PR_KA * 3, PR_LD>;
constexpr int PR_SLOTS = Stager32::SLOTS;

template <bool FIT>
__global__ void __launch_bounds__(PR_THREADS, FIT ? 2 : 3)
pair_rmsd_kernel(const float *__restrict__ traj, uint64_t nframes, uint64_t natoms,
                 const int32_t *__restrict__ sel, uint64_t nsel, const double *__restrict__ sqw,
                 const FrameCentre *__restrict__ centres, double cutoff, double *__restrict__ out)
{
    // upper-triangular tile index -> (ti, tj), tj >= ti
    const unsigned nt = (unsigned)((nframes + PR_T - 1) / PR_T);
    unsigned ti = 0, rest = blockIdx.x;
    while (rest >= nt - ti) { rest -= nt - ti; ++ti; }
    const unsigned tj = ti + rest;
    const uint64_t i0 = (uint64_t)ti * PR_T, j0 = (uint64_t)tj * PR_T;
    const bool offdiag = tj != ti;

    __shared__ __align__(16) double sa[PR_KA * 3 * PR_LD];
    __shared__ __align__(16) double sb[PR_KA * 3 * PR_LD];
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;   // rows 2ty, 2ty+1; columns 2tx, 2tx+1

    constexpr int NACC = FIT ? 9 : 1;
    double acc[4][NACC];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[p][q] = 0.0;

    Stager32 sta, stb;
    sta.init(traj, nframes, natoms, i0, FIT ? centres : nullptr);
    stb.init(traj, nframes, natoms, j0, FIT ? centres : nullptr);
    // the nine-accumulator flavour has no registers to spare: it fetches just before it stores
    constexpr bool PREFETCH = !FIT;
    float va[PR_SLOTS], vb[PR_SLOTS];
    if (PREFETCH) {
        sta.fetch(va, sel, nsel, 0);
        if (offdiag) stb.fetch(vb, sel, nsel, 0);
    }
    const double *pa = sa + 2 * ty;
    const double *pb = (offdiag ? sb : sa) + 2 * tx;

    for (uint64_t k0 = 0; k0 < nsel; k0 += PR_KA) {
        if (!PREFETCH) {
            sta.fetch(va, sel, nsel, k0);
            if (offdiag) stb.fetch(vb, sel, nsel, k0);
        }
        __syncthreads();                       // the previous chunk has been consumed
        sta.template store<FIT>(sa, va, sqw, nsel, k0);
        if (offdiag) stb.template store<FIT>(sb, vb, sqw, nsel, k0);
        __syncthreads();
        if (PREFETCH && k0 + PR_KA < nsel) {   // in flight while this chunk is computed
            sta.fetch(va, sel, nsel, k0 + PR_KA);
            if (offdiag) stb.fetch(vb, sel, nsel, k0 + PR_KA);
        }
        if constexpr (!FIT) {
#pragma unroll 8
            for (int kc = 0; kc < PR_KA * 3; ++kc) {
                const double2 a = *reinterpret_cast<const double2 *>(pa + kc * PR_LD);
                const double2 b = *reinterpret_cast<const double2 *>(pb + kc * PR_LD);
                double d;
                d = a.x - b.x; acc[0][0] = fma(d, d, acc[0][0]);
                d = a.x - b.y; acc[1][0] = fma(d, d, acc[1][0]);
                d = a.y - b.x; acc[2][0] = fma(d, d, acc[2][0]);
                d = a.y - b.y; acc[3][0] = fma(d, d, acc[3][0]);
            }
        } else {
#pragma unroll 2
            for (int ka = 0; ka < PR_KA; ++ka) {
                double2 a[3], b[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    a[c] = *reinterpret_cast<const double2 *>(pa + (ka * 3 + c) * PR_LD);
                    b[c] = *reinterpret_cast<const double2 *>(pb + (ka * 3 + c) * PR_LD);
                }
#pragma unroll
                for (int c1 = 0; c1 < 3; ++c1)
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) {
                        const int q = c1 * 3 + c2;
                        acc[0][q] = fma(a[c1].x, b[c2].x, acc[0][q]);
                        acc[1][q] = fma(a[c1].x, b[c2].y, acc[1][q]);
                        acc[2][q] = fma(a[c1].y, b[c2].x, acc[2][q]);
                        acc[3][q] = fma(a[c1].y, b[c2].y, acc[3][q]);
                    }
            }
        }
    }

    const double n = (double)nsel;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const uint64_t i = i0 + 2 * ty + (p >> 1), j = j0 + 2 * tx + (p & 1);
        if (i >= nframes || j >= nframes) continue;
        double r;
        if constexpr (!FIT) {
            r = sqrt(acc[p][0] / n);
        } else {
            // E0 = (G_a + G_b) / 2 (qcprot.pyx:305-318), A = sum w a b^T
            r = qcp_solve(nullptr, acc[p], 0.5 * (centres[i].g + centres[j].g), n);
        }
        r = r > cutoff ? r : 0.0;                 // diffusionmap.py:262-263
        if (offdiag) {
            out[i * nframes + j] = r;
            out[j * nframes + i] = r;             // diffusionmap.py:264-267
        } else if (j >= i) {
            // inside a diagonal tile both (i, j) and (j, i) are computed; the reference fills the
            // lower triangle from the upper one
            out[i * nframes + j] = r;
            out[j * nframes + i] = r;
        }
    }
}

// The plain metric on 64 x 64 tiles, a thread owning 4 x 4 pairs.  With 2 x 2 pairs the kernel above
// is bound by the shared-memory instruction queue (ncu: mio_throttle 3.9, FP64 pipe 48 %): two LDS.128
// feed eight FP64 instructions.  Here four LDS.128 feed thirty-two.
constexpr int PR_T2 = 64;          // frames per tile side
constexpr int PR_KA2 = 8;          // atoms per staged chunk
constexpr int PR_LD2 = PR_T2 + 8;  // pitch: 144 words = 16 mod 32, conflict-free 8-byte stores per half warp
using Stager64 = PairStager<PR_T2, PR_KA2 * 3, PR_LD2>;

__global__ void __launch_bounds__(PR_THREADS, 2)
pair_rmsd_plain_kernel(const float *__restrict__ traj, uint64_t nframes, uint64_t natoms,
                       const int32_t *__restrict__ sel, uint64_t nsel, const double *__restrict__ sqw,
                       double cutoff, double *__restrict__ out)
{
    const unsigned nt = (unsigned)((nframes + PR_T2 - 1) / PR_T2);
    unsigned ti = 0, rest = blockIdx.x;
    while (rest >= nt - ti) { rest -= nt - ti; ++ti; }
    const unsigned tj = ti + rest;
    const uint64_t i0 = (uint64_t)ti * PR_T2, j0 = (uint64_t)tj * PR_T2;
    const bool offdiag = tj != ti;

    __shared__ __align__(16) double sa[PR_KA2 * 3 * PR_LD2];
    __shared__ __align__(16) double sb[PR_KA2 * 3 * PR_LD2];
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;   // rows 4ty..4ty+3; columns 2tx, 2tx+1, 32+2tx, 33+2tx

    double acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;

    Stager64 sta, stb;
    sta.init(traj, nframes, natoms, i0, nullptr);
    stb.init(traj, nframes, natoms, j0, nullptr);
    float va[Stager64::SLOTS], vb[Stager64::SLOTS];
    sta.fetch(va, sel, nsel, 0);
    if (offdiag) stb.fetch(vb, sel, nsel, 0);
    const double *pa = sa + 4 * ty;
    const double *pb = (offdiag ? sb : sa) + 2 * tx;

    for (uint64_t k0 = 0; k0 < nsel; k0 += PR_KA2) {
        __syncthreads();                       // the previous chunk has been consumed
        sta.store<false>(sa, va, sqw, nsel, k0);
        if (offdiag) stb.store<false>(sb, vb, sqw, nsel, k0);
        __syncthreads();
        if (k0 + PR_KA2 < nsel) {              // in flight while this chunk is computed
            sta.fetch(va, sel, nsel, k0 + PR_KA2);
            if (offdiag) stb.fetch(vb, sel, nsel, k0 + PR_KA2);
        }
#pragma unroll 4
        for (int kc = 0; kc < PR_KA2 * 3; ++kc) {
            const double2 a01 = *reinterpret_cast<const double2 *>(pa + kc * PR_LD2);
            const double2 a23 = *reinterpret_cast<const double2 *>(pa + kc * PR_LD2 + 2);
            const double2 b01 = *reinterpret_cast<const double2 *>(pb + kc * PR_LD2);
            const double2 b23 = *reinterpret_cast<const double2 *>(pb + kc * PR_LD2 + 32);
            const double av[4] = { a01.x, a01.y, a23.x, a23.y }, bv[4] = { b01.x, b01.y, b23.x, b23.y };
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double d = av[p] - bv[q];
                    acc[p][q] = fma(d, d, acc[p][q]);
                }
        }
    }

    const double n = (double)nsel;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint64_t i = i0 + 4 * ty + p, j = j0 + 2 * tx + (q & 1) + 32 * (q >> 1);
            if (i >= nframes || j >= nframes) continue;
            if (!offdiag && j < i) continue;   // the lower triangle of a diagonal tile is filled from the upper
            double r = sqrt(acc[p][q] / n);
            r = r > cutoff ? r : 0.0;          // diffusionmap.py:262-263
            out[i * nframes + j] = r;
            out[j * nframes + i] = r;          // diffusionmap.py:264-267
        }
}

}  // namespace mdagpu

using namespace mdagpu;

extern "C" int mdagpu_pairwise_rmsd(const float *traj, uint64_t nframes, uint64_t natoms, const int32_t *sel,
                                    uint64_t nsel, const double *weights, int superposition, double cutoff,
                                    double *out, int space, mdagpu_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(traj != nullptr && out != nullptr, "mdagpu_pairwise_rmsd: NULL trajectory or output");
    MDA_REQUIRE(nsel > 0 && nsel <= natoms, "mdagpu_pairwise_rmsd: nsel must be in 1..natoms");
    MDA_REQUIRE(sel != nullptr || nsel <= natoms, "mdagpu_pairwise_rmsd: bad selection");
    if (nframes == 0) return MDAGPU_OK;
    const uint64_t nt = (nframes + PR_T - 1) / PR_T;
    MDA_REQUIRE(nt * (nt + 1) / 2 <= 0x7fffffffull, "mdagpu_pairwise_rmsd: too many frames (%llu)",
                (unsigned long long)nframes);
    StagedIn t, s, w;
    StagedOut o;
    MDA_CUDA(t.init(traj, nframes * natoms * 3 * sizeof(float), space, st));
    MDA_CUDA(s.init(sel, nsel * sizeof(int32_t), space, st));
    MDA_CUDA(o.init(out, nframes * nframes * sizeof(double), space, st));
    // sqrt of the relative weights: always prepared on the host (nsel doubles)
    DevBuf sqw;
    std::vector<double> hw;
    const double *w_dev = nullptr;
    if (weights) {
        hw.resize(nsel);
        if (space == MDAGPU_DEVICE)
            MDA_CUDA(cudaMemcpyAsync(hw.data(), weights, nsel * sizeof(double), cudaMemcpyDeviceToHost, st));
        else
            memcpy(hw.data(), weights, nsel * sizeof(double));
        MDA_CUDA(cudaStreamSynchronize(st));
        MDA_CUDA(w.init(hw.data(), nsel * sizeof(double), MDAGPU_HOST, st));   // plain weights for the centroids
        w_dev = w.as<double>();
    }
    std::vector<double> hs;
    if (weights) {
        hs.resize(nsel);
        for (uint64_t k = 0; k < nsel; ++k) {
            MDA_REQUIRE(hw[k] >= 0.0, "mdagpu_pairwise_rmsd: negative weight");
            hs[k] = sqrt(hw[k]);
        }
        MDA_CUDA(sqw.alloc(nsel * sizeof(double), st));
        MDA_CUDA(cudaMemcpyAsync(sqw.p, hs.data(), nsel * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    DevBuf centres;
    if (superposition) {
        MDA_CUDA(centres.alloc(nframes * sizeof(FrameCentre), st));
        MDA_LAUNCH(pair_rmsd_centre_kernel, (unsigned)nframes, 256, 0, st, t.as<float>(), natoms, s.as<int32_t>(), nsel,
                   w_dev, centres.as<FrameCentre>());
        MDA_CHECK_LAUNCH();
    }
    if (superposition) {
        const unsigned grid = (unsigned)(nt * (nt + 1) / 2);
        MDA_LAUNCH_AS("pair_rmsd_kernel", pair_rmsd_kernel<true>, grid, PR_THREADS, 0, st, t.as<float>(), nframes, natoms,
                      s.as<int32_t>(), nsel, weights ? sqw.as<double>() : nullptr, centres.as<FrameCentre>(), cutoff,
                      o.as<double>());
    } else {
        const uint64_t nt2 = (nframes + PR_T2 - 1) / PR_T2;
        MDA_LAUNCH_AS("pair_rmsd_kernel", pair_rmsd_plain_kernel, (unsigned)(nt2 * (nt2 + 1) / 2), PR_THREADS, 0, st,
                      t.as<float>(), nframes, natoms, s.as<int32_t>(), nsel, weights ? sqw.as<double>() : nullptr, cutoff,
                      o.as<double>());
    }
    MDA_CHECK_LAUNCH();
    MDA_CUDA(o.finish());
    MDA_CUDA(cudaStreamSynchronize(st));   // host-side weight buffers are sources of async copies
    return MDAGPU_OK;
}
