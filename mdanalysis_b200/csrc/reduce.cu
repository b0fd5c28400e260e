// reduce.cu -- the small per-frame / over-frame reductions that follow the hot kernels inside the
// analyses, kept on the device so that only the result crosses PCIe:
//   mdagpu_frame_mean        AverageStructure: float64 mean over frames of the fitted atoms
//                            (analysis/align.py:1072-1080 accumulates positions, :1083 divides)
//   mdagpu_group_rmsd        RMSD groupselections: plain RMSD of a group after the fit of `select`
//                            (analysis/rms.py:716-721 -> rmsd(..., center=False, superposition=False))
//   mdagpu_contact_fraction  Contacts: hard_cut_q / radius_cut_q / soft_cut_q of a frame's native
//                            pair distances (analysis/contacts.py:231-312)
#include "common.cuh"

namespace mdagpu {

constexpr int RD_THREADS = 256;

__device__ __forceinline__ double rd_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double rd_block_sum(double v, double *scratch /* [RD_THREADS / 32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = rd_warp_sum(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < RD_THREADS / 32; ++w) s += scratch[w];
    __syncthreads();
    return s;
}

// partial[c][i] = sum over the frames of chunk c of component i of the selected atoms, added in
// frame order; frame_mean_finish adds the chunks in order and divides: one fixed summation tree.
__global__ void __launch_bounds__(RD_THREADS)
frame_mean_partial_kernel(const float *__restrict__ traj, uint64_t nframes, uint64_t natoms,
                          const int32_t *__restrict__ sel, uint64_t nsel, uint64_t frames_per_chunk,
                          double *__restrict__ partial)
{
    const uint64_t i = (uint64_t)blockIdx.x * RD_THREADS + threadIdx.x;
    if (i >= nsel * 3) return;
    const uint64_t a = i / 3, k = i - 3 * a;
    const uint64_t src = (sel ? (uint64_t)sel[a] : a) * 3 + k;
    const uint64_t f0 = (uint64_t)blockIdx.y * frames_per_chunk;
    const uint64_t f1 = f0 + frames_per_chunk < nframes ? f0 + frames_per_chunk : nframes;
    double s = 0.0;
    const float *p = traj + f0 * natoms * 3 + src;
    for (uint64_t f = f0; f < f1; ++f, p += natoms * 3) s += (double)*p;
    partial[(uint64_t)blockIdx.y * nsel * 3 + i] = s;
}

__global__ void __launch_bounds__(RD_THREADS)
frame_mean_finish_kernel(const double *__restrict__ partial, uint64_t nchunks, uint64_t n3, double nframes,
                         double *__restrict__ mean)
{
    const uint64_t i = (uint64_t)blockIdx.x * RD_THREADS + threadIdx.x;
    if (i >= n3) return;
    double s = 0.0;
    for (uint64_t c = 0; c < nchunks; ++c) s += partial[c * n3 + i];
    mean[i] = s / nframes;
}

// one CTA per frame: sqrt(sum_i w_i |x[f, idx_i] - ref_i|^2 / g)
__global__ void __launch_bounds__(RD_THREADS)
group_rmsd_kernel(const float *__restrict__ traj, uint64_t natoms, const int32_t *__restrict__ idx, uint64_t g,
                  const double *__restrict__ ref, const double *__restrict__ w, double *__restrict__ out)
{
    __shared__ double scratch[RD_THREADS / 32];
    const float *frame = traj + (uint64_t)blockIdx.x * natoms * 3;
    double s = 0.0;
    for (uint64_t i = threadIdx.x; i < g; i += RD_THREADS) {
        const uint64_t a = (uint64_t)idx[i];
        const double dx = (double)frame[3 * a] - ref[3 * i], dy = (double)frame[3 * a + 1] - ref[3 * i + 1],
                     dz = (double)frame[3 * a + 2] - ref[3 * i + 2];
        const double wi = w ? w[i] : 1.0;
        s += wi * (dx * dx) + wi * (dy * dy) + wi * (dz * dz);
    }
    s = rd_block_sum(s, scratch);
    if (threadIdx.x == 0) out[blockIdx.x] = sqrt(s / (double)g);
}

// one CTA per frame over its P native-contact distances
__global__ void __launch_bounds__(RD_THREADS)
contact_fraction_kernel(const double *__restrict__ r, uint64_t P, const double *__restrict__ r0, int method,
                        double radius, double beta, double lambda, double *__restrict__ out)
{
    __shared__ double scratch[RD_THREADS / 32];
    const double *row = r + (uint64_t)blockIdx.x * P;
    double s = 0.0;
    for (uint64_t i = threadIdx.x; i < P; i += RD_THREADS) {
        const double d = row[i];
        if (method == 0) s += d <= r0[i] ? 1.0 : 0.0;                             // hard_cut_q
        else if (method == 1) s += d <= radius ? 1.0 : 0.0;                       // radius_cut_q
        else s += 1.0 / (1.0 + exp(beta * (d - lambda * r0[i])));                 // soft_cut_q
    }
    s = rd_block_sum(s, scratch);
    if (threadIdx.x == 0) out[blockIdx.x] = s / (double)P;
}

}  // namespace mdagpu

using namespace mdagpu;

extern "C" int mdagpu_frame_mean(const float *traj, uint64_t nframes, uint64_t natoms, const int32_t *sel,
                                 uint64_t nsel, double *mean, int space, mdagpu_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(traj && mean, "mdagpu_frame_mean: NULL argument");
    MDA_REQUIRE(nframes > 0 && nsel > 0 && nsel <= natoms, "mdagpu_frame_mean: empty input or bad selection");
    StagedIn t, s;
    StagedOut m;
    MDA_CUDA(t.init(traj, nframes * natoms * 3 * sizeof(float), space, st));
    MDA_CUDA(s.init(sel, nsel * sizeof(int32_t), space, st));
    MDA_CUDA(m.init(mean, nsel * 3 * sizeof(double), space, st));
    // enough CTAs to pull HBM rate: (columns / 256) x chunks >= ~8 per SM, at least 8 frames a chunk
    const uint64_t cols = cdiv(nsel * 3, RD_THREADS);
    uint64_t nchunks = ((uint64_t)num_sms() * 8 + cols - 1) / cols;
    if (nchunks > (nframes + 7) / 8) nchunks = (nframes + 7) / 8;
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 65535) nchunks = 65535;
    const uint64_t per = (nframes + nchunks - 1) / nchunks;
    nchunks = (nframes + per - 1) / per;
    DevBuf partial;
    MDA_CUDA(partial.alloc(nchunks * nsel * 3 * sizeof(double), st));
    MDA_LAUNCH(frame_mean_partial_kernel, dim3((unsigned)cols, (unsigned)nchunks), RD_THREADS, 0, st, t.as<float>(), nframes,
               natoms, sel ? s.as<int32_t>() : nullptr, nsel, per, partial.as<double>());
    MDA_CHECK_LAUNCH();
    MDA_LAUNCH(frame_mean_finish_kernel, (unsigned)cols, RD_THREADS, 0, st, partial.as<double>(), nchunks, nsel * 3,
               (double)nframes, m.as<double>());
    MDA_CHECK_LAUNCH();
    MDA_CUDA(m.finish());
    return MDAGPU_OK;
}

extern "C" int mdagpu_group_rmsd(const float *traj, uint64_t nframes, uint64_t natoms, const int32_t *idx,
                                 uint64_t g, const double *ref, const double *weights, double *out, int space,
                                 mdagpu_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(traj && idx && ref && out, "mdagpu_group_rmsd: NULL argument");
    MDA_REQUIRE(g > 0, "mdagpu_group_rmsd: empty group");
    MDA_REQUIRE(nframes <= 0x7fffffffull, "mdagpu_group_rmsd: too many frames in one call");
    if (nframes == 0) return MDAGPU_OK;
    StagedIn t, i, r, w;
    StagedOut o;
    MDA_CUDA(t.init(traj, nframes * natoms * 3 * sizeof(float), space, st));
    MDA_CUDA(i.init(idx, g * sizeof(int32_t), space, st));
    MDA_CUDA(r.init(ref, g * 3 * sizeof(double), space, st));
    MDA_CUDA(w.init(weights, g * sizeof(double), space, st));
    MDA_CUDA(o.init(out, nframes * sizeof(double), space, st));
    MDA_LAUNCH(group_rmsd_kernel, (unsigned)nframes, RD_THREADS, 0, st, t.as<float>(), natoms, i.as<int32_t>(), g,
               r.as<double>(), weights ? w.as<double>() : nullptr, o.as<double>());
    MDA_CHECK_LAUNCH();
    MDA_CUDA(o.finish());
    return MDAGPU_OK;
}

extern "C" int mdagpu_contact_fraction(const double *r, uint64_t nframes, uint64_t npairs, const double *r0,
                                       int method, double radius, double beta, double lambda, double *out,
                                       int space, mdagpu_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(r && out, "mdagpu_contact_fraction: NULL argument");
    MDA_REQUIRE(method >= 0 && method <= 2, "mdagpu_contact_fraction: method must be 0 (hard), 1 (radius) or 2 (soft)");
    MDA_REQUIRE(method == 1 || r0 != nullptr, "mdagpu_contact_fraction: r0 is NULL");
    MDA_REQUIRE(nframes <= 0x7fffffffull, "mdagpu_contact_fraction: too many frames in one call");
    if (nframes == 0) return MDAGPU_OK;
    StagedIn d, d0;
    StagedOut o;
    MDA_CUDA(d.init(r, nframes * npairs * sizeof(double), space, st));
    MDA_CUDA(d0.init(r0, npairs * sizeof(double), space, st));
    MDA_CUDA(o.init(out, nframes * sizeof(double), space, st));
    MDA_LAUNCH(contact_fraction_kernel, (unsigned)nframes, RD_THREADS, 0, st, d.as<double>(), npairs,
               r0 ? d0.as<double>() : nullptr, method, radius, beta, lambda, o.as<double>());
    MDA_CHECK_LAUNCH();
    MDA_CUDA(o.finish());
    return MDAGPU_OK;
}
