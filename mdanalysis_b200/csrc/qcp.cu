// qcp.cu -- batched QCP RMSD / optimal superposition (K8).
// Reference (package/MDAnalysis/):
//   lib/qcprot.pyx:179-277   InnerProduct              -> the A / G accumulation below
//   lib/qcprot.pyx:326-543   FastCalcRMSDAndRotation   -> geom.cuh qcp_solve()
//   analysis/rms.py:685-727  RMSD._single_frame        -> convention 0
//   analysis/align.py:284-366, 843-854  rotation_matrix / _fit_to / AlignTraj._single_frame
//                                                     -> convention 1
//   core/groups.py:1083-1106 center(), :1478-1529 translate()/rotate()
//
// One CTA per frame.  The selected atoms of the frame are read from HBM once,
// parked in shared memory as float32 (10 000 atoms = 120 kB), and used twice:
// pass 1 accumulates the (weighted) centre in float64, pass 2 the 3x3 inner
// product and G1 of the centred float64 coordinates, exactly the quantities the
// reference forms (it also centres in float64 before calling QCP).  A warp
// shuffle + shared-memory tree reduces the 10 sums; one thread then runs the
// Newton-Raphson quartic solve and the adjugate/quaternion step.  With
// `aligned`, all threads apply the rigid transform to every atom of the frame
// (the frame is still L2-resident) with the reference's three float32 roundings.
#include "common.cuh"

namespace mdagpu {

constexpr int QCP_THREADS = 512;

struct RefInfo {       // produced once per call by qcp_ref_kernel
    double com[3];
    double g2;         // sum_i w_i |ref_i - com|^2
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of K doubles per thread; result valid in every thread
template <int K, int THREADS>
__device__ __forceinline__ void block_sum(double (&v)[K], double *scratch /* [K * THREADS/32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = THREADS / 32;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = warp_sum(v[k]);
        if (lane == 0) scratch[k * NW + warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += scratch[k * NW + w];
        v[k] = s;
    }
    __syncthreads();
}

// centre and G2 of the reference selection
__global__ void __launch_bounds__(QCP_THREADS)
qcp_ref_kernel(const float *__restrict__ ref, const double *__restrict__ w, uint64_t nsel,
               RefInfo *__restrict__ out)
{
    __shared__ double scratch[4 * (QCP_THREADS / 32)];
    double s[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (uint64_t i = threadIdx.x; i < nsel; i += QCP_THREADS) {
        const double wi = w ? w[i] : 1.0;
        s[0] = fma(wi, (double)ref[3 * i], s[0]);
        s[1] = fma(wi, (double)ref[3 * i + 1], s[1]);
        s[2] = fma(wi, (double)ref[3 * i + 2], s[2]);
        s[3] += wi;
    }
    block_sum<4, QCP_THREADS>(s, scratch);
    const double cx = s[0] / s[3], cy = s[1] / s[3], cz = s[2] / s[3];
    double g[1] = { 0.0 };
    for (uint64_t i = threadIdx.x; i < nsel; i += QCP_THREADS) {
        const double wi = w ? w[i] : 1.0;
        const double x = (double)ref[3 * i] - cx, y = (double)ref[3 * i + 1] - cy, z = (double)ref[3 * i + 2] - cz;
        g[0] = fma(wi, x * x + y * y + z * z, g[0]);
    }
    block_sum<1, QCP_THREADS>(g, scratch);
    if (threadIdx.x == 0) {
        out->com[0] = cx; out->com[1] = cy; out->com[2] = cz;
        out->g2 = g[0];
    }
}

template <bool WEIGHTED, bool STAGE>
__global__ void __launch_bounds__(QCP_THREADS)
qcp_traj_kernel(const float *traj, uint64_t natoms, const int32_t *__restrict__ sel,
                uint64_t nsel, const float *__restrict__ ref, const double *__restrict__ w,
                const RefInfo *__restrict__ refinfo, double *__restrict__ rmsd_out,
                double *__restrict__ rot_out, float *aligned, int convention,
                int allclose_shortcut)
{
    extern __shared__ float s_xyz[];                       // [nsel*3] when STAGE
    __shared__ double scratch[11 * (QCP_THREADS / 32)];
    __shared__ double s_rot[9];
    __shared__ double s_com[3];
    const uint64_t f = blockIdx.x;
    const float *frame = traj + f * natoms * 3;

    // ---- pass 1: centre of the mobile selection (float64) ------------------------
    double c[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (uint64_t i = threadIdx.x; i < nsel; i += QCP_THREADS) {
        const uint64_t a = sel ? (uint64_t)sel[i] : i;
        const float x = frame[3 * a], y = frame[3 * a + 1], z = frame[3 * a + 2];
        if (STAGE) { s_xyz[3 * i] = x; s_xyz[3 * i + 1] = y; s_xyz[3 * i + 2] = z; }
        if (WEIGHTED) {
            const double wi = w[i];
            c[0] = fma(wi, (double)x, c[0]); c[1] = fma(wi, (double)y, c[1]); c[2] = fma(wi, (double)z, c[2]);
            c[3] += wi;
        } else {
            c[0] += (double)x; c[1] += (double)y; c[2] += (double)z;
        }
    }
    block_sum<4, QCP_THREADS>(c, scratch);
    const double wsum = WEIGHTED ? c[3] : (double)nsel;
    const double mx = c[0] / wsum, my = c[1] / wsum, mz = c[2] / wsum;
    const double rx = refinfo->com[0], ry = refinfo->com[1], rz = refinfo->com[2];

    // ---- pass 2: inner product of the centred coordinates ---------------------------
    // M[a][b] = sum w * mobile_a * ref_b ; G1 = sum w |mobile|^2
    double acc[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) acc[k] = 0.0;
    for (uint64_t i = threadIdx.x; i < nsel; i += QCP_THREADS) {
        float fx, fy, fz;
        if (STAGE) { fx = s_xyz[3 * i]; fy = s_xyz[3 * i + 1]; fz = s_xyz[3 * i + 2]; }
        else {
            const uint64_t a = sel ? (uint64_t)sel[i] : i;
            fx = frame[3 * a]; fy = frame[3 * a + 1]; fz = frame[3 * a + 2];
        }
        const double y0 = (double)fx - mx, y1 = (double)fy - my, y2 = (double)fz - mz;
        const double x0 = (double)ref[3 * i] - rx, x1 = (double)ref[3 * i + 1] - ry, x2 = (double)ref[3 * i + 2] - rz;
        double w0 = y0, w1 = y1, w2 = y2;
        if (WEIGHTED) { const double wi = w[i]; w0 *= wi; w1 *= wi; w2 *= wi; }
        acc[0] = fma(w0, x0, acc[0]); acc[1] = fma(w0, x1, acc[1]); acc[2] = fma(w0, x2, acc[2]);
        acc[3] = fma(w1, x0, acc[3]); acc[4] = fma(w1, x1, acc[4]); acc[5] = fma(w1, x2, acc[5]);
        acc[6] = fma(w2, x0, acc[6]); acc[7] = fma(w2, x1, acc[7]); acc[8] = fma(w2, x2, acc[8]);
        acc[9] = fma(w0, y0, fma(w1, y1, fma(w2, y2, acc[9])));
        if (allclose_shortcut) {
            // np.allclose(mobile, ref): |a - b| <= 1e-8 + 1e-5 |b|   (align.py:290)
            const bool ok = fabs(y0 - x0) <= 1e-8 + 1e-5 * fabs(x0) &&
                            fabs(y1 - x1) <= 1e-8 + 1e-5 * fabs(x1) &&
                            fabs(y2 - x2) <= 1e-8 + 1e-5 * fabs(x2);
            if (!ok) acc[10] += 1.0;
        }
    }
    block_sum<11, QCP_THREADS>(acc, scratch);

    if (threadIdx.x == 0) {
        double A[9], R[9];
        double rms;
        const bool need_rot = rot_out != nullptr || aligned != nullptr;
        if (allclose_shortcut && acc[10] == 0.0) {
            rms = 0.0;
            for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
        } else {
            if (convention == 0) {
                for (int k = 0; k < 9; ++k) A[k] = acc[k];                 // coords1 = mobile (rms.py:699-701)
            } else {
                for (int a = 0; a < 3; ++a)                                  // coords1 = reference (align.py:306)
                    for (int b = 0; b < 3; ++b) A[3 * a + b] = acc[3 * b + a];
            }
            const double E0 = (acc[9] + refinfo->g2) * 0.5;
            rms = qcp_solve(need_rot ? R : nullptr, A, E0, (double)nsel);
        }
        rmsd_out[f] = rms;
        if (rot_out)
            for (int k = 0; k < 9; ++k) rot_out[f * 9 + k] = R[k];
        if (aligned) {
            for (int k = 0; k < 9; ++k) s_rot[k] = R[k];
            s_com[0] = mx; s_com[1] = my; s_com[2] = mz;
        }
    }
    if (aligned == nullptr) return;
    __syncthreads();

    // ---- pass 3: move every atom of the frame --------------------------------------
    // translate(-com), rotate, translate(+ref_com): each step computed in float64
    // and stored back as float32 (core/groups.py:1480, 1527; rms.py:705-711).
    double R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = s_rot[k];
    const double cx = s_com[0], cy = s_com[1], cz = s_com[2];
    float *dst = aligned + f * natoms * 3;
    for (uint64_t a = threadIdx.x; a < natoms; a += QCP_THREADS) {
        const float p0 = (float)((double)frame[3 * a] - cx);
        const float p1 = (float)((double)frame[3 * a + 1] - cy);
        const float p2 = (float)((double)frame[3 * a + 2] - cz);
        double q0, q1, q2;
        if (convention == 0) {        // x . R
            q0 = fma((double)p2, R[6], fma((double)p1, R[3], (double)p0 * R[0]));
            q1 = fma((double)p2, R[7], fma((double)p1, R[4], (double)p0 * R[1]));
            q2 = fma((double)p2, R[8], fma((double)p1, R[5], (double)p0 * R[2]));
        } else {                      // x . R^T
            q0 = fma((double)p2, R[2], fma((double)p1, R[1], (double)p0 * R[0]));
            q1 = fma((double)p2, R[5], fma((double)p1, R[4], (double)p0 * R[3]));
            q2 = fma((double)p2, R[8], fma((double)p1, R[7], (double)p0 * R[6]));
        }
        const float r0 = (float)q0, r1 = (float)q1, r2 = (float)q2;
        dst[3 * a] = (float)((double)r0 + rx);
        dst[3 * a + 1] = (float)((double)r1 + ry);
        dst[3 * a + 2] = (float)((double)r2 + rz);
    }
}

// single pair of already-centred float64 structures (CalcRMSDRotationalMatrix)
__global__ void __launch_bounds__(QCP_THREADS)
qcp_single_kernel(const double *__restrict__ ref, const double *__restrict__ conf, uint64_t n,
                  const double *__restrict__ w, double *__restrict__ rot, double *__restrict__ rmsd)
{
    __shared__ double scratch[11 * (QCP_THREADS / 32)];
    double acc[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) acc[k] = 0.0;
    for (uint64_t i = threadIdx.x; i < n; i += QCP_THREADS) {
        // InnerProduct(A, conf, ref, ...): coords1 = conf, coords2 = ref (qcprot.pyx:319)
        const double y0 = conf[3 * i], y1 = conf[3 * i + 1], y2 = conf[3 * i + 2];
        const double x0 = ref[3 * i], x1 = ref[3 * i + 1], x2 = ref[3 * i + 2];
        const double wi = w ? w[i] : 1.0;
        const double w0 = wi * y0, w1 = wi * y1, w2 = wi * y2;
        acc[0] = fma(w0, x0, acc[0]); acc[1] = fma(w0, x1, acc[1]); acc[2] = fma(w0, x2, acc[2]);
        acc[3] = fma(w1, x0, acc[3]); acc[4] = fma(w1, x1, acc[4]); acc[5] = fma(w1, x2, acc[5]);
        acc[6] = fma(w2, x0, acc[6]); acc[7] = fma(w2, x1, acc[7]); acc[8] = fma(w2, x2, acc[8]);
        acc[9] = fma(w0, y0, fma(w1, y1, fma(w2, y2, acc[9])));
        acc[10] = fma(wi, x0 * x0 + x1 * x1 + x2 * x2, acc[10]);
    }
    block_sum<11, QCP_THREADS>(acc, scratch);
    if (threadIdx.x == 0) {
        double R[9];
        const double E0 = (acc[9] + acc[10]) * 0.5;
        *rmsd = qcp_solve(rot ? R : nullptr, acc, E0, (double)n);
        if (rot)
            for (int k = 0; k < 9; ++k) rot[k] = R[k];
    }
}

}  // namespace mdagpu

using namespace mdagpu;

extern "C" int mdagpu_qcp(const double *ref, const double *conf, uint64_t n, double *rot,
                          const double *weights, double *rmsd, int space, mdagpu_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(ref && conf && rmsd, "mdagpu_qcp: NULL argument");
    MDA_REQUIRE(n > 0, "mdagpu_qcp: empty structures");
    StagedIn r, c, w;
    StagedOut ro;
    DevBuf val;
    MDA_CUDA(r.init(ref, n * 3 * sizeof(double), space, st));
    MDA_CUDA(c.init(conf, n * 3 * sizeof(double), space, st));
    MDA_CUDA(w.init(weights, n * sizeof(double), space, st));
    MDA_CUDA(ro.init(rot, 9 * sizeof(double), space, st));
    MDA_CUDA(val.alloc(sizeof(double), st));
    MDA_LAUNCH(qcp_single_kernel, 1, QCP_THREADS, 0, st, r.as<double>(), c.as<double>(), n,
               weights ? w.as<double>() : nullptr, rot ? ro.as<double>() : nullptr, val.as<double>());
    MDA_CHECK_LAUNCH();
    MDA_CUDA(cudaMemcpyAsync(rmsd, val.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (rot) MDA_CUDA(ro.finish());
    MDA_CUDA(cudaStreamSynchronize(st));
    return MDAGPU_OK;
}

extern "C" int mdagpu_qcp_trajectory(const float *traj, uint64_t nframes, uint64_t natoms,
                                     const int32_t *sel, uint64_t nsel, const float *ref,
                                     const double *weights, double *rmsd, double *rot,
                                     float *aligned, int convention, int allclose_shortcut,
                                     int space, mdagpu_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(traj && ref && rmsd, "mdagpu_qcp_trajectory: NULL argument");
    MDA_REQUIRE(nsel > 0 && nsel <= natoms, "mdagpu_qcp_trajectory: bad selection size %llu of %llu",
                (unsigned long long)nsel, (unsigned long long)natoms);
    MDA_REQUIRE(convention == 0 || convention == 1, "mdagpu_qcp_trajectory: convention must be 0 or 1");
    if (nframes == 0) return MDAGPU_OK;
    const bool inplace = aligned == traj;
    StagedIn t, s, r, w;
    StagedOut rm, ro, al;
    // in-place alignment of a host trajectory: stage once, copy back the same buffer
    if (aligned && inplace && space == MDAGPU_HOST) {
        MDA_CUDA(al.init(aligned, nframes * natoms * 3 * sizeof(float), space, st, true));
        t.dev = al.dev;
    } else {
        MDA_CUDA(t.init(traj, nframes * natoms * 3 * sizeof(float), space, st));
        if (aligned) MDA_CUDA(al.init(aligned, nframes * natoms * 3 * sizeof(float), space, st));
    }
    MDA_CUDA(s.init(sel, nsel * sizeof(int32_t), space, st));
    MDA_CUDA(r.init(ref, nsel * 3 * sizeof(float), space, st));
    MDA_CUDA(w.init(weights, nsel * sizeof(double), space, st));
    MDA_CUDA(rm.init(rmsd, nframes * sizeof(double), space, st));
    MDA_CUDA(ro.init(rot, nframes * 9 * sizeof(double), space, st));
    DevBuf info;
    MDA_CUDA(info.alloc(sizeof(RefInfo), st));
    const double *wd = weights ? w.as<double>() : nullptr;
    const int32_t *sd = sel ? s.as<int32_t>() : nullptr;
    MDA_LAUNCH(qcp_ref_kernel, 1, QCP_THREADS, 0, st, r.as<float>(), wd, nsel, info.as<RefInfo>());
    MDA_CHECK_LAUNCH();

    const size_t stage_bytes = nsel * 3 * sizeof(float);
    const bool stage = stage_bytes <= 200 * 1024;
    const size_t smem = stage ? stage_bytes : 0;
#define QCP_LAUNCH(W, S)                                                                         \
    do {                                                                                         \
        auto kern = qcp_traj_kernel<W, S>;                                                       \
        if (smem > 48 * 1024)                                                                    \
            MDA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        MDA_LAUNCH_AS("qcp_traj_kernel", kern, (unsigned)nframes, QCP_THREADS, smem, st, t.as<float>(), natoms, sd, nsel, \
                   r.as<float>(), wd, info.as<RefInfo>(), rm.as<double>(),                       \
                   rot ? ro.as<double>() : nullptr, aligned ? al.as<float>() : nullptr,          \
                   convention, allclose_shortcut);                                               \
    } while (0)
    MDA_REQUIRE(nframes <= 0x7fffffffull, "mdagpu_qcp_trajectory: too many frames in one call");
    if (weights) { if (stage) QCP_LAUNCH(true, true); else QCP_LAUNCH(true, false); }
    else { if (stage) QCP_LAUNCH(false, true); else QCP_LAUNCH(false, false); }
#undef QCP_LAUNCH
    MDA_CHECK_LAUNCH();
    MDA_CUDA(rm.finish());
    if (rot) MDA_CUDA(ro.finish());
    if (aligned) MDA_CUDA(al.finish());
    return MDAGPU_OK;
}
