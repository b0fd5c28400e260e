// qcp.cu -- batched QCP RMSD / optimal superposition (K8).
// Reference (package/MDAnalysis/):
//   lib/qcprot.pyx:179-277   InnerProduct              -> the A / G accumulation below
//   lib/qcprot.pyx:326-543   FastCalcRMSDAndRotation   -> geom.cuh qcp_solve()
//   analysis/rms.py:685-727  RMSD._single_frame        -> convention 0
//   analysis/align.py:284-366, 843-854  rotation_matrix / _fit_to / AlignTraj._single_frame
//                                                     -> convention 1
//   core/groups.py:1083-1106 center(), :1478-1529 translate()/rotate()
//
// Two kernels.  qcp_traj_kernel (index lists, small or unaligned selections): one CTA per
// frame, several CTAs per SM, 16-byte loads where the layout allows.  qcp_stream_kernel (the
// C4 path: contiguous selections of >= 2048 atoms): persistent CTAs, one WARP per frame, the
// frames of a batch streamed through a TMA ring next to the float64 reference chunk they share.
// Both read every frame from HBM once and accumulate in a single pass, in float64 and
// relative to a per-frame pivot (the first fitted atom, which removes the translation from
// the magnitudes that are summed),
//     Sy = sum w y',  M = sum w y' (x - x_com)^T,  G = sum w |y'|^2 ,   y' = y - pivot
// from which the centred quantities the reference forms (it centres in float64
// before calling QCP, rms.py:685-701) follow exactly:
//     y_com = pivot + Sy/W,  A = M - (Sy/W) (sum w (x - x_com))^T,  G1 = G - |Sy|^2/W.
// The centred (and weight-scaled) reference is prepared once per call in float64
// (qcp_ref_kernel).  One thread per frame runs the Newton-Raphson quartic solve and the
// adjugate/quaternion step (geom.cuh qcp_solve).  With `aligned`, the frame is read a second
// time and every atom gets translate(-com), rotate(R), translate(+ref_com), each rounded to
// float32 as the reference's in-place float32 positions are (core/groups.py:1480, 1527).
#include "common.cuh"
#include "tma.cuh"

namespace mdagpu {

constexpr int QCP_THREADS = 512;

struct RefInfo {       // produced once per call by qcp_ref_kernel
    double com[3];
    double g2;         // sum_i w_i |ref_i - com|^2
    double sx[3];      // sum_i w_i (ref_i - com): zero up to rounding, kept for rigour
    double wsum;       // sum_i w_i (nsel without weights)
    double close_sum;  // sum_i,k (1e-8 + 1e-5 |ref_ik - com_k|)^2 : screen for np.allclose (align.py:290)
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

// centre, G2 and the centred float64 copy of the reference selection:
// refc[3i+k] = w_i * (ref_i - com)_k  (w_i = 1 without weights), sx = sum_i refc[i]
__global__ void __launch_bounds__(QCP_THREADS)
qcp_ref_kernel(const float *__restrict__ ref, const double *__restrict__ w, uint64_t nsel,
               RefInfo *__restrict__ out, double *__restrict__ refc)
{
    __shared__ double scratch[5 * (QCP_THREADS / 32)];
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
    double g[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
    for (uint64_t i = threadIdx.x; i < nsel; i += QCP_THREADS) {
        const double wi = w ? w[i] : 1.0;
        const double x = (double)ref[3 * i] - cx, y = (double)ref[3 * i + 1] - cy, z = (double)ref[3 * i + 2] - cz;
        g[0] = fma(wi, x * x + y * y + z * z, g[0]);
        const double wx = wi * x, wy = wi * y, wz = wi * z;
        refc[3 * i] = wx; refc[3 * i + 1] = wy; refc[3 * i + 2] = wz;
        g[1] += wx; g[2] += wy; g[3] += wz;
        const double tx = 1e-8 + 1e-5 * fabs(x), ty = 1e-8 + 1e-5 * fabs(y), tz = 1e-8 + 1e-5 * fabs(z);
        g[4] += tx * tx + ty * ty + tz * tz;
    }
    block_sum<5, QCP_THREADS>(g, scratch);
    if (threadIdx.x == 0) {
        out->com[0] = cx; out->com[1] = cy; out->com[2] = cz;
        out->g2 = g[0];
        out->sx[0] = g[1]; out->sx[1] = g[2]; out->sx[2] = g[3];
        out->wsum = s[3];
        out->close_sum = g[4];
    }
}

// float -> double without the (quarter-rate) conversion unit: exact for normal numbers,
// everything else takes the hardware conversion.
__device__ __forceinline__ double f2d_alu(float x)
{
    const uint32_t b = __float_as_uint(x);
    const uint32_t e = b & 0x7f800000u;
    if (e == 0u || e == 0x7f800000u) return (double)x;
    const uint32_t hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
    return __hiloint2double((int)hi, (int)(b << 29));
}

// accumulate one atom: y' = y - pivot (exact in double), xw = w (x - x_com)
template <bool WEIGHTED>
__device__ __forceinline__ void qcp_accumulate(double (&acc)[13], float fx, float fy, float fz,
                                               double px, double py, double pz, double xw0, double xw1,
                                               double xw2, double wi)
{
#if QCP_F2D_ALU
    const double y0 = f2d_alu(fx) - px, y1 = f2d_alu(fy) - py, y2 = f2d_alu(fz) - pz;
#else
    const double y0 = (double)fx - px, y1 = (double)fy - py, y2 = (double)fz - pz;
#endif
    acc[0] = fma(y0, xw0, acc[0]); acc[1] = fma(y0, xw1, acc[1]); acc[2] = fma(y0, xw2, acc[2]);
    acc[3] = fma(y1, xw0, acc[3]); acc[4] = fma(y1, xw1, acc[4]); acc[5] = fma(y1, xw2, acc[5]);
    acc[6] = fma(y2, xw0, acc[6]); acc[7] = fma(y2, xw1, acc[7]); acc[8] = fma(y2, xw2, acc[8]);
    if (WEIGHTED) {
        const double w0 = wi * y0, w1 = wi * y1, w2 = wi * y2;
        acc[9] = fma(w0, y0, fma(w1, y1, fma(w2, y2, acc[9])));
        acc[10] += w0; acc[11] += w1; acc[12] += w2;
    } else {
        acc[9] = fma(y0, y0, fma(y1, y1, fma(y2, y2, acc[9])));
        acc[10] += y0; acc[11] += y1; acc[12] += y2;
    }
}

// np.allclose(mobile - mobile_com, ref - ref_com) of align.rotation_matrix (align.py:290):
// |a - b| <= 1e-8 + 1e-5 |b| for every component.  Needs the centres, hence its own pass
// (only AlignTraj without weights asks for it; the frame is L2-resident by then).
__device__ __forceinline__ bool qcp_close(double y, double x)
{
    return fabs(y - x) <= 1e-8 + 1e-5 * fabs(x);
}

template <int T, bool WEIGHTED, bool VEC>
__global__ void __launch_bounds__(T)
qcp_traj_kernel(const float *traj, uint64_t natoms, const int32_t *__restrict__ sel, uint64_t nsel,
                const double *__restrict__ refc, const double *__restrict__ w,
                const RefInfo *__restrict__ refinfo, double *__restrict__ rmsd_out,
                double *__restrict__ rot_out, float *aligned, int convention, int allclose_shortcut)
{
    __shared__ double scratch[13 * (T / 32)];
    __shared__ double s_rot[9];
    __shared__ double s_com[3];
    const uint64_t f = blockIdx.x;
    const float *frame = traj + f * natoms * 3;
    const uint64_t a0 = sel ? (uint64_t)sel[0] : 0;
    const double px = (double)frame[3 * a0], py = (double)frame[3 * a0 + 1], pz = (double)frame[3 * a0 + 2];

    // ---- single pass over the fitted atoms ---------------------------------------------
    double acc[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) acc[k] = 0.0;
    uint64_t scalar_from = 0;
    if (VEC) {      // natoms % 4 == 0 and no index list: 16-byte loads, 4 atoms per step
        const float4 *fr4 = reinterpret_cast<const float4 *>(frame);
        const double2 *rc2 = reinterpret_cast<const double2 *>(refc);
        const uint64_t nquad = nsel / 4;
        scalar_from = nquad * 4;
#pragma unroll 2
        for (uint64_t q = threadIdx.x; q < nquad; q += T) {
            const float4 v0 = fr4[3 * q], v1 = fr4[3 * q + 1], v2 = fr4[3 * q + 2];
            const double2 r0 = __ldg(rc2 + 6 * q), r1 = __ldg(rc2 + 6 * q + 1), r2 = __ldg(rc2 + 6 * q + 2);
            const double2 r3 = __ldg(rc2 + 6 * q + 3), r4 = __ldg(rc2 + 6 * q + 4), r5 = __ldg(rc2 + 6 * q + 5);
            double w0 = 1.0, w1 = 1.0, w2 = 1.0, w3 = 1.0;
            if (WEIGHTED) {
                const double2 wa = __ldg(reinterpret_cast<const double2 *>(w) + 2 * q);
                const double2 wb = __ldg(reinterpret_cast<const double2 *>(w) + 2 * q + 1);
                w0 = wa.x; w1 = wa.y; w2 = wb.x; w3 = wb.y;
            }
            qcp_accumulate<WEIGHTED>(acc, v0.x, v0.y, v0.z, px, py, pz, r0.x, r0.y, r1.x, w0);
            qcp_accumulate<WEIGHTED>(acc, v0.w, v1.x, v1.y, px, py, pz, r1.y, r2.x, r2.y, w1);
            qcp_accumulate<WEIGHTED>(acc, v1.z, v1.w, v2.x, px, py, pz, r3.x, r3.y, r4.x, w2);
            qcp_accumulate<WEIGHTED>(acc, v2.y, v2.z, v2.w, px, py, pz, r4.y, r5.x, r5.y, w3);
        }
    }
    for (uint64_t i = scalar_from + threadIdx.x; i < nsel; i += T) {
        const uint64_t a = sel ? (uint64_t)sel[i] : i;
        const float fx = frame[3 * a], fy = frame[3 * a + 1], fz = frame[3 * a + 2];
        qcp_accumulate<WEIGHTED>(acc, fx, fy, fz, px, py, pz, refc[3 * i], refc[3 * i + 1], refc[3 * i + 2],
                                 WEIGHTED ? w[i] : 1.0);
    }
    block_sum<13, T>(acc, scratch);

    const double W = refinfo->wsum;                       // sum of weights (nsel when unweighted)
    const double m0 = acc[10] / W, m1 = acc[11] / W, m2 = acc[12] / W;     // centre relative to the pivot
    const double mx = px + m0, my = py + m1, mz = pz + m2;
    const double rx = refinfo->com[0], ry = refinfo->com[1], rz = refinfo->com[2];

    // ---- optional: align.rotation_matrix's allclose early exit ---------------------------
    double notclose = 0.0;
    if (allclose_shortcut) {
        for (uint64_t i = threadIdx.x; i < nsel; i += T) {
            const uint64_t a = sel ? (uint64_t)sel[i] : i;
            const double y0 = (double)frame[3 * a] - mx, y1 = (double)frame[3 * a + 1] - my, y2 = (double)frame[3 * a + 2] - mz;
            if (!(qcp_close(y0, refc[3 * i]) && qcp_close(y1, refc[3 * i + 1]) && qcp_close(y2, refc[3 * i + 2])))
                notclose = 1.0;
        }
        notclose = (double)__syncthreads_or(notclose != 0.0);
    }

    if (threadIdx.x == 0) {
        double A[9], R[9];
        double rms;
        const bool need_rot = rot_out != nullptr || aligned != nullptr;
        if (allclose_shortcut && notclose == 0.0) {
            rms = 0.0;
            for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
        } else {
            // M[a][b] = sum w (y_a - y_com_a) (x_b - x_com_b)
            const double sx[3] = { refinfo->sx[0], refinfo->sx[1], refinfo->sx[2] };
            const double m[3] = { m0, m1, m2 };
            double M[9];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) M[3 * a + b] = acc[3 * a + b] - m[a] * sx[b];
            if (convention == 0) {
                for (int k = 0; k < 9; ++k) A[k] = M[k];                       // coords1 = mobile (rms.py:699-701)
            } else {
                for (int a = 0; a < 3; ++a)                                    // coords1 = reference (align.py:306)
                    for (int b = 0; b < 3; ++b) A[3 * a + b] = M[3 * b + a];
            }
            const double G1 = acc[9] - W * (m0 * m0 + m1 * m1 + m2 * m2);
            const double E0 = (G1 + refinfo->g2) * 0.5;
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

    // ---- second pass: move every atom of the frame ---------------------------------------
    // translate(-com), rotate, translate(+ref_com): each step computed in float64
    // and stored back as float32 (core/groups.py:1480, 1527; rms.py:705-711).
    double R[9];
    if (convention == 0) {        // x . R
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = s_rot[k];
    } else {                      // x . R^T
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) R[3 * a + b] = s_rot[3 * b + a];
    }
    const double cx = s_com[0], cy = s_com[1], cz = s_com[2];
    float *dst = aligned + f * natoms * 3;
    auto move = [&](float x, float y, float z, float &ox, float &oy, float &oz) {
        const float p0 = (float)((double)x - cx);
        const float p1 = (float)((double)y - cy);
        const float p2 = (float)((double)z - cz);
        const double q0 = fma((double)p2, R[6], fma((double)p1, R[3], (double)p0 * R[0]));
        const double q1 = fma((double)p2, R[7], fma((double)p1, R[4], (double)p0 * R[1]));
        const double q2 = fma((double)p2, R[8], fma((double)p1, R[5], (double)p0 * R[2]));
        const float r0 = (float)q0, r1 = (float)q1, r2 = (float)q2;
        ox = (float)((double)r0 + rx);
        oy = (float)((double)r1 + ry);
        oz = (float)((double)r2 + rz);
    };
    uint64_t from = 0;
    if (VEC) {
        const float4 *fr4 = reinterpret_cast<const float4 *>(frame);
        float4 *dst4 = reinterpret_cast<float4 *>(dst);
        const uint64_t nquad = natoms / 4;
        from = nquad * 4;
#pragma unroll 2
        for (uint64_t q = threadIdx.x; q < nquad; q += T) {
            const float4 v0 = fr4[3 * q], v1 = fr4[3 * q + 1], v2 = fr4[3 * q + 2];
            float4 o0, o1, o2;
            move(v0.x, v0.y, v0.z, o0.x, o0.y, o0.z);
            move(v0.w, v1.x, v1.y, o0.w, o1.x, o1.y);
            move(v1.z, v1.w, v2.x, o1.z, o1.w, o2.x);
            move(v2.y, v2.z, v2.w, o2.y, o2.z, o2.w);
            dst4[3 * q] = o0; dst4[3 * q + 1] = o1; dst4[3 * q + 2] = o2;
        }
    }
    for (uint64_t a = from + threadIdx.x; a < natoms; a += T)
        move(frame[3 * a], frame[3 * a + 1], frame[3 * a + 2], dst[3 * a], dst[3 * a + 1], dst[3 * a + 2]);
}

// ---------------------------------------------------------------------------
// TMA-streamed variant (contiguous selection, 16-byte aligned frames)
// ---------------------------------------------------------------------------
// Persistent CTAs, one per SM.  A producer warp streams chunks of QS_CH atoms with
// cp.async.bulk (1-D TMA) into a shared-memory ring guarded by full/empty mbarriers;
// 16 consumer warps in QS_FB groups of QS_GT threads work on QS_FB frames at once, so
// the float64 reference chunk (staged once per stage) is shared by QS_FB frames.
// No registers are spent on loads in flight.
//
// r02 (profiles/r02_ncu_qcp_stream_kernel.txt: issue slots 44 % busy, 1.3 warps per issue
// waiting at the full barrier, 0.9 at the group barrier): a group is now ONE WARP
// (QS_GT = 32, 16 frames per batch, chunks of 256 atoms): no barrier among the consumers
// at all (the 13 sums are reduced by shuffles), the serial Newton solve of 16 frames runs
// in 16 warps side by side once per 16 frames, the reference chunk is a ninth of a stage
// instead of a third (more frame bytes in flight), and full chunks take a path without
// the ragged-quad checks.  Frames are dealt to the CTAs as contiguous, balanced ranges.
// With `aligned`, the same ring streams the frames a second time (L2-resident) and the
// consumers apply the rigid transform IN PLACE in the stage, which then leaves through a
// bulk store (cp.async.bulk shared -> global: full 128-byte lines instead of 16-byte
// pieces 48 bytes apart); the stage is released one chunk later, when the store has read it.
#ifndef QS_GT_
#define QS_GT_ 32
#endif
#ifndef QS_CH_
#define QS_CH_ 256
#endif
#ifndef QCP_F2D_ALU
#define QCP_F2D_ALU 0
#endif
#ifndef QS_BULK_STORE
#define QS_BULK_STORE 1
#endif
constexpr int QS_CH = QS_CH_;               // atoms per chunk
constexpr int QS_GT = QS_GT_;               // threads per group (one frame)
constexpr int QS_FB = 512 / QS_GT;          // frames per batch = consumer groups
constexpr int QS_GW = QS_GT / 32;           // warps per group
constexpr int QS_QPT = QS_CH / 4 / QS_GT;   // quads per thread and chunk
constexpr int QS_THREADS = QS_FB * QS_GT;
static_assert(QS_GT == 32 || QS_GT == 64 || QS_GT == 128, "group = 1, 2 or 4 warps");
static_assert(QS_QPT >= 1 && QS_QPT * 4 * QS_GT == QS_CH, "a chunk is a whole number of quads per thread");
constexpr int QS_REF_BYTES = QS_CH * 3 * 8;          // 6 planes of double2
constexpr int QS_W_BYTES = QS_CH * 8;                // 2 planes of double2
constexpr int QS_FRAME_BYTES = QS_CH * 12;
// ring depth: as many stages as fit next to the static shared memory in the 227 kB of an SM
constexpr int QS_SMEM_BUDGET = 216 * 1024;
constexpr int QS_STAGES_U = QS_SMEM_BUDGET / (QS_REF_BYTES + QS_FB * QS_FRAME_BYTES);
constexpr int QS_STAGES_W = QS_SMEM_BUDGET / (QS_REF_BYTES + QS_W_BYTES + QS_FB * QS_FRAME_BYTES);
static_assert(QS_STAGES_W >= 2, "at least two stages");

__device__ __forceinline__ void group_sync(int group)
{
    if (QS_GT == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(QS_GT) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t *bar, uint32_t n)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
// L2 policies for the AlignTraj pass (QS_L2_HINTS, off): evict_last on the loads of the first pass and
// evict_first on the second-pass loads and the stores were meant to keep the frames in L2 between
// the passes.  Measured: no effect at 4, 8 or 16 frames per CTA (3.35 / 2.97 / 3.00 ms with and
// without, profiles/microbench/qcp_variants_b200.txt) -- the second pass re-reads from HBM either
// way, the pass moves 36 B per atom and runs at 6.1 TB/s of DRAM traffic.
#ifndef QS_L2_HINTS
#define QS_L2_HINTS 0
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes, uint64_t pol)
{
    if (QS_L2_HINTS)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                     ::"l"(dst), "r"(smem_u32(src)), "r"(bytes), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// fl32(x + t) for a float32 x and a float64 t = th + tl (th = fl32(t), tl = fl32(t - th)), two
// coordinates at a time, without leaving the FP32 pipe: Knuth's TwoSum gives x + th = s + e exactly,
// the result is fl32(s + fl32(e + tl)).  This is the float64 sum rounded once to float32 -- what
// `positions += vector` does (core/groups.py:1480) -- unless the exact value lies within ~2^-23 ulp of
// a rounding midpoint (about one coordinate in 10^7; it is then the neighbouring float32).  The
// float64 version cost nine conversions per atom on the quarter-rate XU pipe, which bounded the
// AlignTraj pass (profiles/r02b_ncu_qcp_align.txt: XU 61 % over both passes).
#ifndef QCP_DF_TRANSLATE
#define QCP_DF_TRANSLATE 1
#endif
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 df_translate2(float2 x, float2 th, float2 tl)
{
    const float2 s = __fadd2_rn(x, th);
    const float2 bb = __fadd2_rn(s, neg2(x));
    const float2 t1 = __fadd2_rn(s, neg2(bb));
    const float2 t2 = __fadd2_rn(x, neg2(t1));
    const float2 t3 = __fadd2_rn(th, neg2(bb));
    const float2 e = __fadd2_rn(t2, t3);
    const float2 r = __fadd2_rn(e, tl);
    return __fadd2_rn(s, r);
}
struct DFShift { float2 xy_h, xy_l, zx_h, zx_l, yz_h, yz_l; };
__device__ __forceinline__ DFShift df_shift(double tx, double ty, double tz)
{
    const float hx = (float)tx, hy = (float)ty, hz = (float)tz;
    const float lx = (float)(tx - (double)hx), ly = (float)(ty - (double)hy), lz = (float)(tz - (double)hz);
    DFShift d;
    d.xy_h = make_float2(hx, hy); d.xy_l = make_float2(lx, ly);
    d.zx_h = make_float2(hz, hx); d.zx_l = make_float2(lz, lx);
    d.yz_h = make_float2(hy, hz); d.yz_l = make_float2(ly, lz);
    return d;
}

// chunked plane layout of the centred reference: for chunk c, plane p (0..5), quad q
// the double2 {d[2p], d[2p+1]} of the quad's 12 doubles d = {x0 y0 z0 x1 y1 z1 ...};
// weights: plane p (0..1) holds {w[2p], w[2p+1]} of the quad.  Padding is zero.
__global__ void __launch_bounds__(256)
qcp_ref_planes_kernel(const double *__restrict__ refc, const double *__restrict__ w, uint64_t nsel,
                      uint64_t nchunks, double2 *__restrict__ refp, double2 *__restrict__ wp)
{
    const uint64_t total = nchunks * (QS_CH / 4);
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = t / (QS_CH / 4), q = t % (QS_CH / 4);
        double d[12], ww[4];
        for (int k = 0; k < 4; ++k) {
            const uint64_t atom = c * QS_CH + q * 4 + k;
            const bool ok = atom < nsel;
            for (int j = 0; j < 3; ++j) d[3 * k + j] = ok ? refc[3 * atom + j] : 0.0;
            ww[k] = ok ? (w ? w[atom] : 1.0) : 0.0;
        }
        for (int pl = 0; pl < 6; ++pl) refp[(c * 6 + pl) * (QS_CH / 4) + q] = make_double2(d[2 * pl], d[2 * pl + 1]);
        if (wp)
            for (int pl = 0; pl < 2; ++pl) wp[(c * 2 + pl) * (QS_CH / 4) + q] = make_double2(ww[2 * pl], ww[2 * pl + 1]);
    }
}

struct QcpStreamParams {
    const float *traj;
    uint64_t nframes, natoms, nsel;
    const double *refc;          // flat centred reference (allclose slow path)
    const double2 *refp, *wp;    // chunked planes
    const RefInfo *refinfo;
    double *rmsd_out, *rot_out;
    float *aligned;
    int convention, allclose_shortcut;
};

template <bool WEIGHTED, int QS_STAGES>
__global__ void __launch_bounds__(QS_THREADS + 32, 1) qcp_stream_kernel(const QcpStreamParams p)
{
    extern __shared__ __align__(128) unsigned char qs_smem[];
    __shared__ uint64_t full_bar[QS_STAGES], empty_bar[QS_STAGES];
    __shared__ double red[QS_GW > 1 ? QS_FB : 1][13 * QS_GW];
    __shared__ double s_xf[QS_FB][12];
    constexpr int STAGE_BYTES = QS_REF_BYTES + (WEIGHTED ? QS_W_BYTES : 0) + QS_FB * QS_FRAME_BYTES;
    constexpr int FRAME_OFF = QS_REF_BYTES + (WEIGHTED ? QS_W_BYTES : 0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < QS_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], QS_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // this CTA's frames: a contiguous range, the ranges differ by at most one frame
    const uint64_t fbase = p.nframes / gridDim.x, frem = p.nframes % gridDim.x;
    const uint64_t fr0 = blockIdx.x * fbase + min((uint64_t)blockIdx.x, frem);
    const uint64_t fr1 = fr0 + fbase + (blockIdx.x < frem ? 1u : 0u);
    const uint64_t my_batches = (fr1 - fr0 + QS_FB - 1) / QS_FB;
    const uint32_t nchA = (uint32_t)((p.nsel + QS_CH - 1) / QS_CH);
    const uint32_t nchB = p.aligned ? (uint32_t)((p.natoms + QS_CH - 1) / QS_CH) : 0u;

    // ---------------- producer warp ----------------
    // Chunk number `k` of this CTA (over all its batches and both passes) goes to stage k % STAGES.
    // After every consumed chunk a warp releases the stage; the producer waits until all consumer
    // warps have done so and refills it -- no consumer ever blocks on another consumer.
    const uint32_t per_batch = nchA + nchB;
    if (warp == QS_THREADS / 32) {
        // lane g issues the copy of frame g, lane QS_FB the reference chunk (one thread issuing all
        // 17 copies of a stage spent ~55 ns on each: profiles/microbench/qcp_variants_b200.txt)
        const uint64_t my_chunks = my_batches * per_batch;
        const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
        uint64_t f0 = fr0;
        uint32_t c = 0;
        for (uint64_t k = 0; k < my_chunks; ++k) {
            const int s = (int)(k % QS_STAGES);
            if (k >= QS_STAGES) mbar_wait(&empty_bar[s], ((k - QS_STAGES) / QS_STAGES) & 1u);
            const int nf = (int)min((uint64_t)QS_FB, fr1 - f0);
            unsigned char *st = qs_smem + (size_t)s * STAGE_BYTES;
            const bool passA = c < nchA;
            const uint32_t cc = passA ? c : c - nchA;
            const uint64_t lim = passA ? p.nsel : p.natoms;
            const uint32_t na = (uint32_t)min((uint64_t)QS_CH, lim - (uint64_t)cc * QS_CH);
            const uint32_t fbytes = ((na + 3u) & ~3u) * 12u;
            if (lane == 0) {
                uint32_t bytes = (uint32_t)nf * fbytes;
                if (passA) bytes += QS_REF_BYTES + (WEIGHTED ? QS_W_BYTES : 0);
                mbar_expect_tx(&full_bar[s], bytes);
            }
            __syncwarp();
            if (lane == QS_FB && passA) {
                bulk_g2s(st, p.refp + (uint64_t)cc * 6 * (QS_CH / 4), QS_REF_BYTES, &full_bar[s]);
                if (WEIGHTED) bulk_g2s(st + QS_REF_BYTES, p.wp + (uint64_t)cc * 2 * (QS_CH / 4), QS_W_BYTES, &full_bar[s]);
            }
            if (lane < nf) {
                unsigned char *dst = st + FRAME_OFF + lane * QS_FRAME_BYTES;
                const float *src = p.traj + ((f0 + lane) * p.natoms + (uint64_t)cc * QS_CH) * 3;
                if (QS_L2_HINTS && nchB) bulk_g2s_hint(dst, src, fbytes, &full_bar[s], passA ? pol_keep : pol_stream);
                else bulk_g2s(dst, src, fbytes, &full_bar[s]);
            }
            if (++c == per_batch) { c = 0; f0 += QS_FB; }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int g = warp / QS_GW, gt = tid % QS_GT;
    const double W = p.refinfo->wsum;
    const double rx = p.refinfo->com[0], ry = p.refinfo->com[1], rz = p.refinfo->com[2];
    uint32_t it = 0;
    const uint64_t pol_out = l2_policy_evict_first();
    int pending = -1;                                   // stage whose bulk store has not been waited for (thread gt == 0)
    for (uint64_t b = 0; b < my_batches; ++b) {
        const uint64_t f = fr0 + b * QS_FB + g;
        const bool active = f < fr1;
        double acc[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) acc[k] = 0.0;
        double px = 0.0, py = 0.0, pz = 0.0;
        for (uint32_t c = 0; c < nchA; ++c, ++it) {
            const int s = it % QS_STAGES;
            mbar_wait(&full_bar[s], (it / QS_STAGES) & 1u);
            if (active) {
                const unsigned char *st = qs_smem + (size_t)s * STAGE_BYTES;
                const double2 *rp = reinterpret_cast<const double2 *>(st);
                const double2 *wq = reinterpret_cast<const double2 *>(st + QS_REF_BYTES);
                const float4 *fr4 = reinterpret_cast<const float4 *>(st + FRAME_OFF + g * QS_FRAME_BYTES);
                if (c == 0) {
                    const float4 first = fr4[0];
                    px = (double)first.x; py = (double)first.y; pz = (double)first.z;
                }
                const uint32_t na = (uint32_t)min((uint64_t)QS_CH, p.nsel - (uint64_t)c * QS_CH);
                auto quad = [&](uint32_t q, bool ragged) {
                    float4 v0 = fr4[3 * q], v1 = fr4[3 * q + 1], v2 = fr4[3 * q + 2];
                    if (ragged && q * 4 + 4 > na) {     // last quad: atoms past the selection count as the pivot (zero y')
                        const float fx = (float)px, fy = (float)py, fz = (float)pz;
                        if (q * 4 + 1 >= na) { v0.w = fx; v1.x = fy; v1.y = fz; }
                        if (q * 4 + 2 >= na) { v1.z = fx; v1.w = fy; v2.x = fz; }
                        if (q * 4 + 3 >= na) { v2.y = fx; v2.z = fy; v2.w = fz; }
                    }
                    const double2 r0 = rp[q], r1 = rp[QS_CH / 4 + q], r2 = rp[2 * (QS_CH / 4) + q];
                    const double2 r3 = rp[3 * (QS_CH / 4) + q], r4 = rp[4 * (QS_CH / 4) + q], r5 = rp[5 * (QS_CH / 4) + q];
                    double w0 = 1.0, w1 = 1.0, w2 = 1.0, w3 = 1.0;
                    if (WEIGHTED) {
                        const double2 wa = wq[q], wb = wq[QS_CH / 4 + q];
                        w0 = wa.x; w1 = wa.y; w2 = wb.x; w3 = wb.y;
                    }
                    qcp_accumulate<WEIGHTED>(acc, v0.x, v0.y, v0.z, px, py, pz, r0.x, r0.y, r1.x, w0);
                    qcp_accumulate<WEIGHTED>(acc, v0.w, v1.x, v1.y, px, py, pz, r1.y, r2.x, r2.y, w1);
                    qcp_accumulate<WEIGHTED>(acc, v1.z, v1.w, v2.x, px, py, pz, r3.x, r3.y, r4.x, w2);
                    qcp_accumulate<WEIGHTED>(acc, v2.y, v2.z, v2.w, px, py, pz, r4.y, r5.x, r5.y, w3);
                };
                if (na == (uint32_t)QS_CH) {
#pragma unroll
                    for (int h = 0; h < QS_QPT; ++h) quad(gt + h * QS_GT, false);
                } else {
                    for (int h = 0; h < QS_QPT; ++h) {
                        const uint32_t q = gt + h * QS_GT;
                        if (q * 4 >= na) break;
                        quad(q, true);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        // ---- group reduction and solve ------------------------------------------------------
        {
            double sum[13];
            if (QS_GW == 1) {
#pragma unroll
                for (int k = 0; k < 13; ++k) sum[k] = warp_sum(acc[k]);
            } else {
                const int gw = (tid % QS_GT) >> 5;
#pragma unroll
                for (int k = 0; k < 13; ++k) {
                    const double v = warp_sum(acc[k]);
                    if (lane == 0) red[QS_GW > 1 ? g : 0][k * QS_GW + gw] = v;
                }
                group_sync(g);
                if (gt == 0)
                    for (int k = 0; k < 13; ++k) {
                        double v = 0.0;
                        for (int w4 = 0; w4 < QS_GW; ++w4) v += red[QS_GW > 1 ? g : 0][k * QS_GW + w4];
                        sum[k] = v;
                    }
            }
            if (gt == 0 && active) {
                const double m0 = sum[10] / W, m1 = sum[11] / W, m2 = sum[12] / W;
                const double mx = px + m0, my = py + m1, mz = pz + m2;
                const double sx[3] = { p.refinfo->sx[0], p.refinfo->sx[1], p.refinfo->sx[2] };
                const double m[3] = { m0, m1, m2 };
                double M[9], A[9], R[9];
                for (int a = 0; a < 3; ++a)
                    for (int bb = 0; bb < 3; ++bb) M[3 * a + bb] = sum[3 * a + bb] - m[a] * sx[bb];
                const double G1 = sum[9] - W * (m0 * m0 + m1 * m1 + m2 * m2);
                const bool need_rot = p.rot_out != nullptr || p.aligned != nullptr;
                bool close = false;
                if (p.allclose_shortcut) {
                    // np.allclose(mobile, ref) implies sum |a - b|^2 <= close_sum; only then look elementwise
                    const double d2 = G1 + p.refinfo->g2 - 2.0 * (M[0] + M[4] + M[8]);
                    if (d2 <= p.refinfo->close_sum * 1.001 + 1e-9 * (G1 + p.refinfo->g2)) {
                        close = true;
                        const float *frame = p.traj + f * p.natoms * 3;
                        for (uint64_t i = 0; i < p.nsel && close; ++i)
                            close = qcp_close((double)frame[3 * i] - mx, p.refc[3 * i]) &&
                                    qcp_close((double)frame[3 * i + 1] - my, p.refc[3 * i + 1]) &&
                                    qcp_close((double)frame[3 * i + 2] - mz, p.refc[3 * i + 2]);
                    }
                }
                double rms;
                if (close) {
                    rms = 0.0;
                    for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
                } else {
                    if (p.convention == 0) {
                        for (int k = 0; k < 9; ++k) A[k] = M[k];
                    } else {
                        for (int a = 0; a < 3; ++a)
                            for (int bb = 0; bb < 3; ++bb) A[3 * a + bb] = M[3 * bb + a];
                    }
                    rms = qcp_solve(need_rot ? R : nullptr, A, (G1 + p.refinfo->g2) * 0.5, (double)p.nsel);
                }
                p.rmsd_out[f] = rms;
                if (p.rot_out)
                    for (int k = 0; k < 9; ++k) p.rot_out[f * 9 + k] = R[k];
                if (p.aligned) {
                    for (int a = 0; a < 3; ++a)
                        for (int bb = 0; bb < 3; ++bb)
                            s_xf[g][3 * a + bb] = p.convention == 0 ? R[3 * a + bb] : R[3 * bb + a];   // x.R or x.R^T
                    s_xf[g][9] = mx; s_xf[g][10] = my; s_xf[g][11] = mz;
                }
            }
            group_sync(g);        // red[] reusable, s_xf visible
        }
        if (nchB == 0) continue;
        // ---- second pass: rigid transform of every atom (core/groups.py:1480, 1527) --------
        double R[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = s_xf[g][k];
#if QCP_DF_TRANSLATE
        // The two translations as float32 "double-float" sums (see df_translate2): t = -com and
        // t = +ref_com, each split into hi + lo, paired the way the coordinates of a quad pair up
        // ({x y} {z x} {y z}).  The rotation stays in float64.
        const DFShift sub = df_shift(-s_xf[g][9], -s_xf[g][10], -s_xf[g][11]);
        const DFShift add = df_shift(rx, ry, rz);
        auto rotate = [&](float p0, float p1, float p2, float &r0, float &r1, float &r2) {
            const double d0 = (double)p0, d1 = (double)p1, d2 = (double)p2;
            r0 = (float)fma(d2, R[6], fma(d1, R[3], d0 * R[0]));
            r1 = (float)fma(d2, R[7], fma(d1, R[4], d0 * R[1]));
            r2 = (float)fma(d2, R[8], fma(d1, R[5], d0 * R[2]));
        };
        auto move_quad = [&](float4 &v0, float4 &v1, float4 &v2) {
            float2 a = df_translate2(make_float2(v0.x, v0.y), sub.xy_h, sub.xy_l);
            float2 b = df_translate2(make_float2(v0.z, v0.w), sub.zx_h, sub.zx_l);
            float2 c = df_translate2(make_float2(v1.x, v1.y), sub.yz_h, sub.yz_l);
            float2 d = df_translate2(make_float2(v1.z, v1.w), sub.xy_h, sub.xy_l);
            float2 e = df_translate2(make_float2(v2.x, v2.y), sub.zx_h, sub.zx_l);
            float2 f2 = df_translate2(make_float2(v2.z, v2.w), sub.yz_h, sub.yz_l);
            float4 r0, r1, r2;
            rotate(a.x, a.y, b.x, r0.x, r0.y, r0.z);
            rotate(b.y, c.x, c.y, r0.w, r1.x, r1.y);
            rotate(d.x, d.y, e.x, r1.z, r1.w, r2.x);
            rotate(e.y, f2.x, f2.y, r2.y, r2.z, r2.w);
            a = df_translate2(make_float2(r0.x, r0.y), add.xy_h, add.xy_l);
            b = df_translate2(make_float2(r0.z, r0.w), add.zx_h, add.zx_l);
            c = df_translate2(make_float2(r1.x, r1.y), add.yz_h, add.yz_l);
            d = df_translate2(make_float2(r1.z, r1.w), add.xy_h, add.xy_l);
            e = df_translate2(make_float2(r2.x, r2.y), add.zx_h, add.zx_l);
            f2 = df_translate2(make_float2(r2.z, r2.w), add.yz_h, add.yz_l);
            v0 = make_float4(a.x, a.y, b.x, b.y);
            v1 = make_float4(c.x, c.y, d.x, d.y);
            v2 = make_float4(e.x, e.y, f2.x, f2.y);
        };
#else
        const double cx = s_xf[g][9], cy = s_xf[g][10], cz = s_xf[g][11];
        auto move = [&](float x, float y, float z, float &ox, float &oy, float &oz) {
            const float p0 = (float)(f2d_alu(x) - cx);
            const float p1 = (float)(f2d_alu(y) - cy);
            const float p2 = (float)(f2d_alu(z) - cz);
            const double d0 = f2d_alu(p0), d1 = f2d_alu(p1), d2 = f2d_alu(p2);
            const float r0 = (float)fma(d2, R[6], fma(d1, R[3], d0 * R[0]));
            const float r1 = (float)fma(d2, R[7], fma(d1, R[4], d0 * R[1]));
            const float r2 = (float)fma(d2, R[8], fma(d1, R[5], d0 * R[2]));
            ox = (float)(f2d_alu(r0) + rx);
            oy = (float)(f2d_alu(r1) + ry);
            oz = (float)(f2d_alu(r2) + rz);
        };
        auto move_quad = [&](float4 &v0, float4 &v1, float4 &v2) {
            float4 o0, o1, o2;
            move(v0.x, v0.y, v0.z, o0.x, o0.y, o0.z);
            move(v0.w, v1.x, v1.y, o0.w, o1.x, o1.y);
            move(v1.z, v1.w, v2.x, o1.z, o1.w, o2.x);
            move(v2.y, v2.z, v2.w, o2.y, o2.z, o2.w);
            v0 = o0; v1 = o1; v2 = o2;
        };
#endif
        for (uint32_t c = 0; c < nchB; ++c, ++it) {
            const int s = it % QS_STAGES;
            mbar_wait(&full_bar[s], (it / QS_STAGES) & 1u);
            unsigned char *st = qs_smem + (size_t)s * STAGE_BYTES;
            float4 *fr4 = reinterpret_cast<float4 *>(st + FRAME_OFF + g * QS_FRAME_BYTES);
            float *gdst = p.aligned + (f * p.natoms + (uint64_t)c * QS_CH) * 3;
            const uint32_t na = (uint32_t)min((uint64_t)QS_CH, p.natoms - (uint64_t)c * QS_CH);   // multiple of 4
            if (active) {
                float4 *dst4 = QS_BULK_STORE ? fr4 : reinterpret_cast<float4 *>(gdst);
#pragma unroll
                for (int h = 0; h < QS_QPT; ++h) {
                    const uint32_t q = gt + h * QS_GT;
                    if (q * 4 >= na) break;
                    float4 v0 = fr4[3 * q], v1 = fr4[3 * q + 1], v2 = fr4[3 * q + 2];
                    move_quad(v0, v1, v2);
                    dst4[3 * q] = v0; dst4[3 * q + 1] = v1; dst4[3 * q + 2] = v2;
                }
            }
            if (QS_BULK_STORE) {
                // the transformed chunk leaves through the async proxy; the stage is handed back one
                // chunk later, when that store has read it
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                group_sync(g);
                if (gt == 0) {
                    if (pending >= 0) {
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        mbar_arrive_n(&empty_bar[pending], QS_GW);
                    }
                    if (active) bulk_s2g(gdst, fr4, na * 12u, pol_out);
                    pending = s;
                }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
            }
        }
        if (QS_BULK_STORE && gt == 0 && pending >= 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive_n(&empty_bar[pending], QS_GW);
            pending = -1;
        }
    }
    if (QS_BULK_STORE && gt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// single pair of already-centred float64 structures (CalcRMSDRotationalMatrix)
__global__ void __launch_bounds__(QCP_THREADS)
qcp_single_kernel(const double *__restrict__ ref, const double *__restrict__ conf, uint64_t n,
                  const double *__restrict__ w, double *__restrict__ rot, double *__restrict__ rmsd,
                  int inner_only)
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
        const double E0 = (acc[9] + acc[10]) * 0.5;
        if (inner_only) {                                // InnerProduct (qcprot.pyx:179-277): A[9] and E0
            for (int k = 0; k < 9; ++k) rot[k] = acc[k];
            *rmsd = E0;
            return;
        }
        double R[9];
        *rmsd = qcp_solve(rot ? R : nullptr, acc, E0, (double)n);
        if (rot)
            for (int k = 0; k < 9; ++k) rot[k] = R[k];
    }
}

// FastCalcRMSDAndRotation (qcprot.pyx:326-543) on its own: the quartic, Newton, the quaternion
__global__ void qcp_solve_kernel(const double *__restrict__ A, double E0, double n, double *__restrict__ rot,
                                 double *__restrict__ rmsd)
{
    double a[9], R[9];
    for (int k = 0; k < 9; ++k) a[k] = A[k];
    *rmsd = qcp_solve(rot ? R : nullptr, a, E0, n);
    if (rot)
        for (int k = 0; k < 9; ++k) rot[k] = R[k];
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
               weights ? w.as<double>() : nullptr, rot ? ro.as<double>() : nullptr, val.as<double>(), 0);
    MDA_CHECK_LAUNCH();
    MDA_CUDA(cudaMemcpyAsync(rmsd, val.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (rot) MDA_CUDA(ro.finish());
    MDA_CUDA(cudaStreamSynchronize(st));
    return MDAGPU_OK;
}

extern "C" int mdagpu_qcp_inner_product(const double *coords1, const double *coords2, uint64_t n,
                                        const double *weights, double *A, double *E0, int space,
                                        mdagpu_stream_t stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    MDA_REQUIRE(coords1 && coords2 && A && E0, "mdagpu_qcp_inner_product: NULL argument");
    MDA_REQUIRE(n > 0, "mdagpu_qcp_inner_product: empty structures");
    StagedIn c1, c2, w;
    StagedOut a;
    DevBuf val;
    MDA_CUDA(c1.init(coords1, n * 3 * sizeof(double), space, st));
    MDA_CUDA(c2.init(coords2, n * 3 * sizeof(double), space, st));
    MDA_CUDA(w.init(weights, n * sizeof(double), space, st));
    MDA_CUDA(a.init(A, 9 * sizeof(double), space, st));
    MDA_CUDA(val.alloc(sizeof(double), st));
    // the kernel's first structure is coords2 ("ref"), its second coords1 ("conf"): A's rows follow coords1
    MDA_LAUNCH(qcp_single_kernel, 1, QCP_THREADS, 0, st, c2.as<double>(), c1.as<double>(), n,
               weights ? w.as<double>() : nullptr, a.as<double>(), val.as<double>(), 1);
    MDA_CHECK_LAUNCH();
    MDA_CUDA(cudaMemcpyAsync(E0, val.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    MDA_CUDA(a.finish());
    MDA_CUDA(cudaStreamSynchronize(st));
    return MDAGPU_OK;
}

extern "C" int mdagpu_qcp_solve(const double *A, double E0, double n, double *rot, double *rmsd)
{
    MDA_REQUIRE(A && rmsd, "mdagpu_qcp_solve: NULL argument");
    cudaStream_t st = nullptr;
    StagedIn a;
    StagedOut ro;
    DevBuf val;
    MDA_CUDA(a.init(A, 9 * sizeof(double), MDAGPU_HOST, st));
    MDA_CUDA(ro.init(rot, 9 * sizeof(double), MDAGPU_HOST, st));
    MDA_CUDA(val.alloc(sizeof(double), st));
    MDA_LAUNCH(qcp_solve_kernel, 1, 1, 0, st, a.as<double>(), E0, n, rot ? ro.as<double>() : nullptr, val.as<double>());
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
    DevBuf info, refc;
    MDA_CUDA(info.alloc(sizeof(RefInfo), st));
    MDA_CUDA(refc.alloc(nsel * 3 * sizeof(double), st));
    const double *wd = weights ? w.as<double>() : nullptr;
    const int32_t *sd = sel ? s.as<int32_t>() : nullptr;
    MDA_LAUNCH(qcp_ref_kernel, 1, QCP_THREADS, 0, st, r.as<float>(), wd, nsel, info.as<RefInfo>(), refc.as<double>());
    MDA_CHECK_LAUNCH();

    // 16-byte path: every frame starts on a 16-byte boundary and the fitted atoms are a prefix
    const bool vec = sel == nullptr && natoms % 4 == 0 && ((uintptr_t)t.dev % 16 == 0) &&
                     (!aligned || (uintptr_t)al.dev % 16 == 0) && (!weights || (uintptr_t)wd % 16 == 0);
    MDA_REQUIRE(nframes <= 0x7fffffffull, "mdagpu_qcp_trajectory: too many frames in one call");
    if (vec && nsel >= 2048) {
        // TMA-streamed persistent kernel
        const uint64_t nch = (nsel + QS_CH - 1) / QS_CH;
        DevBuf refp, wp;
        MDA_CUDA(refp.alloc(nch * QS_REF_BYTES, st));
        if (weights) MDA_CUDA(wp.alloc(nch * QS_W_BYTES, st));
        MDA_LAUNCH(qcp_ref_planes_kernel, cdiv(nch * (QS_CH / 4), 256), 256, 0, st, refc.as<double>(), wd, nsel, nch,
                   refp.as<double2>(), weights ? wp.as<double2>() : nullptr);
        MDA_CHECK_LAUNCH();
        QcpStreamParams qp;
        qp.traj = t.as<float>(); qp.nframes = nframes; qp.natoms = natoms; qp.nsel = nsel;
        qp.refc = refc.as<double>(); qp.refp = refp.as<double2>(); qp.wp = weights ? wp.as<double2>() : nullptr;
        qp.refinfo = info.as<RefInfo>(); qp.rmsd_out = rm.as<double>(); qp.rot_out = rot ? ro.as<double>() : nullptr;
        qp.aligned = aligned ? al.as<float>() : nullptr; qp.convention = convention; qp.allclose_shortcut = allclose_shortcut;
        // one CTA per SM, each with a contiguous range of >= QS_FB frames (fewer CTAs for short trajectories)
        const uint64_t nbatch = (nframes + QS_FB - 1) / QS_FB;
        const unsigned grid = (unsigned)(nbatch < (uint64_t)num_sms() ? nbatch : (uint64_t)num_sms());
        // ring depth: as many stages as fit in the 227 kB of one SM
        constexpr int SW = QS_STAGES_W, SU = QS_STAGES_U;
        if (weights) {
            const size_t smem = SW * (size_t)(QS_REF_BYTES + QS_W_BYTES + QS_FB * QS_FRAME_BYTES);
            MDA_CUDA(cudaFuncSetAttribute(qcp_stream_kernel<true, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            MDA_LAUNCH_AS("qcp_stream_kernel", (qcp_stream_kernel<true, SW>), grid, QS_THREADS + 32, smem, st, qp);
        } else {
            const size_t smem = SU * (size_t)(QS_REF_BYTES + QS_FB * QS_FRAME_BYTES);
            MDA_CUDA(cudaFuncSetAttribute(qcp_stream_kernel<false, SU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            MDA_LAUNCH_AS("qcp_stream_kernel", (qcp_stream_kernel<false, SU>), grid, QS_THREADS + 32, smem, st, qp);
        }
        MDA_CHECK_LAUNCH();
        MDA_CUDA(rm.finish());
        if (rot) MDA_CUDA(ro.finish());
        if (aligned) MDA_CUDA(al.finish());
        return MDAGPU_OK;
    }
#define QCP_LAUNCH(T, W, V)                                                                      \
    MDA_LAUNCH_AS("qcp_traj_kernel", (qcp_traj_kernel<T, W, V>), (unsigned)nframes, T, 0, st, t.as<float>(), natoms, sd, nsel, \
                  refc.as<double>(), wd, info.as<RefInfo>(), rm.as<double>(),                    \
                  rot ? ro.as<double>() : nullptr, aligned ? al.as<float>() : nullptr,           \
                  convention, allclose_shortcut)
    // small selections: one warp-sized CTA would idle most lanes in the reduction; large: 256 threads
    if (nsel >= 2048) {
        if (weights) { if (vec) QCP_LAUNCH(256, true, true); else QCP_LAUNCH(256, true, false); }
        else { if (vec) QCP_LAUNCH(256, false, true); else QCP_LAUNCH(256, false, false); }
    } else {
        if (weights) { if (vec) QCP_LAUNCH(64, true, true); else QCP_LAUNCH(64, true, false); }
        else { if (vec) QCP_LAUNCH(64, false, true); else QCP_LAUNCH(64, false, false); }
    }
#undef QCP_LAUNCH
    MDA_CHECK_LAUNCH();
    MDA_CUDA(rm.finish());
    if (rot) MDA_CUDA(ro.finish());
    if (aligned) MDA_CUDA(al.finish());
    return MDAGPU_OK;
}
