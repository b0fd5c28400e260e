// geom.cuh -- exact per-pair / per-atom arithmetic of the geometry hot path.
//
// Every function here reproduces, operation for operation, the IEEE-754
// arithmetic of the reference's serial C code so that device results are
// bit-identical to it.  The translation units that include this header are
// compiled with --fmad=false, so "a*b + c" below is two roundings, exactly as
// in the reference build (package/setup.py:262-263: -O3, no -ffast-math, no
// -march, hence no FMA contraction on x86-64).  Where a fused op is wanted
// (FP32 pre-filters) it is written explicitly as fmaf().
//
// Reference (paths relative to package/MDAnalysis/lib/):
//   include/calc_distances.h:39-49    minimum_image            -> mic_ortho_round
//   include/calc_distances.h:51-70    _minimum_image_ortho_lazy-> mic_ortho_lazy
//   include/calc_distances.h:72-114   minimum_image_triclinic  -> mic_triclinic
//   include/calc_distances.h:116-190  _ortho_pbc               -> wrap_ortho_axis
//   include/calc_distances.h:192-379  _triclinic_pbc           -> wrap_triclinic
//   include/calc_distances.h:726-762  _calc_dihedral_angle     -> torsion
//   nsgrid.pyx:435-481                coord2cellxyz / cellxyz2cellid
//   qcprot.pyx:326-543                FastCalcRMSDAndRotation  -> qcp_solve
//
// The functions are __host__ __device__ so that tests/host_check.cu can run
// the very same source on the CPU against the oracle without a GPU; the
// product library only ever calls them from kernels.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MDA_HD __host__ __device__ __forceinline__
#else
#define MDA_HD inline
#endif

namespace mdagpu {

enum { BOX_NONE = 0, BOX_ORTHO = 1, BOX_TRICLINIC = 2 };

// Minimum-image flavour used by a pair kernel.
enum {
    MIC_NONE = 0,         // plain difference
    MIC_ORTHO_ROUND = 1,  // lib.distances with an orthorhombic box
    MIC_ORTHO_LAZY = 2,   // FastNS with an orthorhombic box (wrapped inputs)
    MIC_TRICLINIC = 3     // 27-image search (wrapped inputs)
};

// All box-derived constants a kernel needs, prepared once on the host.
struct BoxParams {
    int mode;        // BOX_*
    float L[3];      // orthorhombic edge lengths (float32 as handed over)
    float invL[3];   // (float)(1.0 / L)               calc_distances.h:402-405
    float half[3];   // 0.5f * L                       nsgrid.pyx:316
    float h[9];      // row-major lower-triangular cell matrix
    double winv[3];  // 1.0 / (double) L               calc_distances.h:135-137
    // triclinic wrap constants                         calc_distances.h:221-230
    double r0, r4, r8, r3, r6, r7, ay, az, bz;
    // FP32 thresholds on |dx| equivalent to the double tests on s = invL*dx
    // in mic_ortho_round (valid because invL*dx is exact in double):
    float x_half[3];  // smallest float x with (double)invL*x >= 0.5
    float x_lim[3];   // below this |x|, |s| < 1.5 is guaranteed
    // triclinic 8-image shortcut (mic_triclinic_fast): 1 / diagonal, tilt ratios, and the
    // squared smallest perpendicular width of the cell (with a safety factor); 0 disables it
    float tri_inv[3], tri_t3, tri_t6, tri_t7, tri_w2;
};

static inline BoxParams make_box_params(int mode, const float *box)
{
    BoxParams b;
    b.mode = mode;
    for (int k = 0; k < 3; ++k) {
        b.L[k] = b.invL[k] = b.half[k] = 0.0f;
        b.winv[k] = 0.0;
        b.x_half[k] = b.x_lim[k] = 0.0f;
    }
    for (int k = 0; k < 9; ++k) b.h[k] = 0.0f;
    b.r0 = b.r4 = b.r8 = b.r3 = b.r6 = b.r7 = b.ay = b.az = b.bz = 0.0;
    b.tri_inv[0] = b.tri_inv[1] = b.tri_inv[2] = b.tri_t3 = b.tri_t6 = b.tri_t7 = b.tri_w2 = 0.0f;
    if (mode == BOX_ORTHO) {
        for (int k = 0; k < 3; ++k) {
            b.L[k] = box[k];
            b.invL[k] = (float)(1.0 / box[k]);
            b.half[k] = (float)(0.5 * (double)box[k]);
            b.winv[k] = 1.0 / (double)box[k];
            if (box[k] > FLT_EPSILON && box[k] < 1e30f && b.invL[k] > 1e-30f && b.invL[k] < 1e30f) {
                // smallest float x such that invL*x >= 0.5 (exact double product)
                float x = (float)(0.5 / (double)b.invL[k]);
                while ((double)b.invL[k] * (double)x >= 0.5) x = nextafterf(x, 0.0f);
                while ((double)b.invL[k] * (double)x < 0.5) x = nextafterf(x, INFINITY);
                b.x_half[k] = x;
                float y = (float)(1.4 / (double)b.invL[k]);
                b.x_lim[k] = y;
            }
        }
    } else if (mode == BOX_TRICLINIC) {
        for (int k = 0; k < 9; ++k) b.h[k] = box[k];
        b.r0 = 1.0 / (double)box[0];
        b.r4 = 1.0 / (double)box[4];
        b.r8 = 1.0 / (double)box[8];
        b.r3 = -box[3] * b.r0 * b.r4;
        b.r6 = (-b.r3 * box[7] - box[6] * b.r0) * b.r8;
        b.r7 = -box[7] * b.r4 * b.r8;
        b.ay = (double)box[3] * b.r4;
        b.az = (double)box[6] * b.r8;
        b.bz = (double)box[7] * b.r8;
        // perpendicular widths w_k = volume / |cross of the other two cell vectors|
        const double a0[3] = { box[0], box[1], box[2] }, a1[3] = { box[3], box[4], box[5] }, a2[3] = { box[6], box[7], box[8] };
        const double c12[3] = { a1[1] * a2[2] - a1[2] * a2[1], a1[2] * a2[0] - a1[0] * a2[2], a1[0] * a2[1] - a1[1] * a2[0] };
        const double c20[3] = { a2[1] * a0[2] - a2[2] * a0[1], a2[2] * a0[0] - a2[0] * a0[2], a2[0] * a0[1] - a2[1] * a0[0] };
        const double c01[3] = { a0[1] * a1[2] - a0[2] * a1[1], a0[2] * a1[0] - a0[0] * a1[2], a0[0] * a1[1] - a0[1] * a1[0] };
        const double vol = fabs(a0[0] * c12[0] + a0[1] * c12[1] + a0[2] * c12[2]);
        const double n12 = sqrt(c12[0] * c12[0] + c12[1] * c12[1] + c12[2] * c12[2]);
        const double n20 = sqrt(c20[0] * c20[0] + c20[1] * c20[1] + c20[2] * c20[2]);
        const double n01 = sqrt(c01[0] * c01[0] + c01[1] * c01[1] + c01[2] * c01[2]);
        const bool sane = box[0] > 0.0f && box[4] > 0.0f && box[8] > 0.0f && box[1] == 0.0f && box[2] == 0.0f &&
                          box[5] == 0.0f && vol > 0.0 && vol < 1e30 && n12 > 0.0 && n20 > 0.0 && n01 > 0.0;
        if (sane) {
            double w = vol / n12;
            if (vol / n20 < w) w = vol / n20;
            if (vol / n01 < w) w = vol / n01;
            w *= 1.0 - 2e-3;
            b.tri_inv[0] = (float)b.r0; b.tri_inv[1] = (float)b.r4; b.tri_inv[2] = (float)b.r8;
            b.tri_t3 = box[3]; b.tri_t6 = box[6]; b.tri_t7 = box[7];
            b.tri_w2 = (float)(w * w);
        }
    }
    return b;
}

struct D3 { double x, y, z; };
struct F3 { float x, y, z; };

// float32 subtraction, widened afterwards (calc_distances.h:390-392)
MDA_HD D3 sep_f32(float px, float py, float pz, float qx, float qy, float qz)
{
    float tx = px - qx, ty = py - qy, tz = pz - qz;
    D3 d;
    d.x = (double)tx; d.y = (double)ty; d.z = (double)tz;
    return d;
}

MDA_HD double norm2(const D3 &d) { return (d.x * d.x) + (d.y * d.y) + (d.z * d.z); }

// ---- orthorhombic, round-based (calc_distances.h:39-49) -------------------
MDA_HD double mic_round_axis(double x, float L, float invL)
{
    if (L > FLT_EPSILON) {
        double s = (double)invL * x;
        return (double)L * (s - round(s));
    }
    return x;
}
// Same result, cheaper: tx is the float32 difference that x was widened from.
// |s| < 1.5 is decided on tx in FP32 (s = invL*tx is an exact double product),
// in which case round(s) is -1, 0 or +1 by two FP32 compares.
MDA_HD double mic_round_axis_fast(float tx, float L, float invL, float x_half, float x_lim)
{
    double x = (double)tx;
    if (!(L > FLT_EPSILON)) return x;
    double s = (double)invL * x;
    double r;
    if (fabsf(tx) < x_lim)
        r = (tx >= x_half) ? 1.0 : ((tx <= -x_half) ? -1.0 : 0.0);
    else
        r = round(s);
    return (double)L * (s - r);
}
MDA_HD void mic_ortho_round(D3 &d, const BoxParams &b)
{
    d.x = mic_round_axis(d.x, b.L[0], b.invL[0]);
    d.y = mic_round_axis(d.y, b.L[1], b.invL[1]);
    d.z = mic_round_axis(d.z, b.L[2], b.invL[2]);
}

// ---- orthorhombic, lazy (calc_distances.h:51-70) ---------------------------
MDA_HD double mic_lazy_axis(double x, float L, float half)
{
    if (x > (double)half) return x - (double)L;
    if (x <= -(double)half) return x + (double)L;
    return x;
}
MDA_HD void mic_ortho_lazy(D3 &d, const BoxParams &b)
{
    d.x = mic_lazy_axis(d.x, b.L[0], b.half[0]);
    d.y = mic_lazy_axis(d.y, b.L[1], b.half[1]);
    d.z = mic_lazy_axis(d.z, b.L[2], b.half[2]);
}

// ---- triclinic, exhaustive 27 images (calc_distances.h:72-114) -------------
// Returns the squared length of the winning image and overwrites d with it.
MDA_HD double mic_triclinic(D3 &d, const float *h)
{
    double best = (double)FLT_MAX;
    double bx = 0.0, by = 0.0, bz = 0.0;
    for (int a = -1; a <= 1; ++a) {
        double xa = d.x + (double)(h[0] * (float)a);
        for (int b = -1; b <= 1; ++b) {
            double xab = xa + (double)(h[3] * (float)b);
            double yb = d.y + (double)(h[4] * (float)b);
            for (int c = -1; c <= 1; ++c) {
                double x = xab + (double)(h[6] * (float)c);
                double y = yb + (double)(h[7] * (float)c);
                double z = d.z + (double)(h[8] * (float)c);
                double r2 = x * x + y * y + z * z;
                if (r2 < best) { best = r2; bx = x; by = y; bz = z; }
            }
        }
    }
    d.x = bx; d.y = by; d.z = bz;
    return best;
}

// One image (a,b,c) evaluated exactly as the loop above would.
MDA_HD double tric_image_d2(const D3 &d, const float *h, int a, int b, int c, D3 &img)
{
    double xa = d.x + (double)(h[0] * (float)a);
    double xab = xa + (double)(h[3] * (float)b);
    double yb = d.y + (double)(h[4] * (float)b);
    img.x = xab + (double)(h[6] * (float)c);
    img.y = yb + (double)(h[7] * (float)c);
    img.z = d.z + (double)(h[8] * (float)c);
    return img.x * img.x + img.y * img.y + img.z * img.z;
}

// The reference's loop restricted to the eight images n_k in {0, k1}, visited in the loop's own
// order (a outermost, -1 < 0 < 1, strict '<'): when its winner is known to be one of the eight
// (see mic_triclinic_fast8) this returns what the 27-image loop returns, near-ties included.
MDA_HD double tric_eight_exact(const D3 &d, const float *h, int a1, int b1, int c1, D3 &img)
{
    double best = (double)FLT_MAX;
    double bx = 0.0, by = 0.0, bz = 0.0;
    for (int ia = 0; ia < 2; ++ia) {
        const int a = (ia == 0) == (a1 < 0) ? a1 : 0;
        double xa = d.x + (double)(h[0] * (float)a);
        for (int ib = 0; ib < 2; ++ib) {
            const int b = (ib == 0) == (b1 < 0) ? b1 : 0;
            double xab = xa + (double)(h[3] * (float)b);
            double yb = d.y + (double)(h[4] * (float)b);
            for (int ic = 0; ic < 2; ++ic) {
                const int c = (ic == 0) == (c1 < 0) ? c1 : 0;
                double x = xab + (double)(h[6] * (float)c);
                double y = yb + (double)(h[7] * (float)c);
                double z = d.z + (double)(h[8] * (float)c);
                double r2 = x * x + y * y + z * z;
                if (r2 < best) { best = r2; bx = x; by = y; bz = z; }
            }
        }
    }
    img.x = bx; img.y = by; img.z = bz;
    return best;
}

// Bit-identical to mic_triclinic(), but picks the winning image with an FP32
// scan first; the double evaluation is done for the winner only.  If the FP32
// scan cannot separate winner and runner-up by a safe margin (ties on lattice
// data, huge offsets) it falls back to the exhaustive double loop.
MDA_HD double mic_triclinic_fast(float tx, float ty, float tz, const float *h, D3 &out)
{
    float best = FLT_MAX, second = FLT_MAX;
    int ba = 0, bb = 0, bc = 0;
#pragma unroll
    for (int a = -1; a <= 1; ++a) {
        float xa = tx + h[0] * (float)a;
#pragma unroll
        for (int b = -1; b <= 1; ++b) {
            float xab = xa + h[3] * (float)b;
            float yb = ty + h[4] * (float)b;
#pragma unroll
            for (int c = -1; c <= 1; ++c) {
                float x = xab + h[6] * (float)c;
                float y = yb + h[7] * (float)c;
                float z = tz + h[8] * (float)c;
                float r2 = fmaf(x, x, fmaf(y, y, z * z));
                if (r2 < best) { second = best; best = r2; ba = a; bb = b; bc = c; }
                else if (r2 < second) second = r2;
            }
        }
    }
    D3 d;
    d.x = (double)tx; d.y = (double)ty; d.z = (double)tz;
    // FP32 evaluation error of each r2 is below 8 ulp-equivalents (2^-20
    // relative) of the largest r2 term involved; demand a much wider gap.
    if (second - best > 1e-4f * second && best < 1e30f)
        return tric_image_d2(d, h, ba, bb, bc, out);
    out = d;
    return mic_triclinic(out, h);
}

// The same winner from EIGHT images when that is provable.  With s = d H^-1 the fractional
// coordinates of d, image (a, b, c) has fractional coordinates (s + n) and is at least
// w_k |s_k + n_k| long (w_k: perpendicular width of the cell along axis k).  The 19 images
// that shift some axis the wrong way (|s_k + n_k| >= 1) are therefore >= min w_k long: if the
// best of the other eight (n_k in {0, -sign(s_k)}) is shorter than that, it is the winner of
// the reference's 27-image loop.  Otherwise -- and whenever the FP32 values cannot separate
// winner and runner-up -- the exhaustive path above decides.  (A wrong sign of an s_k near
// zero only swaps one wrong-way image for another.)
MDA_HD double mic_triclinic_fast8(float tx, float ty, float tz, const BoxParams &bp, D3 &out)
{
    const float *h = bp.h;
    if (bp.tri_w2 > 0.0f) {
        const float sz = tz * bp.tri_inv[2];
        const float sy = (ty - sz * bp.tri_t7) * bp.tri_inv[1];
        const float sx = (tx - sy * bp.tri_t3 - sz * bp.tri_t6) * bp.tri_inv[0];
        const int a1 = sx >= 0.0f ? -1 : 1, b1 = sy >= 0.0f ? -1 : 1, c1 = sz >= 0.0f ? -1 : 1;
        float best = FLT_MAX, second = FLT_MAX;
        int ba = 0, bb = 0, bc = 0;
#pragma unroll
        for (int ia = 0; ia < 2; ++ia) {
            const int a = ia ? a1 : 0;
            const float xa = tx + h[0] * (float)a;
#pragma unroll
            for (int ib = 0; ib < 2; ++ib) {
                const int b = ib ? b1 : 0;
                const float xab = xa + h[3] * (float)b;
                const float yb = ty + h[4] * (float)b;
#pragma unroll
                for (int ic = 0; ic < 2; ++ic) {
                    const int c = ic ? c1 : 0;
                    const float x = xab + h[6] * (float)c;
                    const float y = yb + h[7] * (float)c;
                    const float z = tz + h[8] * (float)c;
                    const float r2 = fmaf(x, x, fmaf(y, y, z * z));
                    if (r2 < best) { second = best; best = r2; ba = a; bb = b; bc = c; }
                    else if (r2 < second) second = r2;
                }
            }
        }
        if (best < bp.tri_w2) {
            D3 d;
            d.x = (double)tx; d.y = (double)ty; d.z = (double)tz;
            if (second - best > 1e-4f * second) return tric_image_d2(d, h, ba, bb, bc, out);
            return tric_eight_exact(d, h, a1, b1, c1, out);
        }
    }
    return mic_triclinic_fast(tx, ty, tz, h, out);
}

// ---- wraps -------------------------------------------------------------------
MDA_HD float wrap_ortho_axis(float x, float L, double invL)
{
    float stored = x, probe = x;
    if (probe < 0.0f) {
        probe += L;
        if (probe < 0.0f) {
            int s = (int)floor((double)stored * invL);
            stored -= (float)s * L;
            if (stored < 0.0f) stored += L;
        } else {
            stored = probe;
        }
    }
    if (probe >= L) {
        probe -= L;
        if (probe >= L) {
            int s = (int)floor((double)stored * invL);
            stored -= (float)s * L;
            if (stored >= L) stored -= L;
        } else {
            stored = probe;
        }
    }
    return stored;
}

MDA_HD void wrap_ortho(float &x, float &y, float &z, const BoxParams &b)
{
    if (b.L[0] == 0.0f && b.L[1] == 0.0f && b.L[2] == 0.0f) return;
    x = wrap_ortho_axis(x, b.L[0], b.winv[0]);
    y = wrap_ortho_axis(y, b.L[1], b.winv[1]);
    z = wrap_ortho_axis(z, b.L[2], b.winv[2]);
}

MDA_HD int tric_single_pass(float *c, const BoxParams &w, bool strict)
{
    const float *h = w.h;
    if (c[2] < 0.0f) {
        c[0] += h[6]; c[1] += h[7]; c[2] += h[8];
        if (strict && c[2] < 0.0f) return 1;
    }
    if (c[2] >= h[8]) {
        c[0] -= h[6]; c[1] -= h[7]; c[2] -= h[8];
        if (strict && c[2] >= h[8]) return 1;
    }
    double lo = c[2] * w.bz;
    double hi = lo + h[4];
    if (c[1] < lo) {
        c[0] += h[3]; c[1] += h[4];
        if (strict && c[1] < lo) return 1;
    }
    if (c[1] >= hi) {
        c[0] -= h[3]; c[1] -= h[4];
        if (strict && c[1] >= hi) return 1;
    }
    lo = c[1] * w.ay + c[2] * w.az;
    hi = lo + h[0];
    if (c[0] < lo) {
        c[0] += h[0];
        if (strict && c[0] < lo) return 1;
    }
    if (c[0] >= hi) {
        c[0] -= h[0];
        if (strict && c[0] >= hi) return 1;
    }
    return 0;
}

MDA_HD void wrap_triclinic(float &x, float &y, float &z, const BoxParams &w)
{
    const float *h = w.h;
    if (h[0] == 0.0f && h[4] == 0.0f && h[8] == 0.0f) return;
    float c[3] = { x, y, z };
    if (tric_single_pass(c, w, true)) {
        float p0 = x, p1 = y, p2 = z;
        int s = (int)floor(p2 * w.r8);
        p2 -= (float)s * h[8];
        p1 -= (float)s * h[7];
        p0 -= (float)s * h[6];
        s = (int)floor(p1 * w.r4 + p2 * w.r7);
        p1 -= (float)s * h[4];
        p0 -= (float)s * h[3];
        s = (int)floor(p0 * w.r0 + p1 * w.r3 + p2 * w.r6);
        p0 -= (float)s * h[0];
        c[0] = p0; c[1] = p1; c[2] = p2;
        tric_single_pass(c, w, false);
    }
    x = c[0]; y = c[1]; z = c[2];
}

MDA_HD void wrap_any(float &x, float &y, float &z, const BoxParams &b)
{
    if (b.mode == BOX_ORTHO) wrap_ortho(x, y, z, b);
    else if (b.mode == BOX_TRICLINIC) wrap_triclinic(x, y, z, b);
}

// ---- exact squared pair distance for each minimum-image flavour ----------------
// p - q with the reference's operand order (caller picks which is which).
template <int MIC>
MDA_HD double pair_d2_exact(float px, float py, float pz, float qx, float qy, float qz,
                            const BoxParams &b)
{
    float tx = px - qx, ty = py - qy, tz = pz - qz;
    if (MIC == MIC_NONE) {
        double x = (double)tx, y = (double)ty, z = (double)tz;
        return (x * x) + (y * y) + (z * z);
    } else if (MIC == MIC_ORTHO_ROUND) {
        double x = mic_round_axis_fast(tx, b.L[0], b.invL[0], b.x_half[0], b.x_lim[0]);
        double y = mic_round_axis_fast(ty, b.L[1], b.invL[1], b.x_half[1], b.x_lim[1]);
        double z = mic_round_axis_fast(tz, b.L[2], b.invL[2], b.x_half[2], b.x_lim[2]);
        return (x * x) + (y * y) + (z * z);
    } else if (MIC == MIC_ORTHO_LAZY) {
        double x = mic_lazy_axis((double)tx, b.L[0], b.half[0]);
        double y = mic_lazy_axis((double)ty, b.L[1], b.half[1]);
        double z = mic_lazy_axis((double)tz, b.L[2], b.half[2]);
        return x * x + y * y + z * z;
    } else {
        D3 img;
        return mic_triclinic_fast8(tx, ty, tz, b, img);
    }
}

// Minimum-image vector (needed by angles / dihedrals), lib.distances flavours.
MDA_HD D3 pair_vec_exact(int mode, float px, float py, float pz, float qx, float qy, float qz,
                         const BoxParams &b)
{
    D3 d = sep_f32(px, py, pz, qx, qy, qz);
    if (mode == BOX_ORTHO) mic_ortho_round(d, b);
    else if (mode == BOX_TRICLINIC) mic_triclinic(d, b.h);
    return d;
}

// ---- angles / dihedrals (calc_distances.h:618-762) -------------------------------
MDA_HD double angle_from(const D3 &u, const D3 &v)
{
    double dot = u.x * v.x + u.y * v.y + u.z * v.z;
    double cx = u.y * v.z - u.z * v.y;
    double cy = -u.x * v.z + u.z * v.x;
    double cz = u.x * v.y - u.y * v.x;
    double cross = sqrt(cx * cx + cy * cy + cz * cz);
    return atan2(cross, dot);
}

MDA_HD double torsion(const D3 &va, const D3 &vb, const D3 &vc)
{
    double n1x = -va.y * vb.z + va.z * vb.y;
    double n1y = va.x * vb.z - va.z * vb.x;
    double n1z = -va.x * vb.y + va.y * vb.x;
    double n2x = -vb.y * vc.z + vb.z * vc.y;
    double n2y = vb.x * vc.z - vb.z * vc.x;
    double n2z = -vb.x * vc.y + vb.y * vc.x;
    double x = (n1x * n2x + n1y * n2y + n1z * n2z);
    double mx = n1y * n2z - n1z * n2y;
    double my = -n1x * n2z + n1z * n2x;
    double mz = n1x * n2y - n1y * n2x;
    double len_b = sqrt(vb.x * vb.x + vb.y * vb.y + vb.z * vb.z);
    double y = (mx * vb.x + my * vb.y + mz * vb.z) / len_b;
    if (fabs(x) == 0.0 && fabs(y) == 0.0) return NAN;
    return atan2(y, x);
}

// ---- FastNS grid geometry (nsgrid.pyx:294-481) -------------------------------------
struct GridGeom {
    int ncells[3];
    int periodic[3];
    double cs[9];      // cell widths on the diagonal, tilt ratios off it
    double origin[3];  // 0 for the reference's grids; lower corner of library-made grids
    int triclinic;
    int pbc;
    int64_t ncell_total;
};

MDA_HD void cell_of(const GridGeom &g, float x, float y, float z, int *c)
{
    // (double)v - 0.0 == (double)v exactly, so the origin does not disturb parity
    const double dx = (double)x - g.origin[0], dy = (double)y - g.origin[1], dz = (double)z - g.origin[2];
    c[2] = (int)(dz / g.cs[8]);
    c[1] = (int)((dy - dz * g.cs[7]) / g.cs[4]);
    c[0] = (int)((dx - dy * g.cs[3] - dz * g.cs[6]) / g.cs[0]);
    c[0] %= g.ncells[0];
    c[1] %= g.ncells[1];
    c[2] %= g.ncells[2];
}

// Same cell as cell_of(); returns false when the C remainder had to fold an index back
// into range (a coordinate that the float64 cell arithmetic puts outside the wrapped
// cell, e.g. on the upper face of a triclinic cell): such an atom lives in a cell that
// is not where it is in space, which the per-cell image prediction of rowtile.cuh
// cannot handle (the reference's per-pair minimum image can).
MDA_HD bool cell_of_checked(const GridGeom &g, float x, float y, float z, int *c)
{
    const double dx = (double)x - g.origin[0], dy = (double)y - g.origin[1], dz = (double)z - g.origin[2];
    c[2] = (int)(dz / g.cs[8]);
    c[1] = (int)((dy - dz * g.cs[7]) / g.cs[4]);
    c[0] = (int)((dx - dy * g.cs[3] - dz * g.cs[6]) / g.cs[0]);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (c[k] < 0 || c[k] >= g.ncells[k]) ok = false;
        c[k] %= g.ncells[k];
    }
    return ok;
}

MDA_HD int64_t cell_linear(const GridGeom &g, const int *c)
{
    return (int64_t)c[0] + (int64_t)c[1] * g.ncells[0] + (int64_t)c[2] * g.ncells[0] * g.ncells[1];
}

// neighbour cell with periodic wrap; -1 when outside (nsgrid.pyx:448-481)
MDA_HD int64_t cell_neighbor(const GridGeom &g, int cx, int cy, int cz)
{
    int c[3] = { cx, cy, cz };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (c[k] < 0) {
            if (!g.periodic[k]) return -1;
            c[k] = g.ncells[k] - 1;
        } else if (c[k] == g.ncells[k]) {
            if (!g.periodic[k]) return -1;
            c[k] = 0;
        }
    }
    return cell_linear(g, c);
}

// ---- QCP solve (qcprot.pyx:326-543), all double -------------------------------------
MDA_HD double qcp_solve(double *rot, const double *A, double E0, double N)
{
    const double Sxx = A[0], Sxy = A[1], Sxz = A[2];
    const double Syx = A[3], Syy = A[4], Syz = A[5];
    const double Szx = A[6], Szy = A[7], Szz = A[8];
    const double Sxx2 = Sxx * Sxx, Syy2 = Syy * Syy, Szz2 = Szz * Szz;
    const double Sxy2 = Sxy * Sxy, Syz2 = Syz * Syz, Sxz2 = Sxz * Sxz;
    const double Syx2 = Syx * Syx, Szy2 = Szy * Szy, Szx2 = Szx * Szx;

    const double t_a = 2.0 * (Syz * Szy - Syy * Szz);
    const double t_b = Syy2 + Szz2 - Sxx2 + Syz2 + Szy2;
    const double C2 = -2.0 * (Sxx2 + Syy2 + Szz2 + Sxy2 + Syx2 + Sxz2 + Szx2 + Syz2 + Szy2);
    const double C1 = 8.0 * (Sxx * Syz * Szy + Syy * Szx * Sxz + Szz * Sxy * Syx
                             - Sxx * Syy * Szz - Syz * Szx * Sxy - Szy * Syx * Sxz);
    const double xz_p = Sxz + Szx, yz_p = Syz + Szy, xy_p = Sxy + Syx;
    const double yz_m = Syz - Szy, xz_m = Sxz - Szx, xy_m = Sxy - Syx;
    const double xx_p_yy = Sxx + Syy, xx_m_yy = Sxx - Syy;
    const double t_c = Sxy2 + Sxz2 - Syx2 - Szx2;
    const double C0 = t_c * t_c
        + (t_b + t_a) * (t_b - t_a)
        + (-(xz_p) * (yz_m) + (xy_m) * (xx_m_yy - Szz)) * (-(xz_m) * (yz_p) + (xy_m) * (xx_m_yy + Szz))
        + (-(xz_p) * (yz_p) - (xy_p) * (xx_p_yy - Szz)) * (-(xz_m) * (yz_m) - (xy_p) * (xx_p_yy + Szz))
        + (+(xy_p) * (yz_p) + (xz_p) * (xx_m_yy + Szz)) * (-(xy_m) * (yz_m) + (xz_p) * (xx_p_yy + Szz))
        + (+(xy_p) * (yz_m) + (xz_m) * (xx_m_yy - Szz)) * (-(xy_m) * (yz_p) + (xz_m) * (xx_p_yy - Szz));

    double lam = E0;
    for (int it = 0; it < 50; ++it) {
        const double prev = lam;
        const double l2 = lam * lam;
        const double b = (l2 + C2) * lam;
        const double a = b + C1;
        lam -= (a * lam + C0) / (2.0 * l2 * lam + b + a);
        if (fabs(lam - prev) < fabs(1e-14 * lam)) break;
    }
    const double rms = sqrt(fabs(2.0 * (E0 - lam) / N));
    if (!rot) return rms;

    const double a11 = xx_p_yy + Szz - lam, a12 = yz_m, a13 = -xz_m, a14 = xy_m;
    const double a21 = yz_m, a22 = xx_m_yy - Szz - lam, a23 = xy_p, a24 = xz_p;
    const double a31 = a13, a32 = a23, a33 = Syy - Sxx - Szz - lam, a34 = yz_p;
    const double a41 = a14, a42 = a24, a43 = a34, a44 = Szz - xx_p_yy - lam;
    const double m3344 = a33 * a44 - a43 * a34, m3244 = a32 * a44 - a42 * a34;
    const double m3243 = a32 * a43 - a42 * a33, m3143 = a31 * a43 - a41 * a33;
    const double m3144 = a31 * a44 - a41 * a34, m3142 = a31 * a42 - a41 * a32;
    double q1 = a22 * m3344 - a23 * m3244 + a24 * m3243;
    double q2 = -a21 * m3344 + a23 * m3144 - a24 * m3143;
    double q3 = a21 * m3244 - a22 * m3144 + a24 * m3142;
    double q4 = -a21 * m3243 + a22 * m3143 - a23 * m3142;
    double qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
    const double tiny = 1e-6;
    if (qq < tiny) {
        q1 = a12 * m3344 - a13 * m3244 + a14 * m3243;
        q2 = -a11 * m3344 + a13 * m3144 - a14 * m3143;
        q3 = a11 * m3244 - a12 * m3144 + a14 * m3142;
        q4 = -a11 * m3243 + a12 * m3143 - a13 * m3142;
        qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
        if (qq < tiny) {
            const double m1324 = a13 * a24 - a14 * a23, m1224 = a12 * a24 - a14 * a22;
            const double m1223 = a12 * a23 - a13 * a22, m1124 = a11 * a24 - a14 * a21;
            const double m1123 = a11 * a23 - a13 * a21, m1122 = a11 * a22 - a12 * a21;
            q1 = a42 * m1324 - a43 * m1224 + a44 * m1223;
            q2 = -a41 * m1324 + a43 * m1124 - a44 * m1123;
            q3 = a41 * m1224 - a42 * m1124 + a44 * m1122;
            q4 = -a41 * m1223 + a42 * m1123 - a43 * m1122;
            qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
            if (qq < tiny) {
                q1 = a32 * m1324 - a33 * m1224 + a34 * m1223;
                q2 = -a31 * m1324 + a33 * m1124 - a34 * m1123;
                q3 = a31 * m1224 - a32 * m1124 + a34 * m1122;
                q4 = -a31 * m1223 + a32 * m1123 - a33 * m1122;
                qq = q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4;
                if (qq < tiny) {
                    for (int k = 0; k < 9; ++k) rot[k] = (k % 4 == 0) ? 1.0 : 0.0;
                    return rms;
                }
            }
        }
    }
    const double nq = sqrt(qq);
    q1 /= nq; q2 /= nq; q3 /= nq; q4 /= nq;
    const double aa = q1 * q1, xx = q2 * q2, yy = q3 * q3, zz = q4 * q4;
    const double xy = q2 * q3, az = q1 * q4, zx = q4 * q2, ay = q1 * q3, yz = q3 * q4, ax = q1 * q2;
    rot[0] = aa + xx - yy - zz; rot[1] = 2 * (xy + az);     rot[2] = 2 * (zx - ay);
    rot[3] = 2 * (xy - az);     rot[4] = aa - xx + yy - zz; rot[5] = 2 * (yz + ax);
    rot[6] = 2 * (zx + ay);     rot[7] = 2 * (yz - ax);     rot[8] = aa - xx - yy + zz;
    return rms;
}

// ---- histogram bin lookup with np.histogram's uniform-bin semantics ---------------
// (numpy/lib/_histograms_impl.py: index from the scaled value, then corrected
// against the actual float64 edges; last bin closed.)  `edges` has nbins+1
// entries; the caller has already established lo <= d <= hi.
MDA_HD int histogram_bin(double d, const double *edges, int nbins, double lo, double hi)
{
    double norm = (double)nbins / (hi - lo);
    double f = (d - lo) * norm;
    int k = (int)f;
    if (k >= nbins) k = nbins - 1;           // d == hi lands in the last bin
    if (d < edges[k]) --k;
    else if (d >= edges[k + 1] && k != nbins - 1) ++k;
    return k;
}

}  // namespace mdagpu
