// host_check.cu -- TEST INFRASTRUCTURE ONLY (never part of libmdagpu.so).
//
// Compiles the __host__ __device__ arithmetic of mdanalysis_b200/csrc/geom.cuh
// for the CPU so that the very source the kernels execute can be compared with
// the oracle bit for bit on a machine without a GPU (-m "not gpu" tests).  It is
// a checker of device code, not a compute path: nothing in mdanalysis_b200
// loads this library.
#include <stdint.h>
#include <string.h>

#include "../mdanalysis_b200/csrc/geom.cuh"

using namespace mdagpu;
#define HC_API extern "C" __attribute__((visibility("default")))

HC_API void hc_wrap(float *xyz, uint64_t n, int mode, const float *box)
{
    BoxParams b = make_box_params(mode, box);
    for (uint64_t i = 0; i < n; ++i) wrap_any(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], b);
}

template <int MIC>
static void pair_d2_loop(const float *p, const float *q, uint64_t n, const BoxParams &b, double *out)
{
    for (uint64_t i = 0; i < n; ++i)
        out[i] = pair_d2_exact<MIC>(p[3 * i], p[3 * i + 1], p[3 * i + 2], q[3 * i], q[3 * i + 1],
                                    q[3 * i + 2], b);
}

// out[i] = exact squared minimum-image distance of p[i] - q[i] in flavour `mic`
HC_API void hc_pair_d2(int mic, const float *p, const float *q, uint64_t n, int box_mode,
                       const float *box, double *out)
{
    BoxParams b = make_box_params(box_mode, box);
    switch (mic) {
    case MIC_NONE: pair_d2_loop<MIC_NONE>(p, q, n, b, out); break;
    case MIC_ORTHO_ROUND: pair_d2_loop<MIC_ORTHO_ROUND>(p, q, n, b, out); break;
    case MIC_ORTHO_LAZY: pair_d2_loop<MIC_ORTHO_LAZY>(p, q, n, b, out); break;
    default: pair_d2_loop<MIC_TRICLINIC>(p, q, n, b, out); break;
    }
}

// kind 2: bonds, 3: angles, 4: dihedrals; same per-tuple code path as tuple_kernel
HC_API void hc_tuples(int kind, const float *a, const float *b_, const float *c, const float *d,
                      uint64_t n, int box_mode, const float *box, double *out)
{
    BoxParams box_p = make_box_params(box_mode, box);
    auto load = [&](const float *p, uint64_t i, float &x, float &y, float &z) {
        x = p[3 * i]; y = p[3 * i + 1]; z = p[3 * i + 2];
        if (box_p.mode == BOX_TRICLINIC) wrap_triclinic(x, y, z, box_p);
    };
    for (uint64_t i = 0; i < n; ++i) {
        float ax, ay, az, bx, by, bz, cx, cy, cz, dx, dy, dz;
        load(a, i, ax, ay, az);
        load(b_, i, bx, by, bz);
        if (kind == 2) {
            D3 v = pair_vec_exact(box_mode, ax, ay, az, bx, by, bz, box_p);
            out[i] = sqrt(norm2(v));
        } else if (kind == 3) {
            load(c, i, cx, cy, cz);
            D3 u = pair_vec_exact(box_mode, ax, ay, az, bx, by, bz, box_p);
            D3 v = pair_vec_exact(box_mode, cx, cy, cz, bx, by, bz, box_p);
            out[i] = angle_from(u, v);
        } else {
            load(c, i, cx, cy, cz);
            load(d, i, dx, dy, dz);
            D3 va = pair_vec_exact(box_mode, bx, by, bz, ax, ay, az, box_p);
            D3 vb = pair_vec_exact(box_mode, cx, cy, cz, bx, by, bz, box_p);
            D3 vc = pair_vec_exact(box_mode, dx, dy, dz, cx, cy, cz, box_p);
            out[i] = torsion(va, vb, vc);
        }
    }
}

HC_API void hc_cells(const float *xyz, uint64_t n, const int *ncells, const double *cs, int64_t *out)
{
    GridGeom g;
    memset(&g, 0, sizeof g);
    for (int k = 0; k < 3; ++k) g.ncells[k] = ncells[k];
    for (int k = 0; k < 9; ++k) g.cs[k] = cs[k];
    for (uint64_t i = 0; i < n; ++i) {
        int c[3];
        cell_of(g, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], c);
        out[i] = cell_linear(g, c);
    }
}

HC_API double hc_qcp_solve(double *rot, const double *A, double E0, double N)
{
    return qcp_solve(rot, A, E0, N);
}

HC_API int hc_histogram_bin(double d, const double *edges, int nbins)
{
    return histogram_bin(d, edges, nbins, edges[0], edges[nbins]);
}
