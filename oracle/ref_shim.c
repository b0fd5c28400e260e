/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Exposes the reference's own header-only C kernels
 * (/root/reference/package/MDAnalysis/lib/include/calc_distances.h, found via
 * -I at build time; the header is NOT copied into this repository) through
 * extern "C"-style entry points, so that the unmodified reference arithmetic
 * can (a) pin oracle.c and (b) serve as the CPU baseline ("kind":"reference")
 * on the GPU box, where /root/reference does not exist but the prebuilt
 * oracle/_ref/*.so travels with the snapshot.
 *
 * The header uses uint64_t without including <stdint.h> (Cython normally
 * provides it), hence the include order below.
 */
#include <stdint.h>
#include <stddef.h>
#include "calc_distances.h"

#define REF_API __attribute__((visibility("default")))
typedef float (*xyz_t)[3];

REF_API int ref_used_openmp(void) { return USED_OPENMP; }

REF_API void ref_ortho_pbc(float *c, uint64_t n, float *box) { _ortho_pbc((xyz_t)c, n, box); }
REF_API void ref_triclinic_pbc(float *c, uint64_t n, float *box) { _triclinic_pbc((xyz_t)c, n, box); }

REF_API void ref_distance_array(float *r, uint64_t n, float *c, uint64_t m, int mode, float *box, double *out)
{
    if (mode == 0) _calc_distance_array((xyz_t)r, n, (xyz_t)c, m, out);
    else if (mode == 1) _calc_distance_array_ortho((xyz_t)r, n, (xyz_t)c, m, box, out);
    else _calc_distance_array_triclinic((xyz_t)r, n, (xyz_t)c, m, box, out);
}
REF_API void ref_self_distance_array(float *r, uint64_t n, int mode, float *box, double *out)
{
    if (mode == 0) _calc_self_distance_array((xyz_t)r, n, out);
    else if (mode == 1) _calc_self_distance_array_ortho((xyz_t)r, n, box, out);
    else _calc_self_distance_array_triclinic((xyz_t)r, n, box, out);
}
REF_API void ref_bonds(float *a, float *b, uint64_t n, int mode, float *box, double *out)
{
    if (mode == 0) _calc_bond_distance((xyz_t)a, (xyz_t)b, n, out);
    else if (mode == 1) _calc_bond_distance_ortho((xyz_t)a, (xyz_t)b, n, box, out);
    else _calc_bond_distance_triclinic((xyz_t)a, (xyz_t)b, n, box, out);
}
REF_API void ref_angles(float *a, float *b, float *c, uint64_t n, int mode, float *box, double *out)
{
    if (mode == 0) _calc_angle((xyz_t)a, (xyz_t)b, (xyz_t)c, n, out);
    else if (mode == 1) _calc_angle_ortho((xyz_t)a, (xyz_t)b, (xyz_t)c, n, box, out);
    else _calc_angle_triclinic((xyz_t)a, (xyz_t)b, (xyz_t)c, n, box, out);
}
REF_API void ref_dihedrals(float *a, float *b, float *c, float *d, uint64_t n, int mode, float *box, double *out)
{
    if (mode == 0) _calc_dihedral((xyz_t)a, (xyz_t)b, (xyz_t)c, (xyz_t)d, n, out);
    else if (mode == 1) _calc_dihedral_ortho((xyz_t)a, (xyz_t)b, (xyz_t)c, (xyz_t)d, n, box, out);
    else _calc_dihedral_triclinic((xyz_t)a, (xyz_t)b, (xyz_t)c, (xyz_t)d, n, box, out);
}
REF_API void ref_coord_transform(float *c, uint64_t n, double *m) { _coord_transform((xyz_t)c, n, m); }
REF_API void ref_minimum_image(double *x, float *box, float *inv) { minimum_image(x, box, inv); }
REF_API void ref_minimum_image_triclinic(double *x, float *box) { minimum_image_triclinic(x, box); }
REF_API void ref_minimum_image_ortho_lazy(double *x, float *box, float *half) { _minimum_image_ortho_lazy(x, box, half); }
