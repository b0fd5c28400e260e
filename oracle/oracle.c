/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * A plain-C, single-threaded CPU restatement of the arithmetic of the
 * MDAnalysis geometry hot path, written from the reference's published
 * behaviour (it is NOT a copy of the reference sources).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; mdanalysis_b200/ never does.
 *
 * Parity status: PINNED.  oracle/check_against_reference.py compares every
 * entry point bit-for-bit (distances, wraps, pair sets, d2) or to 1e-12 (QCP)
 * with the reference built in this container, and tests/test_oracle_golden.py
 * re-checks it against fixtures generated from that reference
 * (tests/golden/, made by tests/golden/make_golden.py) and the reference's
 * own known-answer values.
 *
 * Arithmetic rules taken from the reference (all file:line relative to
 * /root/reference/package/MDAnalysis/lib/):
 *   - coordinates are float32; every pair difference is a float32 subtraction
 *     that is then widened to double (include/calc_distances.h:390-392).
 *   - no FMA contraction (reference is built -O3 without -march/-ffast-math,
 *     package/setup.py:262-263); build this file with -ffp-contract=off.
 *
 * Build: see oracle/Makefile (gcc -std=c11 -O3 -ffp-contract=off -fsigned-zeros).
 */
#define _DEFAULT_SOURCE 1
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

enum { ORC_BOX_NONE = 0, ORC_BOX_ORTHO = 1, ORC_BOX_TRICLINIC = 2 };

typedef struct { double v[3]; } dvec3;

/* ------------------------------------------------------------------------- */
/* separation vectors and minimum images                                      */
/* ------------------------------------------------------------------------- */

/* float32 subtraction p - q, widened afterwards (calc_distances.h:390-392). */
static inline dvec3 sep_f32(const float *p, const float *q)
{
    dvec3 d;
    for (int k = 0; k < 3; ++k) {
        float t = p[k] - q[k];
        d.v[k] = (double)t;
    }
    return d;
}

/* Orthorhombic minimum image via scaled rounding (calc_distances.h:39-49):
 * s = inv*x ; x = L*(s - round(s)), inv being the *float32* reciprocal of L
 * (calc_distances.h:402-405); axes with L <= FLT_EPSILON are left alone. */
static inline void mic_ortho_round(dvec3 *d, const float *L, const float *invL)
{
    for (int k = 0; k < 3; ++k) {
        if (L[k] > FLT_EPSILON) {
            double s = (double)invL[k] * d->v[k];
            d->v[k] = (double)L[k] * (s - round(s));
        }
    }
}

/* "Lazy" orthorhombic minimum image used by the grid search
 * (calc_distances.h:51-70): one conditional +-L, valid for |x| < 1.5 L. */
static inline void mic_ortho_lazy(dvec3 *d, const float *L, const float *halfL)
{
    for (int k = 0; k < 3; ++k) {
        double x = d->v[k];
        if (x > (double)halfL[k])
            x -= (double)L[k];
        else if (x <= -(double)halfL[k])
            x += (double)L[k];
        d->v[k] = x;
    }
}

/* Triclinic minimum image: exhaustive search over the 27 neighbouring images,
 * x outermost, strict "<" so the first minimum wins, start value FLT_MAX
 * (calc_distances.h:72-114).  h is the row-major lower-triangular cell
 * (h[1]=h[2]=h[5]=0).  int*float products are float32 (exact for -1,0,1). */
static inline void mic_triclinic(dvec3 *d, const float *h)
{
    double best = (double)FLT_MAX;
    double bx = 0.0, by = 0.0, bz = 0.0;
    for (int a = -1; a <= 1; ++a) {
        double x_a = d->v[0] + (double)(h[0] * (float)a);
        for (int b = -1; b <= 1; ++b) {
            double x_ab = x_a + (double)(h[3] * (float)b);
            double y_b = d->v[1] + (double)(h[4] * (float)b);
            for (int c = -1; c <= 1; ++c) {
                double x = x_ab + (double)(h[6] * (float)c);
                double y = y_b + (double)(h[7] * (float)c);
                double z = d->v[2] + (double)(h[8] * (float)c);
                double r2 = x * x + y * y + z * z;
                if (r2 < best) {
                    best = r2;
                    bx = x; by = y; bz = z;
                }
            }
        }
    }
    d->v[0] = bx; d->v[1] = by; d->v[2] = bz;
}

static inline double norm2(const dvec3 *d)
{
    return (d->v[0] * d->v[0]) + (d->v[1] * d->v[1]) + (d->v[2] * d->v[2]);
}

/* ------------------------------------------------------------------------- */
/* wrapping coordinates into the primary cell                                 */
/* ------------------------------------------------------------------------- */

/* One component of the orthorhombic wrap (calc_distances.h:116-190).
 * Single +-L shifts are float32; the multi-shift estimate uses the double
 * reciprocal; lower bound is always tested before the upper bound. */
static inline float wrap_axis(float x, float L, double invL)
{
    float stored = x;   /* what the reference keeps in coords[i][j] */
    float probe = x;    /* the reference's scratch copy "crd"        */
    if (probe < 0.0f) {
        probe += L;
        if (probe < 0.0f) {
            int s = (int)floor((double)stored * invL);
            stored -= (float)s * L;
            if (stored < 0.0f)
                stored += L;
        } else {
            stored = probe;
        }
    }
    if (probe >= L) {
        probe -= L;
        if (probe >= L) {
            int s = (int)floor((double)stored * invL);
            stored -= (float)s * L;
            if (stored >= L)
                stored -= L;
        } else {
            stored = probe;
        }
    }
    return stored;
}

ORC_API void orc_wrap_ortho(float *xyz, uint64_t n, const float *L)
{
    if (L[0] == 0.0f && L[1] == 0.0f && L[2] == 0.0f)
        return;
    const double inv[3] = { 1.0 / (double)L[0], 1.0 / (double)L[1], 1.0 / (double)L[2] };
    for (uint64_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k)
            xyz[3 * i + k] = wrap_axis(xyz[3 * i + k], L[k], inv[k]);
}

/* Constants of the triclinic wrap (calc_distances.h:221-230). */
typedef struct {
    const float *h;
    double r0, r4, r8, r3, r6, r7;   /* entries of the inverse cell matrix */
    double ay, az, bz;               /* tilt factors for the single shifts  */
} tric_wrap_t;

static tric_wrap_t tric_wrap_setup(const float *h)
{
    tric_wrap_t w;
    w.h = h;
    w.r0 = 1.0 / (double)h[0];
    w.r4 = 1.0 / (double)h[4];
    w.r8 = 1.0 / (double)h[8];
    w.r3 = -h[3] * w.r0 * w.r4;
    w.r6 = (-w.r3 * h[7] - h[6] * w.r0) * w.r8;
    w.r7 = -h[7] * w.r4 * w.r8;
    w.ay = (double)h[3] * w.r4;
    w.az = (double)h[6] * w.r8;
    w.bz = (double)h[7] * w.r8;
    return w;
}

/* Single-image shifts along c, then b, then a (calc_distances.h:251-312 and
 * the re-check block :333-370).  Returns 1 as soon as one shift turned out
 * not to be enough; when `strict` is 0 the pass never bails out. */
static inline int tric_single_pass(float *c, const tric_wrap_t *w, int strict)
{
    const float *h = w->h;
    if (c[2] < 0.0f) {
        c[0] += h[6]; c[1] += h[7]; c[2] += h[8];
        if (strict && c[2] < 0.0f) return 1;
    }
    if (c[2] >= h[8]) {
        c[0] -= h[6]; c[1] -= h[7]; c[2] -= h[8];
        if (strict && c[2] >= h[8]) return 1;
    }
    double lo = c[2] * w->bz;
    double hi = lo + h[4];
    if (c[1] < lo) {
        c[0] += h[3]; c[1] += h[4];
        if (strict && c[1] < lo) return 1;
    }
    if (c[1] >= hi) {
        c[0] -= h[3]; c[1] -= h[4];
        if (strict && c[1] >= hi) return 1;
    }
    lo = c[1] * w->ay + c[2] * w->az;
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

static inline void wrap_tric_one(float *p, const tric_wrap_t *w)
{
    const float *h = w->h;
    float c[3] = { p[0], p[1], p[2] };
    /* NB: in the reference a failed shift along c does not stop it from
     * falling through "if (!msr)"; tric_single_pass returns at the same
     * points because nothing computed after a failure is used. */
    if (tric_single_pass(c, w, 1)) {
        /* multi-image shifts (calc_distances.h:317-329), float32 updates */
        int s = (int)floor(p[2] * w->r8);
        p[2] -= (float)s * h[8];
        p[1] -= (float)s * h[7];
        p[0] -= (float)s * h[6];
        s = (int)floor(p[1] * w->r4 + p[2] * w->r7);
        p[1] -= (float)s * h[4];
        p[0] -= (float)s * h[3];
        s = (int)floor(p[0] * w->r0 + p[1] * w->r3 + p[2] * w->r6);
        p[0] -= (float)s * h[0];
        c[0] = p[0]; c[1] = p[1]; c[2] = p[2];
        tric_single_pass(c, w, 0);
    }
    p[0] = c[0]; p[1] = c[1]; p[2] = c[2];
}

ORC_API void orc_wrap_triclinic(float *xyz, uint64_t n, const float *h)
{
    if (h[0] == 0.0f && h[4] == 0.0f && h[8] == 0.0f)
        return;
    tric_wrap_t w = tric_wrap_setup(h);
    for (uint64_t i = 0; i < n; ++i)
        wrap_tric_one(xyz + 3 * i, &w);
}

/* ------------------------------------------------------------------------- */
/* box context shared by the pairwise routines                                */
/* ------------------------------------------------------------------------- */
typedef struct {
    int mode;
    const float *box;     /* float[3] (ortho) or float[9] (triclinic) */
    float inv[3];
} boxctx;

static boxctx boxctx_make(int mode, const float *box)
{
    boxctx b;
    b.mode = mode;
    b.box = box;
    b.inv[0] = b.inv[1] = b.inv[2] = 0.0f;
    if (mode == ORC_BOX_ORTHO)
        for (int k = 0; k < 3; ++k)
            b.inv[k] = (float)(1.0 / box[k]);
    return b;
}

static inline void mic_apply(dvec3 *d, const boxctx *b)
{
    if (b->mode == ORC_BOX_ORTHO)
        mic_ortho_round(d, b->box, b->inv);
    else if (b->mode == ORC_BOX_TRICLINIC)
        mic_triclinic(d, b->box);
}

/* wrapped private copy for the triclinic variants (calc_distances.h:429-430) */
static float *dup_wrapped(const float *xyz, uint64_t n, const boxctx *b)
{
    float *w = (float *)malloc((n ? n : 1) * 3 * sizeof(float));
    memcpy(w, xyz, n * 3 * sizeof(float));
    if (b->mode == ORC_BOX_TRICLINIC)
        orc_wrap_triclinic(w, n, b->box);
    return w;
}

/* ------------------------------------------------------------------------- */
/* distance_array / self_distance_array (calc_distances.h:381-531)            */
/* ------------------------------------------------------------------------- */
ORC_API void orc_distance_array(const float *ref, uint64_t n, const float *conf, uint64_t m,
                                int mode, const float *box, double *out)
{
    boxctx b = boxctx_make(mode, box);
    const float *r = ref, *c = conf;
    float *rw = NULL, *cw = NULL;
    if (mode == ORC_BOX_TRICLINIC) {
        r = rw = dup_wrapped(ref, n, &b);
        c = cw = dup_wrapped(conf, m, &b);
    }
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t j = 0; j < m; ++j) {
            dvec3 d = sep_f32(c + 3 * j, r + 3 * i);
            mic_apply(&d, &b);
            out[i * m + j] = sqrt(norm2(&d));
        }
    free(rw); free(cw);
}

ORC_API void orc_self_distance_array(const float *ref, uint64_t n, int mode, const float *box,
                                     double *out)
{
    boxctx b = boxctx_make(mode, box);
    const float *r = ref;
    float *rw = NULL;
    if (mode == ORC_BOX_TRICLINIC)
        r = rw = dup_wrapped(ref, n, &b);
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; ++i)
        for (uint64_t j = i + 1; j < n; ++j) {
            dvec3 d = sep_f32(r + 3 * j, r + 3 * i);
            mic_apply(&d, &b);
            out[k++] = sqrt(norm2(&d));
        }
    free(rw);
}

/* ------------------------------------------------------------------------- */
/* bonds, angles, dihedrals (calc_distances.h:557-860)                        */
/* ------------------------------------------------------------------------- */
ORC_API void orc_bonds(const float *a, const float *b_, uint64_t n, int mode, const float *box,
                       double *out)
{
    boxctx b = boxctx_make(mode, box);
    float *aw = dup_wrapped(a, n, &b), *bw = dup_wrapped(b_, n, &b);
    for (uint64_t i = 0; i < n; ++i) {
        dvec3 d = sep_f32(aw + 3 * i, bw + 3 * i);
        mic_apply(&d, &b);
        out[i] = sqrt(norm2(&d));
    }
    free(aw); free(bw);
}

ORC_API void orc_angles(const float *a, const float *b_, const float *c, uint64_t n, int mode,
                        const float *box, double *out)
{
    boxctx b = boxctx_make(mode, box);
    float *aw = dup_wrapped(a, n, &b), *bw = dup_wrapped(b_, n, &b), *cw = dup_wrapped(c, n, &b);
    for (uint64_t i = 0; i < n; ++i) {
        dvec3 u = sep_f32(aw + 3 * i, bw + 3 * i);   /* apex -> first  */
        mic_apply(&u, &b);
        dvec3 v = sep_f32(cw + 3 * i, bw + 3 * i);   /* apex -> third  */
        mic_apply(&v, &b);
        double dot = u.v[0] * v.v[0] + u.v[1] * v.v[1] + u.v[2] * v.v[2];
        double cx = u.v[1] * v.v[2] - u.v[2] * v.v[1];
        double cy = -u.v[0] * v.v[2] + u.v[2] * v.v[0];
        double cz = u.v[0] * v.v[1] - u.v[1] * v.v[0];
        double cross = sqrt(cx * cx + cy * cy + cz * cz);
        out[i] = atan2(cross, dot);
    }
    free(aw); free(bw); free(cw);
}

/* signed torsion from three bond vectors (calc_distances.h:726-762) */
static double torsion(const dvec3 *p, const dvec3 *q, const dvec3 *r)
{
    const double *va = p->v, *vb = q->v, *vc = r->v;
    double n1x = -va[1] * vb[2] + va[2] * vb[1];
    double n1y = va[0] * vb[2] - va[2] * vb[0];
    double n1z = -va[0] * vb[1] + va[1] * vb[0];
    double n2x = -vb[1] * vc[2] + vb[2] * vc[1];
    double n2y = vb[0] * vc[2] - vb[2] * vc[0];
    double n2z = -vb[0] * vc[1] + vb[1] * vc[0];
    double x = (n1x * n2x + n1y * n2y + n1z * n2z);
    double mx = n1y * n2z - n1z * n2y;
    double my = -n1x * n2z + n1z * n2x;
    double mz = n1x * n2y - n1y * n2x;
    double len_b = sqrt(vb[0] * vb[0] + vb[1] * vb[1] + vb[2] * vb[2]);
    double y = (mx * vb[0] + my * vb[1] + mz * vb[2]) / len_b;
    if (fabs(x) == 0.0 && fabs(y) == 0.0)
        return NAN;
    return atan2(y, x);
}

ORC_API void orc_dihedrals(const float *a, const float *b_, const float *c, const float *d_,
                           uint64_t n, int mode, const float *box, double *out)
{
    boxctx b = boxctx_make(mode, box);
    float *aw = dup_wrapped(a, n, &b), *bw = dup_wrapped(b_, n, &b);
    float *cw = dup_wrapped(c, n, &b), *dw = dup_wrapped(d_, n, &b);
    for (uint64_t i = 0; i < n; ++i) {
        dvec3 p = sep_f32(bw + 3 * i, aw + 3 * i);
        mic_apply(&p, &b);
        dvec3 q = sep_f32(cw + 3 * i, bw + 3 * i);
        mic_apply(&q, &b);
        dvec3 r = sep_f32(dw + 3 * i, cw + 3 * i);
        mic_apply(&r, &b);
        out[i] = torsion(&p, &q, &r);
    }
    free(aw); free(bw); free(cw); free(dw);
}

/* coords <- coords . M with a float32 accumulator (calc_distances.h:533-555) */
ORC_API void orc_coord_transform(float *xyz, uint64_t n, const double *M)
{
    for (uint64_t i = 0; i < n; ++i) {
        float *p = xyz + 3 * i;
        float q[3];
        for (int j = 0; j < 3; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 3; ++k)
                acc = (float)((double)acc + (double)p[k] * M[3 * k + j]);
            q[j] = acc;
        }
        p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
    }
}

/* ------------------------------------------------------------------------- */
/* FastNS cell-list neighbour search (nsgrid.pyx:226-644)                     */
/* ------------------------------------------------------------------------- */
typedef struct {
    double cutoff;
    int triclinic, pbc;
    int periodic[3];
    int ncells[3];
    double cs[9];            /* cell widths on the diagonal, tilts off it */
    float dims[6], half[3], h[9];
    float max_cutoff;
    uint64_t n;
    float *xyz;              /* wrapped copy */
    int32_t *head, *next;
} orc_grid;

typedef struct {
    int32_t *pairs;          /* 2*count */
    double *d2;
    uint64_t count, cap;
} orc_pairs;

static void pairs_push(orc_pairs *r, int32_t i, int32_t j, double d2)
{
    if (r->count == r->cap) {
        r->cap = r->cap ? r->cap * 2 : 1024;
        r->pairs = (int32_t *)realloc(r->pairs, r->cap * 2 * sizeof(int32_t));
        r->d2 = (double *)realloc(r->d2, r->cap * sizeof(double));
    }
    r->pairs[2 * r->count] = i;
    r->pairs[2 * r->count + 1] = j;
    r->d2[r->count] = d2;
    r->count++;
}

static float sin_deg_f32(float deg)
{
    /* nsgrid.pyx:119-122: float32 angle, double libm sine, float32 result */
    deg = (float)((double)deg * (M_PI / 180.));
    return (float)sin((double)deg);
}

static float fmin_f32(float a, float b) { return a < b ? a : b; }
static float fmax_f32(float a, float b) { return a > b ? a : b; }

/* grid geometry (nsgrid.pyx:294-387).  box6 = [lx,ly,lz,alpha,beta,gamma]
 * float32; h9 = float32 triclinic_vectors(box6) supplied by the caller
 * (mdamath.py:309-405 is host-side NumPy). */
ORC_API void orc_grid_geometry(double cutoff_in, const float *box6, const float *h9, int pbc,
                               int *ncells, double *cellsize, int *periodic, int *triclinic,
                               float *max_cutoff_out)
{
    float cutoff = (float)(cutoff_in > 1.0 ? cutoff_in : 1.0);
    *triclinic = (h9[3] != 0.0f || h9[6] != 0.0f || h9[7] != 0.0f);

    float mc = h9[4] * sin_deg_f32(box6[3]);
    mc = fmin_f32(mc, h9[8] * sin_deg_f32(box6[3]));
    mc = fmin_f32(mc, h9[0] * sin_deg_f32(box6[4]));
    mc = fmin_f32(mc, h9[8] * sin_deg_f32(box6[4]));
    mc = fmin_f32(mc, h9[0] * sin_deg_f32(box6[5]));
    mc = fmin_f32(mc, h9[4] * sin_deg_f32(box6[5]));
    mc /= 2;
    *max_cutoff_out = mc;

    float min_cell = cutoff;
    if (*triclinic)
        for (int k = 3; k < 6; ++k)
            min_cell = fmax_f32(cutoff / sin_deg_f32(box6[k]), min_cell);
    min_cell = (float)((double)min_cell + 0.001);

    const float edge[3] = { h9[0], h9[4], h9[8] };
    for (int k = 0; k < 3; ++k) {
        double nc = floor((double)(edge[k] / min_cell));
        ncells[k] = (int)(nc < 1290.0 ? nc : 1290.0);
    }
    for (int k = 0; k < 3; ++k) {
        if (pbc) {
            if (ncells[k] <= 3) { ncells[k] = 1; periodic[k] = 0; }
            else periodic[k] = 1;
        } else {
            if (ncells[k] <= 2) ncells[k] = 1;
            periodic[k] = 0;
        }
    }
    memset(cellsize, 0, 9 * sizeof(double));
    cellsize[0] = (double)h9[0] / (double)ncells[0];
    cellsize[3] = (double)(h9[3] / h9[4]);
    cellsize[4] = (double)h9[4] / (double)ncells[1];
    cellsize[6] = (double)(h9[6] / h9[8]);
    cellsize[7] = (double)(h9[7] / h9[8]);
    cellsize[8] = (double)h9[8] / (double)ncells[2];
}

/* cell coordinates of a wrapped position (nsgrid.pyx:435-446); C remainder */
static inline void cell_of(const orc_grid *g, const float *p, int *c)
{
    c[2] = (int)((double)p[2] / g->cs[8]);
    c[1] = (int)(((double)p[1] - (double)p[2] * g->cs[7]) / g->cs[4]);
    c[0] = (int)(((double)p[0] - (double)p[1] * g->cs[3] - (double)p[2] * g->cs[6]) / g->cs[0]);
    c[0] %= g->ncells[0];
    c[1] %= g->ncells[1];
    c[2] %= g->ncells[2];
}

/* neighbour cell id with periodic wrap, -1 when outside (nsgrid.pyx:448-481) */
static inline int cell_id(const orc_grid *g, int cx, int cy, int cz)
{
    int c[3] = { cx, cy, cz };
    for (int k = 0; k < 3; ++k) {
        if (c[k] < 0) {
            if (!g->periodic[k]) return -1;
            c[k] = g->ncells[k] - 1;
        } else if (c[k] == g->ncells[k]) {
            if (!g->periodic[k]) return -1;
            c[k] = 0;
        }
    }
    return c[0] + c[1] * g->ncells[0] + c[2] * g->ncells[0] * g->ncells[1];
}

static inline double grid_d2(const orc_grid *g, const float *a, const float *b)
{
    dvec3 d = sep_f32(a, b);
    if (g->pbc) {
        if (g->triclinic) mic_triclinic(&d, g->h);
        else mic_ortho_lazy(&d, g->dims, g->half);
    }
    return d.v[0] * d.v[0] + d.v[1] * d.v[1] + d.v[2] * d.v[2];
}

ORC_API orc_grid *orc_grid_create(double cutoff, const float *xyz, uint64_t n, const float *box6,
                                  const float *h9, int pbc)
{
    orc_grid *g = (orc_grid *)calloc(1, sizeof(orc_grid));
    g->cutoff = cutoff;
    g->pbc = pbc;
    g->n = n;
    memcpy(g->dims, box6, 6 * sizeof(float));
    memcpy(g->h, h9, 9 * sizeof(float));
    for (int k = 0; k < 3; ++k)
        g->half[k] = (float)(0.5 * (double)box6[k]);
    orc_grid_geometry(cutoff, box6, h9, pbc, g->ncells, g->cs, g->periodic, &g->triclinic,
                      &g->max_cutoff);
    uint64_t ncell = (uint64_t)g->ncells[0] * g->ncells[1] * g->ncells[2];
    g->head = (int32_t *)malloc(ncell * sizeof(int32_t));
    g->next = (int32_t *)malloc((n ? n : 1) * sizeof(int32_t));
    for (uint64_t c = 0; c < ncell; ++c) g->head[c] = -1;
    g->xyz = (float *)malloc((n ? n : 1) * 3 * sizeof(float));
    memcpy(g->xyz, xyz, n * 3 * sizeof(float));
    if (g->triclinic) orc_wrap_triclinic(g->xyz, n, g->h);
    else orc_wrap_ortho(g->xyz, n, g->dims);
    for (uint64_t i = 0; i < n; ++i) {
        int c[3];
        cell_of(g, g->xyz + 3 * i, c);
        int id = c[0] + c[1] * g->ncells[0] + c[2] * g->ncells[0] * g->ncells[1];
        g->next[i] = g->head[id];
        g->head[id] = (int32_t)i;
    }
    return g;
}

ORC_API float orc_grid_max_cutoff(const orc_grid *g) { return g->max_cutoff; }
ORC_API void orc_grid_info(const orc_grid *g, int *ncells, double *cellsize, int *periodic,
                           int *triclinic)
{
    memcpy(ncells, g->ncells, sizeof g->ncells);
    memcpy(cellsize, g->cs, sizeof g->cs);
    memcpy(periodic, g->periodic, sizeof g->periodic);
    *triclinic = g->triclinic;
}

ORC_API void orc_grid_free(orc_grid *g)
{
    if (!g) return;
    free(g->head); free(g->next); free(g->xyz); free(g);
}

ORC_API orc_pairs *orc_pairs_new(void) { return (orc_pairs *)calloc(1, sizeof(orc_pairs)); }
ORC_API uint64_t orc_pairs_count(const orc_pairs *r) { return r->count; }
ORC_API void orc_pairs_export(const orc_pairs *r, int64_t *pairs, double *d2)
{
    for (uint64_t k = 0; k < 2 * r->count; ++k) pairs[k] = r->pairs[k];
    memcpy(d2, r->d2, r->count * sizeof(double));
}
ORC_API void orc_pairs_free(orc_pairs *r)
{
    if (!r) return;
    free(r->pairs); free(r->d2); free(r);
}

/* query set against the grid, 27-cell stencil (nsgrid.pyx:499-580) */
ORC_API void orc_grid_search(const orc_grid *g, const float *q, uint64_t nq, orc_pairs *out)
{
    const double rc2 = g->cutoff * g->cutoff;
    for (uint64_t i = 0; i < nq; ++i) {
        float p[3] = { q[3 * i], q[3 * i + 1], q[3 * i + 2] };
        if (g->triclinic) orc_wrap_triclinic(p, 1, g->h);
        else orc_wrap_ortho(p, 1, g->dims);
        int c[3];
        cell_of(g, p, c);
        for (int ox = -1; ox <= 1; ++ox)
            for (int oy = -1; oy <= 1; ++oy)
                for (int oz = -1; oz <= 1; ++oz) {
                    int id = cell_id(g, c[0] + ox, c[1] + oy, c[2] + oz);
                    if (id < 0) continue;
                    for (int32_t j = g->head[id]; j != -1; j = g->next[j]) {
                        double d2 = grid_d2(g, p, g->xyz + 3 * (uint64_t)j);
                        if (d2 <= rc2) pairs_push(out, (int32_t)i, j, d2);
                    }
                }
    }
}

/* all pairs inside the grid: same cell + 13-cell half shell (nsgrid.pyx:582-644) */
ORC_API void orc_grid_self_search(const orc_grid *g, orc_pairs *out)
{
    static const int shell[13][3] = {
        { 1, 0, 0 }, { 1, 1, 0 }, { 0, 1, 0 }, { -1, 1, 0 }, { 1, 0, -1 }, { 1, 1, -1 },
        { 0, 1, -1 }, { -1, 1, -1 }, { 1, 0, 1 }, { 1, 1, 1 }, { 0, 1, 1 }, { -1, 1, 1 },
        { 0, 0, 1 } };
    const double rc2 = g->cutoff * g->cutoff;
    for (int cx = 0; cx < g->ncells[0]; ++cx)
        for (int cy = 0; cy < g->ncells[1]; ++cy)
            for (int cz = 0; cz < g->ncells[2]; ++cz) {
                int ci = cell_id(g, cx, cy, cz);
                for (int32_t i = g->head[ci]; i != -1; i = g->next[i]) {
                    const float *pi = g->xyz + 3 * (uint64_t)i;
                    for (int32_t j = g->next[i]; j != -1; j = g->next[j]) {
                        double d2 = grid_d2(g, pi, g->xyz + 3 * (uint64_t)j);
                        if (d2 <= rc2) pairs_push(out, i, j, d2);
                    }
                    for (int s = 0; s < 13; ++s) {
                        int cj = cell_id(g, cx + shell[s][0], cy + shell[s][1], cz + shell[s][2]);
                        if (cj < 0) continue;
                        for (int32_t j = g->head[cj]; j != -1; j = g->next[j]) {
                            double d2 = grid_d2(g, pi, g->xyz + 3 * (uint64_t)j);
                            if (d2 <= rc2) pairs_push(out, i, j, d2);
                        }
                    }
                }
            }
}

/* ------------------------------------------------------------------------- */
/* QCP (qcprot.pyx:179-543)                                                   */
/* ------------------------------------------------------------------------- */

/* A[3a+b] = sum_i w_i c1[i][a] c2[i][b]; returns E0 = (G1+G2)/2.  Sequential
 * accumulation in the input precision (qcprot.pyx:179-277). */
#define ORC_INNER(NAME, T)                                                                  \
    ORC_API T NAME(T *A, const T *c1, const T *c2, int N, const T *w)                       \
    {                                                                                       \
        T G1 = 0, G2 = 0;                                                                   \
        for (int k = 0; k < 9; ++k) A[k] = 0;                                               \
        for (int i = 0; i < N; ++i) {                                                       \
            T x1, y1, z1, x2 = c2[3 * i], y2 = c2[3 * i + 1], z2 = c2[3 * i + 2];           \
            if (w) {                                                                        \
                x1 = w[i] * c1[3 * i]; y1 = w[i] * c1[3 * i + 1]; z1 = w[i] * c1[3 * i + 2];\
                G1 += x1 * c1[3 * i] + y1 * c1[3 * i + 1] + z1 * c1[3 * i + 2];             \
                G2 += w[i] * (x2 * x2 + y2 * y2 + z2 * z2);                                 \
            } else {                                                                        \
                x1 = c1[3 * i]; y1 = c1[3 * i + 1]; z1 = c1[3 * i + 2];                     \
                G1 += (x1 * x1 + y1 * y1 + z1 * z1);                                        \
                G2 += (x2 * x2 + y2 * y2 + z2 * z2);                                        \
            }                                                                               \
            A[0] += (x1 * x2); A[1] += (x1 * y2); A[2] += (x1 * z2);                        \
            A[3] += (y1 * x2); A[4] += (y1 * y2); A[5] += (y1 * z2);                        \
            A[6] += (z1 * x2); A[7] += (z1 * y2); A[8] += (z1 * z2);                        \
        }                                                                                   \
        return (G1 + G2) * (T)0.5;                                                          \
    }
ORC_INNER(orc_inner_product_f64, double)
ORC_INNER(orc_inner_product_f32, float)

/* largest root of the QCP quartic by Newton iteration, then (optionally) the
 * rotation from a column of the adjugate (qcprot.pyx:326-543).  rot may be
 * NULL. All arithmetic double. */
ORC_API double orc_qcp_solve(double *rot, const double *A, double E0, int N)
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

/* CalcRMSDRotationalMatrix(ref, conf, N, rot, w): inner product is taken with
 * (conf, ref) in that order (qcprot.pyx:319). */
ORC_API double orc_qcp_f64(const double *ref, const double *conf, int N, double *rot,
                           const double *w)
{
    double A[9];
    double E0 = orc_inner_product_f64(A, conf, ref, N, w);
    return orc_qcp_solve(rot, A, E0, N);
}

ORC_API double orc_qcp_f32(const float *ref, const float *conf, int N, float *rot, const float *w)
{
    float A[9];
    float E0 = orc_inner_product_f32(A, conf, ref, N, w);
    double Ad[9], R[9];
    for (int k = 0; k < 9; ++k) Ad[k] = (double)A[k];
    double rms = orc_qcp_solve(rot ? R : NULL, Ad, (double)E0, N);
    if (rot)
        for (int k = 0; k < 9; ++k) rot[k] = (float)R[k];
    return rms;
}
