/*
 * mdagpu.h -- C ABI of libmdagpu.so, the B200 (sm_100a) backend for the
 * MDAnalysis geometry hot path.
 *
 * Every entry point is what a reference-side FFI (a Cython `cdef extern`, or
 * ctypes as in INTEGRATION.md) would bind in place of the routine cited next to
 * it.  Paths are relative to /root/reference/package/MDAnalysis/.
 *
 * Conventions
 *   - coordinates: float32, C-contiguous (n,3)           (lib/include/calc_distances.h:30)
 *   - box for the distance kernels: box_mode + `box`
 *       MDAGPU_BOX_NONE      box ignored (may be NULL)
 *       MDAGPU_BOX_ORTHO     float32[3]  edge lengths     (lib/util.py:2522-2523)
 *       MDAGPU_BOX_TRICLINIC float32[9]  row-major lower-triangular cell matrix
 *                                                         (lib/util.py:2524, lib/mdamath.py:309-405)
 *   - results: float64; indices: int64
 *   - `space`: MDAGPU_HOST   -> all array arguments are host pointers; the call
 *                               stages them through device memory and returns
 *                               after the results are back on the host.
 *              MDAGPU_DEVICE -> all array arguments are device pointers on the
 *                               current device; the call only enqueues work on
 *                               `stream` (small scalar parameters such as `box`
 *                               are always host pointers).
 *   - `stream`: a cudaStream_t (NULL = legacy default stream).
 *   - return value: MDAGPU_OK or a negative MDAGPU_ERR_*; the message is
 *     available from mdagpu_last_error() (thread-local).
 *   - the library keeps no reference to caller memory after a call returns
 *     (device-space calls: after the enqueued work has completed).
 */
#ifndef MDAGPU_H
#define MDAGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDAGPU_ABI_VERSION 1

#define MDAGPU_OK 0
#define MDAGPU_ERR_CUDA (-1)
#define MDAGPU_ERR_INVALID (-2)
#define MDAGPU_ERR_NOMEM (-3)

#define MDAGPU_HOST 0
#define MDAGPU_DEVICE 1

#define MDAGPU_BOX_NONE 0
#define MDAGPU_BOX_ORTHO 1
#define MDAGPU_BOX_TRICLINIC 2

/* minimum-image flavour of the pair engine (capped distances / RDF / contacts) */
#define MDAGPU_MIC_NONE 0        /* no box                                              */
#define MDAGPU_MIC_ORTHO_ROUND 1 /* lib.distances ortho   (calc_distances.h:39-49)      */
#define MDAGPU_MIC_ORTHO_LAZY 2  /* FastNS ortho          (calc_distances.h:51-70)      */
#define MDAGPU_MIC_TRICLINIC 3   /* 27-image search       (calc_distances.h:72-114)     */

typedef void *mdagpu_stream_t;

#if defined(__GNUC__)
#define MDAGPU_API __attribute__((visibility("default")))
#else
#define MDAGPU_API
#endif

/* ---- library ---------------------------------------------------------------- */
MDAGPU_API int mdagpu_abi_version(void);
MDAGPU_API const char *mdagpu_last_error(void);
MDAGPU_API int mdagpu_device_count(int *count);
MDAGPU_API int mdagpu_set_device(int device);
/* number of kernel launches issued by this library in this process so far */
MDAGPU_API unsigned long long mdagpu_launch_count(void);

/* Copies between caller-owned host memory and device memory in stream order (csrc/hostcopy.cu):
 * pageable host memory is staged through a pinned ring by a pool of host threads (about 40 GB/s
 * against the driver's 10-20 GB/s), pinned memory goes straight to the DMA engine.
 * host_to_device returns once `host` has been read; device_to_host once `host` holds the data. */
MDAGPU_API int mdagpu_host_to_device(void *dev, const void *host, uint64_t bytes, mdagpu_stream_t stream);
MDAGPU_API int mdagpu_device_to_host(void *host, const void *dev, uint64_t bytes, mdagpu_stream_t stream);
/* Per-kernel device timing for benchmarks: when enabled, CUDA events bracket every
 * launch on its stream; mdagpu_profile_report() synchronises, writes
 * "name<TAB>launches<TAB>total_ms" lines into buf and clears the records. */
MDAGPU_API int mdagpu_profile_enable(int on);
MDAGPU_API int mdagpu_profile_report(char *buf, uint64_t size);

/* ---- lib.distances kernels ---------------------------------------------------- */

/* _calc_distance_array{,_ortho,_triclinic}  lib/include/calc_distances.h:381-446
 * (shims lib/c_distances.pyx:48-78).  out[i*m+j] = |conf[j]-ref[i]|_mic.
 * `nframes` > 1 evaluates a batch: ref[f] (n,3), conf[f] (m,3), out[f] (n,m),
 * one box for all frames. */
MDAGPU_API int mdagpu_distance_array(const float *ref, uint64_t n, const float *conf, uint64_t m,
                          int box_mode, const float *box, double *out, uint64_t nframes,
                          int space, mdagpu_stream_t stream);

/* _calc_self_distance_array{,_ortho,_triclinic}  calc_distances.h:448-531
 * (shims c_distances.pyx:81-105).  Packed upper triangle,
 * out[i*(2n-i-1)/2 + (j-i-1)] for j > i; nframes as above. */
MDAGPU_API int mdagpu_self_distance_array(const float *ref, uint64_t n, int box_mode, const float *box,
                               double *out, uint64_t nframes, int space,
                               mdagpu_stream_t stream);

/* _calc_bond_distance*  calc_distances.h:557-616  (c_distances.pyx:121-145) */
MDAGPU_API int mdagpu_calc_bonds(const float *a, const float *b, uint64_t n, int box_mode, const float *box,
                      double *out, int space, mdagpu_stream_t stream);
/* _calc_angle*          calc_distances.h:618-724  (c_distances.pyx:148-175) */
MDAGPU_API int mdagpu_calc_angles(const float *a, const float *b, const float *c, uint64_t n, int box_mode,
                       const float *box, double *out, int space, mdagpu_stream_t stream);
/* _calc_dihedral*       calc_distances.h:726-860  (c_distances.pyx:178-208) */
MDAGPU_API int mdagpu_calc_dihedrals(const float *a, const float *b, const float *c, const float *d,
                          uint64_t n, int box_mode, const float *box, double *out, int space,
                          mdagpu_stream_t stream);
/* _coord_transform      calc_distances.h:533-555  (c_distances.pyx:108-118); in place,
 * `matrix` = 9 doubles on the host. */
MDAGPU_API int mdagpu_coord_transform(float *coords, uint64_t n, const double *matrix, int space,
                           mdagpu_stream_t stream);
/* _ortho_pbc / _triclinic_pbc  calc_distances.h:116-379  (c_distances.pyx:211-222); in place */
MDAGPU_API int mdagpu_wrap(float *coords, uint64_t n, int box_mode, const float *box, int space,
                mdagpu_stream_t stream);

/* minimize_vectors  lib/distances.py:1790-1830 -> _minimize_vectors_{ortho,triclinic}
 * (c_distances.pyx:372-414).  vectors/out: [n,3] float32 (is_double = 0) or float64; box: 3
 * (BOX_ORTHO) or 9 (BOX_TRICLINIC, flattened triclinic_vectors) values of the SAME dtype,
 * host memory.  Same dtype out as in, as the reference. */
MDAGPU_API int mdagpu_minimize_vectors(const void *vectors, uint64_t n, int is_double, int box_mode,
                            const void *box, void *out, int space, mdagpu_stream_t stream);

/* Contacts._single_frame  analysis/contacts.py:486-499: r = distance_array(A, B, box)[mask].
 * Evaluates only the listed elements (idx_i into A, idx_j into B) of every frame's distance
 * matrix: out[f, p] = |B[f, idx_j[p]] - A[f, idx_i[p]]| with distance_array's arithmetic.
 * A [nframes, na, 3], B [nframes, nb, 3] float32; boxes: 3 (ortho) or 9 (triclinic) float32
 * per frame (boxes_per_frame) or once, host memory; out float64 [nframes, npairs]. */
MDAGPU_API int mdagpu_pair_list_distances(const float *A, uint64_t na, const float *B, uint64_t nb,
                               uint64_t nframes, const long long *idx_i, const long long *idx_j,
                               uint64_t npairs, int box_mode, const float *boxes, int boxes_per_frame,
                               double *out, int space, mdagpu_stream_t stream);

/* ---- FastNS: cell-list neighbour search (lib/nsgrid.pyx) ----------------------- */
typedef struct mdagpu_grid mdagpu_grid;
typedef struct mdagpu_pairs mdagpu_pairs;

/* FastNS.__init__ -> _prepare_box + _pack_grid   nsgrid.pyx:226-419.
 * box6 = [lx,ly,lz,alpha,beta,gamma] float32, h9 = float32 triclinic_vectors(box6)
 * (both host).  The grid geometry (ncells, cellsize, periodic flags) is the
 * reference's, so the 27/13-cell stencils visit exactly the reference's
 * candidates.  Returns MDAGPU_ERR_INVALID (message as the reference's
 * ValueError) when cutoff > max_cutoff. */
MDAGPU_API int mdagpu_grid_create(double cutoff, const float *coords, uint64_t n, const float *box6,
                       const float *h9, int pbc, int space, mdagpu_stream_t stream,
                       mdagpu_grid **out);
MDAGPU_API int mdagpu_grid_info(const mdagpu_grid *g, int *ncells3, double *cellsize9, int *periodic3,
                     int *triclinic, float *max_cutoff);
/* geometry only (no coordinates), for callers that validate a cutoff */
MDAGPU_API int mdagpu_grid_geometry(double cutoff, const float *box6, const float *h9, int pbc, int *ncells3,
                         double *cellsize9, int *periodic3, int *triclinic, float *max_cutoff);
/* Host only, no GPU work.  Is the (2W + 1)^3 stencil COMPLETE on the grid FastNS builds for this box
 * and cutoff, i.e. does FastNS (W = 1) return every pair within the cutoff?  (nsgrid.pyx:341-346
 * elongates the cells by cutoff / sin(angle), which does not always cover the shear of its own cell
 * coordinates, nsgrid.pyx:435-446.)  *complete = 0 or 1.  The triclinic InterRDF path leaves the
 * reference's grid only for cells where this holds. */
MDAGPU_API int mdagpu_grid_stencil_complete(double cutoff, const float *box6, const float *h9, int W, int *complete);
MDAGPU_API void mdagpu_grid_free(mdagpu_grid *g);

/* FastNS.search        nsgrid.pyx:499-580 : (query i, grid atom j) with d2 <= cutoff^2 */
MDAGPU_API int mdagpu_grid_search(mdagpu_grid *g, const float *queries, uint64_t nq, int space,
                       mdagpu_stream_t stream, mdagpu_pairs **out);
/* FastNS.self_search   nsgrid.pyx:582-644 : each unordered pair once */
MDAGPU_API int mdagpu_grid_self_search(mdagpu_grid *g, mdagpu_stream_t stream, mdagpu_pairs **out);

/* NSResults            nsgrid.pyx:127-192 */
MDAGPU_API uint64_t mdagpu_pairs_count(const mdagpu_pairs *p);
/* get_pairs() -> int64[count,2]; get_pair_distances() -> sqrt(d2) float64[count].
 * Either output may be NULL.  `space` says where the destination lives. */
MDAGPU_API int mdagpu_pairs_export(const mdagpu_pairs *p, int64_t *pairs, double *distances, int space,
                        mdagpu_stream_t stream);
MDAGPU_API void mdagpu_pairs_free(mdagpu_pairs *p);

/* _bruteforce_capped / _bruteforce_capped_self   lib/distances.py:589-673, 1050-1130:
 * all (i,j) with min_cutoff < d <= max_cutoff where d is the float64 distance
 * of distance_array / self_distance_array (box_mode as for those).  conf ==
 * NULL selects the self variant (i < j).  min_cutoff < 0 disables the lower
 * bound.  Never materialises the n*m matrix. */
MDAGPU_API int mdagpu_capped_bruteforce(const float *ref, uint64_t n, const float *conf, uint64_t m,
                             double max_cutoff, double min_cutoff, int box_mode, const float *box,
                             int space, mdagpu_stream_t stream, mdagpu_pairs **out);

/* ---- QCP (lib/qcprot.pyx) ------------------------------------------------------ */

/* CalcRMSDRotationalMatrix  qcprot.pyx:282-320, float64 flavour: ref, conf (N,3)
 * already centred; rot (9) and weights (N) may be NULL.  Host or device arrays;
 * *rmsd is a host scalar (the call synchronises the stream). */
MDAGPU_API int mdagpu_qcp(const double *ref, const double *conf, uint64_t n, double *rot,
               const double *weights, double *rmsd, int space, mdagpu_stream_t stream);

/* The two public halves of the above (lib/qcprot.pyx:179-277 InnerProduct, :326-543
 * FastCalcRMSDAndRotation), for callers that use them separately:
 *   inner_product: A[9] = sum_i w_i coords1_i (x) coords2_i (row = component of coords1),
 *                  *E0 = (sum w |coords1|^2 + sum w |coords2|^2) / 2; float64 [n,3] inputs
 *   solve:         RMSD (and rot[9], unless NULL) from A, E0 and the atom count n; host pointers */
MDAGPU_API int mdagpu_qcp_inner_product(const double *coords1, const double *coords2, uint64_t n,
               const double *weights, double *A, double *E0, int space, mdagpu_stream_t stream);
MDAGPU_API int mdagpu_qcp_solve(const double *A, double E0, double n, double *rot, double *rmsd);

/* Batched driver for analysis.rms.RMSD._single_frame (rms.py:685-727) and
 * analysis.align.AlignTraj._single_frame / _fit_to (align.py:310-366, 843-854).
 *   traj      float32 [F, natoms, 3]
 *   sel       int32 [nsel] indices of the fitted atoms inside a frame, or NULL
 *             for "the first nsel atoms" (nsel == natoms: all atoms)
 *   ref       float32 [nsel,3] un-centred reference positions of those atoms
 *   weights   float64 [nsel] relative weights (already divided by their mean,
 *             rms.py:639-643) or NULL
 *   rmsd      float64 [F]                      (out)
 *   rot       float64 [F,9] or NULL            (out; R with conf @ R ~ ref)
 *   aligned   if non-NULL (float32 [F,natoms,3], may alias traj): every atom of
 *             the frame is moved as mobile.translate(-com).rotate(R).translate(ref_com)
 *             with the three float32 roundings of core/groups.py:1478-1529.
 *   allclose_shortcut  1 reproduces align.rotation_matrix's early exit
 *             (align.py:290-291: R = I, rmsd = 0 when np.allclose(mobile, ref)
 *             and weights is NULL).
 * Centres are float64 means / weighted means as AtomGroup.center
 * (core/groups.py:1083-1106). */
MDAGPU_API int mdagpu_qcp_trajectory(const float *traj, uint64_t nframes, uint64_t natoms, const int32_t *sel,
                          uint64_t nsel, const float *ref, const double *weights, double *rmsd,
                          double *rot, float *aligned, int align_convention,
                          int allclose_shortcut, int space, mdagpu_stream_t stream);

/* All-pairs RMSD matrix of a trajectory: analysis.diffusionmap.DistanceMatrix._single_frame
 * (diffusionmap.py:238-268, a Python loop over frame pairs) with the metric rms.rmsd
 * (rms.py:178-283).
 *   traj      float32 [F, natoms, 3]
 *   sel       int32 [nsel] indices of the atoms inside a frame, or NULL for the first nsel
 *   weights   float64 [nsel] RELATIVE weights (divided by their mean, rms.py:270) or NULL
 *   superposition  0: sqrt(sum(w (a - b)^2) / nsel)  (the DistanceMatrix default metric)
 *                  1: rmsd(a, b, superposition=True): both frames centred on their weighted
 *                     mean, then QCP (qcprot.pyx:282-543)
 *   cutoff    entries <= cutoff are stored as 0 (diffusionmap.py:262-263)
 *   out       float64 [F, F], symmetric (out) */
MDAGPU_API int mdagpu_pairwise_rmsd(const float *traj, uint64_t nframes, uint64_t natoms, const int32_t *sel,
                         uint64_t nsel, const double *weights, int superposition, double cutoff,
                         double *out, int space, mdagpu_stream_t stream);

/* ---- pair engine over many frames: InterRDF and contacts -------------------------- */

/* analysis/rdf.py:268-291 for all frames at once: capped_distance(A[f], B[f], cutoff,
 * box[f]) followed by np.histogram, accumulated over frames.
 *   A, B        float32 [F,na,3], [F,nb,3]; B == NULL or B == A: the same group
 *               (pairs (i,i) included, exactly like capped_distance(x, x))
 *   box6, h9    host arrays, float32 [nbox,6] unit cells and their float32
 *               triclinic_vectors [nbox,9]; nbox = F (boxes_per_frame = 1) or 1
 * The caller (host logic mirroring lib/distances.py:516-586) has decided the
 * reference's method for this frame set and passes its arithmetic:
 *   use_grid    1: FastNS semantics (reference grid geometry, 27-cell stencil,
 *               nsgrid.pyx:294-580); 0: brute force (all pairs)
 *   mic         MDAGPU_MIC_* flavour of the pair distance
 *   wrap        1: coordinates are wrapped into the cell first (FastNS, triclinic)
 *   cutoff      search radius (grid geometry);
 *   d2_max      pairs with d2 <= d2_max are kept: cutoff^2 for FastNS
 *               (nsgrid.pyx:540), the largest d2 with sqrt(d2) <= cutoff for the
 *               brute-force path (distances.py:661-664)
 *   thresholds  host float64 [nbins+1]: t[k] = smallest d2 whose sqrt is >= edges[k]
 *               for k < nbins, t[nbins] = largest d2 whose sqrt is <= edges[nbins];
 *               bin k receives t[k] <= d2 < t[k+1] (last bin closed), which is
 *               np.histogram(np.sqrt(d2), bins=nbins, range=(r_lo, r_hi))
 *   excl_a/b    exclusion_block (0,0 = none)                       rdf.py:274-278
 *   counts      uint64 [nbins], accumulated (+=); host or device per `space` */
MDAGPU_API int mdagpu_rdf_accumulate(const float *A, uint64_t na, const float *B, uint64_t nb,
                          uint64_t nframes, const float *box6, const float *h9,
                          int boxes_per_frame, int mic, int wrap, double cutoff, double d2_max,
                          int use_grid, const double *thresholds, int nbins, double r_lo,
                          double r_hi, int excl_a, int excl_b, unsigned long long *counts,
                          int space, mdagpu_stream_t stream);

/* InterRDF(exclude_same=...)  analysis/rdf.py:280-285: as mdagpu_rdf_accumulate, but a pair (i, j) is
 * dropped when ids_a[i] == ids_b[j] (host int32 [na], [nb]: residue / segment / chain index of each
 * atom; ids_b may be NULL for the same group). */
MDAGPU_API int mdagpu_rdf_accumulate_ids(const float *A, uint64_t na, const float *B, uint64_t nb,
                              uint64_t nframes, const float *box6, const float *h9,
                              int boxes_per_frame, int mic, int wrap, double cutoff, double d2_max,
                              int use_grid, const double *thresholds, int nbins, double r_lo,
                              double r_hi, const int32_t *ids_a, const int32_t *ids_b,
                              unsigned long long *counts, int space, mdagpu_stream_t stream);

/* InterRDF_s  analysis/rdf.py:585-597: as mdagpu_rdf_accumulate, but one histogram per
 * (site i of A, site j of B): counts is uint64 [na, nb, nbins] (nb = na when B is NULL /
 * the same array), accumulated into. */
MDAGPU_API int mdagpu_rdf_sites_accumulate(const float *A, uint64_t na, const float *B, uint64_t nb,
                                uint64_t nframes, const float *box6, const float *h9,
                                int boxes_per_frame, int mic, int wrap, double cutoff,
                                double d2_max, int use_grid, const double *thresholds,
                                int nbins, double r_lo, double r_hi,
                                unsigned long long *counts, int space, mdagpu_stream_t stream);

/* Fused self_distance_array -> contact count / per-atom degree
 * (analysis/contacts.py:315-338 `d <= radius`, analysis/distances.py:102-108)
 * for nframes frames [F,n,3]: pairs i<j whose self_distance_array distance has
 * d2 <= d2_max (the caller converts `d <= radius` as above; `cutoff` = radius is
 * used for the cell size only).  counts uint64 [F]; degree int32 [F,n] or NULL.
 * Never materialises the n(n-1)/2 distances. */
MDAGPU_API int mdagpu_contacts_self(const float *xyz, uint64_t n, uint64_t nframes, int box_mode,
                         const float *box, double d2_max, double cutoff,
                         unsigned long long *counts, int32_t *degree, int space,
                         mdagpu_stream_t stream);

/* ---- trajectory ingest: the step in front of the path ---- */

/* ---- reductions that follow the hot kernels inside the analyses (csrc/reduce.cu) ----
 * frame_mean:       mean[nsel,3] (float64) = mean over the frames of traj[f, sel] -- AverageStructure
 *                   (analysis/align.py:1072-1083); sel = NULL: the first nsel atoms
 * group_rmsd:       out[f] = sqrt(sum_i w_i |traj[f, idx_i] - ref_i|^2 / g) -- RMSD groupselections
 *                   (analysis/rms.py:716-721); weights already divided by their mean, or NULL
 * contact_fraction: out[f] = fraction of native contacts of frame f from its pair distances r[f, :]
 *                   method 0 hard_cut_q (r <= r0), 1 radius_cut_q (r <= radius),
 *                   2 soft_cut_q 1 / (1 + exp(beta (r - lambda r0))) -- analysis/contacts.py:231-312 */
MDAGPU_API int mdagpu_frame_mean(const float *traj, uint64_t nframes, uint64_t natoms, const int32_t *sel,
               uint64_t nsel, double *mean, int space, mdagpu_stream_t stream);
MDAGPU_API int mdagpu_group_rmsd(const float *traj, uint64_t nframes, uint64_t natoms, const int32_t *idx,
               uint64_t g, const double *ref, const double *weights, double *out, int space,
               mdagpu_stream_t stream);
MDAGPU_API int mdagpu_contact_fraction(const double *r, uint64_t nframes, uint64_t npairs, const double *r0,
               int method, double radius, double beta, double lambda, double *out, int space,
               mdagpu_stream_t stream);

/* Bulk DCD reader: coordinates/DCD.py:279-344 (DCDReader.timeseries) ->
 * lib/formats/libdcd.pyx:583-700 (DCDFile.readframes), C side lib/formats/include/readdcd.h.
 * Host code; CHARMM and X-PLOR flavours, either endianness, 4-byte record markers,
 * no fixed atoms. */
typedef struct mdagpu_dcd mdagpu_dcd;
MDAGPU_API int mdagpu_dcd_open(const char *path, mdagpu_dcd **out);
MDAGPU_API void mdagpu_dcd_close(mdagpu_dcd *d);
/* header fields (libdcd.pyx:279-330 `header`): natoms, number of frames (from the file
 * size, as the reference), unit cell records present, ISTART, NSAVC, DELTA, CHARMM version */
MDAGPU_API int mdagpu_dcd_header(const mdagpu_dcd *d, int64_t *natoms, int64_t *nframes, int *has_cell,
                      int32_t *istart, int32_t *nsavc, double *delta, int *charmm);
/* frames range(start, stop, step) of atoms `indices` (NULL: all) into
 *   xyz       float32 [n, nsel, 3] ('fac' order; pinned memory recommended)
 *   unitcell  float64 [n, 6] in the FILE's order [A, gamma, B, beta, alpha, C], or NULL */
MDAGPU_API int mdagpu_dcd_readframes(mdagpu_dcd *d, int64_t start, int64_t stop, int64_t step,
                          const int64_t *indices, int64_t nsel, float *xyz, double *unitcell);

#ifdef __cplusplus
}
#endif
#endif /* MDAGPU_H */
