/*
 * picca_b200.h -- C ABI of libpicca_b200.so: the B200 (sm_100a) implementation of picca's forest
 * pair-counting hot path.  Plain pointers and sizes only; no torch / C++ types.
 *
 * The reference has no native layer: its "plugin API" for this path is the Python module surface
 * of picca.cf / picca.xcf (module globals + fill_neighs / compute_xi / compute_dmat, reference
 * py/picca/cf.py:28-79,82,138,390 and py/picca/xcf.py:27-68,71,126,325).  picca_b200/cf.py and
 * picca_b200/xcf.py mirror that surface and bind the entry points below with ctypes; each entry
 * point names the reference code it replaces.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer inside pb2_catalog / pb2_pairs / passed as `d_*` is a DEVICE pointer;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions return 0 on success, a positive cudaError_t value or a negative PB2_E* code on
 *     failure; pb2_last_error() returns the message for the calling thread;
 *   - all launches are asynchronous on `stream` unless stated otherwise.
 */
#ifndef PICCA_B200_H
#define PICCA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB2_ABI_VERSION 23

/* layout constants of the packed copies read by the diagonal-lane xi kernel */
#ifndef PB2_DIAG_LANES
#define PB2_DIAG_LANES 2                     /* adjacent diagonals per lane */
#endif
#define PB2_DIAG_PAD (34 * PB2_DIAG_LANES)   /* dummy pixels either side, interleaved copy */
#define PB2_DIAG_ROW_PAD 8                   /* dummy pixels after a line of sight, natural copy */
#ifndef PB2_DIAG_CHUNK_ROWS
#define PB2_DIAG_CHUNK_ROWS 32               /* rows staged in shared memory per TMA bulk copy */
#endif

#define PB2_EINVAL (-1)   /* bad argument */
#define PB2_ECONFIG (-2)  /* configuration not supported by the kernels (message says which) */

/* Module globals of picca.cf / picca.xcf, snapshotted at call time (cf.py:28-79, xcf.py:27-68). */
typedef struct pb2_params {
    int32_t num_bins_r_par;         /* np  */
    int32_t num_bins_r_trans;       /* nt  */
    int32_t num_model_bins_r_par;   /* npm */
    int32_t num_model_bins_r_trans; /* ntm */
    double r_par_min;
    double r_par_max;
    double r_trans_max;
    int32_t has_z_min_pairs;
    int32_t has_z_max_pairs;
    double z_min_pairs;
    double z_max_pairs;
    int32_t has_zerr_cut;
    int32_t x_correlation;
    double zerr_cut_deg;
    double zerr_cut_kms;
    int32_t rmu_binning;
    int32_t ang_correlation;
    int32_t remove_same_half_plate_close_pairs;
    int32_t redshift_evolution_in_distortion_matrix;
    double z_ref;
    double alpha;
    double alpha2; /* cf: alpha2 ; xcf: alpha_obj */
    double ang_max;
} pb2_params;

/*
 * A catalogue of lines of sight packed as SoA + CSR in HBM (replaces the dict[healpix] ->
 * list[Delta|QSO] object graph the reference walks, py/picca/data.py:14-162,238-373).
 * Lines of sight are stored in ascending-HEALPix, list order -- the iteration order of
 * fill_neighs (cf.py:91-122).  A quasar catalogue (xcf `objs`) is the same structure with one
 * pixel per object: r_comov/dist_m/weights scalars, z = z_qso.
 */
typedef struct pb2_catalog {
    int64_t n_los;            /* number of forests / objects */
    int64_t n_pix;            /* total pixels = offset[n_los] */
    const int64_t *offset;    /* [n_los+1] first pixel of each line of sight */
    /* per pixel, fp64 */
    const double *r_comov;    /* Delta.r_comov (or 10**log_lambda when ang_correlation) */
    const double *dist_m;     /* Delta.dist_m  (same remark) */
    const double *z;          /* Delta.z */
    const double *weights;    /* Delta.weights */
    const double *delta_w;    /* Delta.delta * Delta.weights, 0 where weights == 0 */
    const double *z_w;        /* Delta.z * Delta.weights */
    const double *log_lambda; /* Delta.log_lambda (distortion matrix only; may be NULL) */
    /* ---- packed copies read by the diagonal-lane xi kernel (pb2_xi_diag.cu); NULL for object
     * catalogues.  Zero-weight pixels (never counted by the reference, cf.py:318,331) are
     * compacted away; dg_count[f] pixels of line of sight f remain.  A pixel is a 48-byte record
     * (r_comov, dist_m, weights, delta_w, z / 2, 0).
     * Natural order (row records, window searches): pixel i of line of sight f is record
     * dg_offset[f] + i of dg_rec, followed by PB2_DIAG_ROW_PAD dummies (distance 1e299, zeros
     * elsewhere); PB2_DIAG_CHUNK_ROWS more dummies end the array. */
    const int64_t *dg_offset;    /* [n_los] */
    const int32_t *dg_count;     /* [n_los] */
    const double *dg_rec;
    /* Interleaved by PB2_DIAG_LANES (column records): with jp = j + PB2_DIAG_PAD, pixel j of line
     * of sight f is record il_offset[f] + jp / PB2_DIAG_LANES of plane jp % PB2_DIAG_LANES; plane
     * p starts at record p * il_total of il_rec.  Every line of sight is padded with PB2_DIAG_PAD
     * dummies (distance 1e300, zeros elsewhere) on both sides, and every plane ends with
     * PB2_DIAG_CHUNK_ROWS + 64 more, so a warp may copy whole chunks of a diagonal block past
     * either end without a bounds check. */
    const int64_t *il_offset;    /* [n_los] */
    int64_t il_total;            /* records per plane */
    const double *il_rec;
    int32_t dg_lanes;            /* PB2_DIAG_LANES the copies were packed for */
    int32_t dg_max_pix;          /* longest compacted line of sight */
    int32_t dg_ok;               /* 1 if every r_comov, dist_m, z, weight, delta_w is finite */
    int32_t dg_reserved;
    double dg_reach;             /* max(|r_comov|, |dist_m|) over the catalogue */
    /* ---- per-forest prefix sums read by the forest x object kernel (pb2_xcf.cu), filled by
     * pb2_build_prefix; may be NULL.  Entry offset[f] + f + i (six doubles) = sums over pixels
     * [0, i) of line of sight f of weights, delta_w, (r_comov - r_comov[0]) weights,
     * (dist_m - dist_m[0]) weights, z_w and the number of non-zero weights; n + 1 entries. */
    const double *px_rec;
    /* per line of sight */
    const double *x_cart, *y_cart, *z_cart, *ra, *dec, *cos_dec, *z_qso;
    const int64_t *thingid, *plate, *fiberid;
    const int32_t *order;     /* Delta.order, -1 when None */
    const int32_t *row;       /* index of the line of sight's HEALPix pixel in sorted(data) */
    /* per HEALPix pixel: member range and a bounding cap (centre, angular radius) of members */
    int32_t n_hp;
    int32_t sorted;           /* 1 if r_comov and dist_m are non-decreasing inside every forest */
    const int32_t *hp_first;  /* [n_hp+1] */
    const double *cap_x, *cap_y, *cap_z, *cap_rad;
    int32_t max_pix;          /* longest line of sight */
    int32_t reserved;
} pb2_catalog;

/*
 * Neighbour (forest-pair) list in CSR form over the `n_f1` lines of sight selected by f1_index.
 * Order inside a segment = ascending index in the second catalogue = the reference's neighbour
 * order (ascending HEALPix id, then list order), which the --rej RNG contract depends on.
 */
typedef struct pb2_pairs {
    int64_t n_f1;
    int64_t n_pairs;
    const int32_t *f1_index;   /* [n_f1] line-of-sight index in catalogue 1 */
    const int64_t *nb_offset;  /* [n_f1+1] */
    const int32_t *nb_f1;      /* [n_pairs] position k in f1_index of the owning line of sight */
    const int32_t *nb_f2;      /* [n_pairs] line-of-sight index in catalogue 2 */
    const double *nb_ang;      /* [n_pairs] angular separation (get_angle_between) */
    const double *nb_cos;      /* [n_pairs] cos(ang/2) */
    const double *nb_sin;      /* [n_pairs] sin(ang/2) */
    const uint8_t *nb_keep;    /* [n_pairs] or NULL: 0 = pair dropped by the --rej draw (dmat) */
} pb2_pairs;

int32_t pb2_abi_version(void);
const char *pb2_last_error(void);
int32_t pb2_sizeof_params(void);
int32_t pb2_sizeof_catalog(void);
int32_t pb2_sizeof_pairs(void);
int32_t pb2_diag_lanes(void); /* PB2_DIAG_LANES the library was built with */

/* ---- neighbour search: replaces cf.fill_neighs (cf.py:82-135) and xcf.fill_neighs
 * (xcf.py:71-123), including QSO.get_angle_between (data.py:106-162).
 * mode 0: auto  (same catalogue, thingid != and ra1 > ra2, cf.py:109,129-135)
 * mode 1: cross (two delta catalogues, thingid != only, cf.py:99-110)
 * mode 2: xcf   (forest x object: thingid !=, optional zerr cut xcf.py:102-115 and the r_par
 *                pre-filter xcf.py:117-121)
 * Pass 1 writes the neighbour count of every selected line of sight; the caller turns counts into
 * nb_offset (exclusive scan) and pass 2 fills nb_f1/nb_f2/nb_ang/nb_cos/nb_sin in reference order. */
int32_t pb2_neigh_count(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                        int32_t mode, int64_t n_f1, const int32_t *d_f1_index,
                        int32_t *d_count, void *stream);
int32_t pb2_neigh_fill(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                       int32_t mode, int64_t n_f1, const int32_t *d_f1_index,
                       const int64_t *d_nb_offset, int32_t *d_nb_f1, int32_t *d_nb_f2,
                       double *d_nb_ang, double *d_nb_cos, double *d_nb_sin, void *stream);

/* ---- auto / delta-delta cross correlation: replaces cf.compute_xi's pair loop and
 * cf.compute_xi_forest_pairs_fast (cf.py:161-240, 250-387).
 * d_out is [n_rows][6][np*nt]: un-normalised sums of weight, xi, r_par, r_trans, z (fp64) and
 * num_pairs (int64 stored in the same 8-byte slots); it is accumulated into (caller zeroes it).
 * d_out_row[k] is the output row of f1_index[k].  The per-call normalisation of cf.py:242-246
 * is pb2_xi_normalise.  `variant`: 0 = product (diagonal-lane kernel for the standard binning
 * without per-pair cuts, the general row-tile kernel otherwise), 1 = brute-force validation
 * kernel, 2 = force the general row-tile kernel, 3 = force the specialised row-tile kernel
 * (same results; used by tests and for comparison). */
int32_t pb2_xi_auto(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                    const pb2_pairs *pairs, const int32_t *d_out_row, int64_t n_rows,
                    double *d_out, int32_t variant, void *stream);

/* ---- forest x object correlation: replaces xcf.compute_xi's loop and
 * xcf.compute_xi_forest_pairs_fast (xcf.py:149-213, 223-322).  Same output layout; `n_rows` =
 * rows of d_out (sizes the bin-major scratch histogram of the product kernel).  `variant`: 0 =
 * product (prefix-sum kernel with transposed reductions when the catalogue carries prefix records,
 * the forests are sorted and no per-pair z cut is set; the general lane = object kernel otherwise),
 * 1 = validation (every pair through the reference expression), 2 = force the general kernel,
 * 4 = prefix-sum kernel with per-lane reductions (same results; used by the tests). */
int32_t pb2_xi_cross(const pb2_catalog *cat1, const pb2_catalog *objs, const pb2_params *par,
                     const pb2_pairs *pairs, const int32_t *d_out_row, int64_t n_rows,
                     double *d_out, int32_t variant, void *stream);

/* ---- per-pixel products of the packed catalogue, formed in HBM when the host deferred them
 * (catalog.pack(defer_products=True)): delta_w = Delta.delta * Delta.weights -- the product the
 * reference forms first, cf.py:367-368 -- set to 0 where the weight is 0, and z_w = z * weights. */
/* ---- one pass over the packed SoA in HBM for what catalog.pack otherwise computes with ~12 NumPy
 * passes on the host: d_count[f] = pixels of line of sight f with a non-zero weight (the reference
 * never visits the others, cf.py:318,331); d_flags[0] != 0: a kept pixel is not finite;
 * d_flags[1] != 0: r_comov or dist_m decreases inside a line of sight (or is not finite);
 * d_flags[2]: bit pattern of max(|r_comov|, |dist_m|) over the kept pixels.  d_flags zeroed by the
 * caller. */
int32_t pb2_catalog_stats(int64_t n_los, const int64_t *d_offset, const double *d_weights,
                          const double *d_r_comov, const double *d_dist_m, const double *d_z,
                          const double *d_delta_w, int32_t *d_count, int64_t *d_flags, void *stream);
int32_t pb2_derive_products(int64_t n_pix, const double *d_weights, const double *d_delta,
                            const double *d_z, double *d_delta_w, double *d_z_w, void *stream);

/* Write the packed record copies of the diagonal-lane xi kernel (cat->dg_rec: dg_total records,
 * cat->il_rec: PB2_DIAG_LANES * il_total records; layout above) from the SoA arrays of the
 * catalogue on the device.  cat->dg_offset / dg_count / il_offset / il_total describe the layout. */
int32_t pb2_pack_diag(const pb2_catalog *cat, int64_t dg_total, void *stream);

/* Fill the prefix records of a delta catalogue: d_px_rec holds 6 * (n_pix + n_los) doubles (see
 * pb2_catalog.px_rec); the caller then stores the pointer in the catalogue it passes on. */
int32_t pb2_build_prefix(const pb2_catalog *cat, double *d_px_rec, void *stream);

/* xi, r_par, r_trans, z /= weights where weights > 0, per row (cf.py:242-246, xcf.py:215-219). */
int32_t pb2_xi_normalise(int64_t n_rows, int32_t nb, double *d_out, void *stream);

/* ---- distortion matrix: replaces cf.compute_dmat's pair loop + cf.compute_dmat_forest_pairs_fast
 * (cf.py:424-502, 520-887) and the xcf equivalents (xcf.py:360-409, 427-674).  pairs->nb_keep
 * carries the host-drawn --rej mask (cf.py:444 / xcf.py:379).  Outputs are accumulated into:
 * d_dmat [nb][nbm], d_weights_dmat [nb], d_r_par_eff/d_r_trans_eff/d_z_eff/d_weight_eff [nbm].
 * d_scratch/scratch_bytes: workspace (query the size with pb2_dmat_scratch_bytes). */
int64_t pb2_dmat_scratch_bytes(const pb2_catalog *cat1, const pb2_catalog *cat2,
                               const pb2_params *par, int32_t cross);
int32_t pb2_dmat_auto(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                      const pb2_pairs *pairs, double *d_weights_dmat, double *d_dmat,
                      double *d_r_par_eff, double *d_r_trans_eff, double *d_z_eff,
                      double *d_weight_eff, void *d_scratch, int64_t scratch_bytes, void *stream);
int32_t pb2_dmat_cross(const pb2_catalog *cat1, const pb2_catalog *objs, const pb2_params *par,
                       const pb2_pairs *pairs, double *d_weights_dmat, double *d_dmat,
                       double *d_r_par_eff, double *d_r_trans_eff, double *d_z_eff,
                       double *d_weight_eff, void *d_scratch, int64_t scratch_bytes, void *stream);

/* ---- metal distortion matrix, forest x forest: replaces the pair loop of
 * cf.compute_metal_dmat (py/picca/cf.py:890-1232) for one (absorber of forest 1, absorber of
 * forest 2) pass; the caller runs the swapped pass too when cf.py:1089-1091 asks for it.
 * d_z*, d_rc*, d_dm*, d_pw*: per pixel of the catalogue, the redshift / r_comov / dist_m the pixel
 * has for that absorber (cf.py:944-946, :976-978) and (1+z)^(alpha_abs-1); evol_den =
 * (1+z_ref)^(alpha_abs1+alpha_abs2-2) (cf.py:1033-1037).  pairs->nb_keep = the --rej mask.
 * Outputs are accumulated into, as pb2_dmat_auto. */
int32_t pb2_metal_dmat_auto(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                            const pb2_pairs *pairs, const double *d_z1, const double *d_rc1,
                            const double *d_dm1, const double *d_pw1, const double *d_z2,
                            const double *d_rc2, const double *d_dm2, const double *d_pw2,
                            double evol_den, double *d_weights_dmat, double *d_dmat,
                            double *d_r_par_eff, double *d_r_trans_eff, double *d_z_eff,
                            double *d_weight_eff, void *stream);

/* forest x object: replaces the pair loop of xcf.compute_metal_dmat (py/picca/xcf.py:677-835);
 * d_pw1 = ((1+z_abs)/(1+z_ref))^(alpha_abs-1) per forest pixel (xcf.py:769-771). */
int32_t pb2_metal_dmat_cross(const pb2_catalog *cat1, const pb2_catalog *objs, const pb2_params *par,
                             const pb2_pairs *pairs, const double *d_z1, const double *d_rc1,
                             const double *d_dm1, const double *d_pw1, double *d_weights_dmat,
                             double *d_dmat, double *d_r_par_eff, double *d_r_trans_eff,
                             double *d_z_eff, double *d_weight_eff, void *stream);

/* ---- Wick expansion of the covariance.
 * pb2_wick_auto replaces the forest-pair loop of cf.compute_wick_terms and
 * cf.compute_wickT123_pairs (py/picca/cf.py:1326-1494, :1497-1626; diagrams T1-T3, i.e.
 * max_diagram <= 3) over the forest pairs flagged in pairs->nb_keep (the caller's per-forest
 * --rej draw, cf.py:1378).  d_var*: get_variance_1d(log_lambda) per pixel of the catalogue
 * (cf.py:1412); d_ze*: ((1+z)/(1+z_ref))^(alpha-1) per pixel (cf.py:1556-1557); the 1-D
 * correlation xi_1d (picca_wick.py:412-417: scipy interp1d, kind "nearest", extrapolating) comes
 * as a table: n_x values d_xy and the n_x - 1 decision bounds d_xb between them.
 * Outputs are accumulated into: weights_wick[nb], num_pairs_wick[nb] (int64), t1/t2/t3 [nb][nb].
 * pb2_wick_cross: xcf.compute_wick_terms / compute_wickT1234_pairs (py/picca/xcf.py:838-1153,
 * :1219-1351; T1-T4) over the forests flagged in d_keep_f1 [pairs->n_f1] with ALL their
 * neighbouring objects; d_ze_obj: ((1+z_q)/(1+z_ref))^(alpha_obj-1) per object;
 * max_neighbours: the longest neighbour list among the flagged forests.
 * Scratch: pb2_wick_scratch_bytes(longest forest of catalogue 1, longest forest of catalogue 2
 * or max_neighbours, cross). */
int64_t pb2_wick_scratch_bytes(int64_t max_pix1, int64_t max_rows2, int32_t cross);
int32_t pb2_wick_auto(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                      const pb2_pairs *pairs, const double *d_var1, const double *d_ze1,
                      const double *d_var2, const double *d_ze2, int32_t n_x1,
                      const double *d_xb1, const double *d_xy1, int32_t n_x2,
                      const double *d_xb2, const double *d_xy2, double *d_weights_wick,
                      int64_t *d_num_pairs_wick, double *d_t1, double *d_t2, double *d_t3,
                      void *d_scratch, int64_t scratch_bytes, void *stream);
int32_t pb2_wick_cross(const pb2_catalog *cat1, const pb2_catalog *objs, const pb2_params *par,
                       const pb2_pairs *pairs, const uint8_t *d_keep_f1, int64_t max_neighbours,
                       const double *d_var1, const double *d_ze1, const double *d_ze_obj,
                       int32_t n_x1, const double *d_xb1, const double *d_xy1,
                       double *d_weights_wick, int64_t *d_num_pairs_wick, double *d_t1,
                       double *d_t2, double *d_t3, double *d_t4, void *d_scratch,
                       int64_t scratch_bytes, void *stream);

/* ---- object x object pair counting: replaces the pair loop of
 * co.compute_xi / co.compute_xi_forest_pairs (py/picca/co.py:77-132, :135-202).  pairs = the
 * neighbour list of pb2_neigh_* in mode 1 (co.py:43-69); the mean-redshift cut of co.py:70-74 is
 * applied per pair when has_z_cut; take_abs = `not x_correlation or type_corr in (DR, RD)`
 * (co.py:172).  d_out [n_rows][5][np*nt]: weight, r_par*w, r_trans*w, z*w sums and int64 counts. */
int32_t pb2_co_pairs(const pb2_catalog *objs1, const pb2_catalog *objs2, const pb2_params *par,
                     const pb2_pairs *pairs, int32_t take_abs, int32_t has_z_cut, double z_cut_min,
                     double z_cut_max, const int32_t *d_out_row, int64_t n_rows, double *d_out,
                     void *stream);

/* ---- sub-sample covariance of the per-HEALPix blocks (the consumer of the WE/DA columns the
 * pair kernels produce).
 * pb2_cov_subsample replaces utils.compute_cov (py/picca/utils.py:100-128): d_xi, d_weights are
 * [n_samples][nb] row-major fp64 (one row per HEALPix pixel); outputs d_cov [nb][nb],
 * d_mean_xi [nb] (weighted mean, utils.py:113-116) and d_sum_weights [nb].  Workspace of
 * pb2_cov_scratch_bytes(n_samples, nb) bytes.
 * pb2_cov_smooth replaces utils.smooth_cov (py/picca/utils.py:153-249) for a covariance without
 * zero variances (the caller handles the reference's early return, utils.py:187-190): the
 * correlation coefficient is averaged over bin pairs with equal
 * (round(|dr_par|/delta_r_par), round(|dr_trans|/delta_r_trans)) and, with per_r_par, equal
 * int(r_par/delta_r_par).  n_dp, n_dt: extents of the two rounded differences; rp_lo, n_rp: lowest
 * value and extent of int(r_par/delta_r_par) (per_r_par only).  d_table_sum / d_table_count:
 * n_rp*n_dp*n_dt entries of workspace; *d_bad is set to 1 if a key fell outside the extents. */
int64_t pb2_cov_scratch_bytes(int64_t n_samples, int32_t nb);
int32_t pb2_cov_subsample(int64_t n_samples, int32_t nb, const double *d_xi, const double *d_weights,
                          double *d_cov, double *d_mean_xi, double *d_sum_weights, void *d_scratch,
                          int64_t scratch_bytes, void *stream);
/* utils.compute_cov_boot (py/picca/utils.py:131-150; picca_export.py --num-boot-cov): d_idx
 * [n_boot][n_samples] int32 = the host-drawn `default_rng(seed).choice(nhpx, size=nhpx)` of every
 * realisation (the reference's RNG stream); output d_cov [nb][nb] = np.cov of the bootstrap means. */
int64_t pb2_cov_boot_scratch_bytes(int64_t n_samples, int32_t nb, int32_t n_boot);
int32_t pb2_cov_boot(int64_t n_samples, int32_t nb, int32_t n_boot, const double *d_xi,
                     const double *d_weights, const int32_t *d_idx, double *d_cov, void *d_scratch,
                     int64_t scratch_bytes, void *stream);
int32_t pb2_cov_smooth(int32_t nb, const double *d_cov, const double *d_r_par,
                       const double *d_r_trans, double delta_r_par, double delta_r_trans,
                       int32_t per_r_par, int32_t n_dp, int32_t n_dt, int32_t rp_lo, int32_t n_rp,
                       double *d_table_sum, uint64_t *d_table_count, int32_t *d_bad,
                       double *d_cov_smooth, void *stream);

/* ---- delta loader: FITS delta files -> the SoA CSR buffers.
 * Host side (plain C, no GPU): pb2_fits_scan walks the HDUs of a decompressed FITS buffer,
 * info[h] = {header_off, data_off, data_bytes, bitpix, naxis, naxis1, naxis2, tfields}; returns the
 * number of HDUs or -1.  pb2_fits_cards extracts header cards of many HDUs (keys: n_keys x 8 blank-
 * padded characters; per HDU and key: kind 0 missing / 1 number / 2 string / 3 logical, value as
 * double, as int64 for integer literals, and as a 24-character string).  Together they replace the
 * per-HDU fitsio calls of io.read_delta_file (py/picca/io.py:354-360) and Delta.from_fitsio
 * (py/picca/data.py:392-474).
 * pb2_delta_unpack: big-endian BinTable rows (the raw file bytes, uploaded as they are) -> native
 * fp64 arrays at the CSR offsets: forest f has rows at d_raw + d_row0[f], d_row_bytes[f] apart,
 * with the wavelength (LOGLAM or LAMBDA), DELTA and WEIGHT columns at d_col_off[3f..3f+2].
 * pb2_delta_prepare: the per-forest loop of io.read_deltas (py/picca/io.py:493-507): z =
 * 10^log_lambda / lambda_abs - 1 (or d_z_in when given: parity mode with the host's power),
 * r_comov / dist_m by linear interpolation on the n_table-point cosmology table exactly as
 * scipy's interp1d evaluates it (py/picca/constants.py:211-229; skipped when d_tab_z is NULL, the
 * reference's cosmo=None), weights *= ((1+z)/(1+z_ref))^(alpha-1), Delta.project
 * (py/picca/data.py:622-655) when project != 0 (d_order: 0/1 per forest), d_z_range[2f..2f+1] =
 * min/max z of forest f.  *d_status = 1 if a redshift fell outside the table (interp1d raises). */
int64_t pb2_fits_scan(const uint8_t *buf, int64_t len, int64_t max_hdu, int64_t *info);
int32_t pb2_fits_cards(const uint8_t *buf, int64_t len, int64_t n_hdu, const int64_t *header_off,
                       int32_t n_keys, const char *keys, int32_t *kind, double *num, int64_t *inum,
                       char *str);
int32_t pb2_delta_unpack(int64_t n_los, const uint8_t *d_raw, const int64_t *d_row0,
                         const int32_t *d_row_bytes, const int32_t *d_col_off,
                         const int64_t *d_offset, double *d_log_lambda, double *d_delta,
                         double *d_weights, void *stream);
/* ImageHDU flavour (Delta.from_image, py/picca/data.py:519-620): images [n_forest][n_lambda] of
 * big-endian fp64 at byte offsets delta_off / weight_off of d_raw, the common wavelength grid at
 * lambda_off; forest f is image row d_rows[f] and keeps the pixels with WEIGHT > 0.
 * pb2_delta_image_count gives the kept pixels per forest; after an exclusive scan,
 * pb2_delta_image_unpack compacts them into the CSR arrays. */
int32_t pb2_delta_image_count(int64_t n_los, const uint8_t *d_raw, int64_t weight_off,
                              int32_t n_lambda, const int32_t *d_rows, int32_t *d_count,
                              void *stream);
int32_t pb2_delta_image_unpack(int64_t n_los, const uint8_t *d_raw, int64_t lambda_off,
                               int64_t delta_off, int64_t weight_off, int32_t n_lambda,
                               const int32_t *d_rows, const int64_t *d_offset, double *d_log_lambda,
                               double *d_delta, double *d_weights, void *stream);
/* Delta.rebin (py/picca/data.py:657-686; read_deltas' rebin_factor, io.py:362-378).
 * pb2_fits_hierarch: one long keyword of the ESO HIERARCH convention (WAVE_SOLUTION, DELTA_LAMBDA)
 * from the header at header_off.  pb2_delta_wave: wave = 10^log_lambda per pixel (log10 first
 * when the file stores LAMBDA).  pb2_delta_rebin: d_new_offset == NULL -> d_count[f] = surviving
 * bins of forest f; else the bins are written at d_new_offset[f] (mid-point wavelength, weighted
 * mean delta, weight sum).  d_dwave[f] = DELTA_LAMBDA of the forest's file.  *d_status = 2 if a
 * forest's wavelengths are not ascending. */
int32_t pb2_fits_hierarch(const uint8_t *buf, int64_t len, int64_t header_off, const char *key,
                          int32_t *kind, double *num, char *str24);
int32_t pb2_delta_wave(int64_t n_pix, int32_t wave_is_lambda, double *d_log_lambda, double *d_wave,
                       void *stream);
int32_t pb2_delta_rebin(int64_t n_los, const int64_t *d_offset, const double *d_wave,
                        const double *d_delta, const double *d_weights, const double *d_dwave,
                        int32_t factor, int32_t *d_count, int32_t *d_status,
                        const int64_t *d_new_offset, double *d_new_wave, double *d_new_delta,
                        double *d_new_weights, void *stream);
int32_t pb2_delta_prepare(int64_t n_los, const int64_t *d_offset, const int32_t *d_order,
                          double lambda_abs, double alpha, double z_ref, int32_t n_table,
                          const double *d_tab_z, const double *d_tab_r_comov,
                          const double *d_tab_dist_m, int32_t project, int32_t wave_is_lambda,
                          const double *d_z_in, double *d_log_lambda, double *d_delta,
                          double *d_weights, double *d_z, double *d_r_comov, double *d_dist_m,
                          double *d_z_range, int32_t *d_status, void *stream);

/* measurement: statistics of the last pb2_dmat_auto call that used d_scratch -- out3[0] the FP64
 * ops the reference algorithm would execute as written (N_sel (15 U + 4) +
 * 40 N_inrange per used forest pair, cf.py:623-887), out3[1] the sum of U, out3[2] the in-range
 * pixel pairs.  Synchronises the stream. */
int32_t pb2_dmat_stats(const void *d_scratch, double *out3, void *stream);

/* ---- measurement helpers
 * pb2_fp64_peak: dependent-free DFMA microbenchmark; returns achieved FP64 op/s (1 DFMA = 1 op,
 * i.e. warp-instruction lanes per second) -- the measured denominator of the FP64 roofline.
 * Synchronous.  pb2_launch_count: number of kernels this library has launched in the process. */
int32_t pb2_fp64_peak(int32_t iters, double *ops_per_second, double *elapsed_ms);
int64_t pb2_launch_count(void);
/* time (ms) spent in the most recent pair-kernel launch sequence, measured with CUDA events on the
 * launching stream when timing is enabled with pb2_set_timing(1) (adds a stream sync). */
int32_t pb2_set_timing(int32_t enable);
double pb2_last_kernel_ms(void);

#ifdef __cplusplus
}
#endif
#endif /* PICCA_B200_H */
