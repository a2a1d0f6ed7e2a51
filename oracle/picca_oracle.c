/*
 * picca_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity referee, never the product path).
 *
 * Plain-C, scalar restatement of the reference's Numba pixel-pair kernels, statement by
 * statement and in the same floating-point evaluation order, so that on the same host it
 * reproduces the reference bit for bit (same libm sin/cos/pow/floor, no FMA contraction:
 * build with -ffp-contract=off, no -ffast-math).
 *
 *   orc_xi_auto_pair    <- reference py/picca/cf.py:250-387   (compute_xi_forest_pairs_fast)
 *   orc_xi_cross_forest <- reference py/picca/xcf.py:223-322  (compute_xi_forest_pairs_fast)
 *   orc_dmat_auto_pair  <- reference py/picca/cf.py:520-887   (compute_dmat_forest_pairs_fast)
 *   orc_dmat_cross_forest <- reference py/picca/xcf.py:427-674 (compute_dmat_forest_pairs_fast)
 *   orc_wick_t123_pair  <- reference py/picca/cf.py:1497-1626  (compute_wickT123_pairs)
 *   orc_wick_t1234_forest <- reference py/picca/xcf.py:1219-1351 (compute_wickT1234_pairs)
 *   orc_xi_auto_batch / orc_xi_cross_batch: the compute_xi loops (cf.py:138-247, xcf.py:126-220)
 *       over a CSR catalogue + neighbour list, one histogram row per HEALPix pixel, OpenMP over
 *       pixels -- the analogue of the reference's Pool.map over pixels (picca_cf.py:454-457).
 *       Used as the CPU baseline by bench.py and by tests at sizes too big for Python loops.
 *
 * Pinned against the live reference and its golden FITS files by tests/test_oracle_vs_reference.py
 * (run where /root/reference exists) and against tests/golden/ everywhere.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SPEED_LIGHT 299792.458 /* reference py/picca/constants.py: SPEED_LIGHT (km/s) */

/* Mirrors the module globals of picca.cf / picca.xcf (cf.py:28-79, xcf.py:27-68). */
typedef struct {
    int32_t num_bins_r_par;        /* np  */
    int32_t num_bins_r_trans;      /* nt  */
    int32_t num_model_bins_r_par;  /* npm */
    int32_t num_model_bins_r_trans;/* ntm */
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
} orc_params;

/* ------------------------------------------------------------------------------------------
 * cf.compute_xi_forest_pairs_fast, cf.py:250-387
 * ---------------------------------------------------------------------------------------- */
void orc_xi_auto_pair(const orc_params *P, int64_t n1, const double *z1, const double *r_comov1,
                      const double *dist_m1, const double *weights1, const double *delta1,
                      double z_qso_1, int64_t n2, const double *z2, const double *r_comov2,
                      const double *dist_m2, const double *weights2, const double *delta2,
                      double z_qso_2, double ang, int32_t same_half_plate, double *rebin_weight,
                      double *rebin_xi, double *rebin_r_par, double *rebin_r_trans, double *rebin_z,
                      int64_t *rebin_num_pairs)
{
    const double r_par_max = P->r_par_max, r_par_min = P->r_par_min, r_trans_max = P->r_trans_max;
    const int num_bins_r_par = P->num_bins_r_par, num_bins_r_trans = P->num_bins_r_trans;
    /* loop-invariant libm calls, hoisted exactly as LLVM hoists them in the Numba build: the
     * values are identical to evaluating np.cos(ang / 2) inside the loop (cf.py:356-357) */
    const double cos_half = cos(ang / 2), sin_half = sin(ang / 2);
    for (int64_t i = 0; i < n1; i++) {
        if (weights1[i] == 0) continue; /* cf.py:318 */

        if (P->has_zerr_cut && (ang < P->zerr_cut_deg * M_PI / 180.0)) { /* cf.py:321-328 */
            double z_qF = 0.5 * (z1[i] + z_qso_2);
            double dv_kms = fabs(z1[i] - z_qso_2) / (1 + z_qF);
            dv_kms *= ORC_SPEED_LIGHT;
            if (dv_kms < P->zerr_cut_kms) continue;
        }

        for (int64_t j = 0; j < n2; j++) {
            if (weights2[j] == 0) continue; /* cf.py:331 */

            double z = (z1[i] + z2[j]) / 2; /* cf.py:334 */

            if ((P->has_z_min_pairs && z < P->z_min_pairs) ||
                (P->has_z_max_pairs && z > P->z_max_pairs)) /* cf.py:336-339 */
                continue;

            if (P->has_zerr_cut && (ang < P->zerr_cut_deg * M_PI / 180.0)) { /* cf.py:341-348 */
                double z_qF = 0.5 * (z2[j] + z_qso_1);
                double dv_kms = fabs(z2[j] - z_qso_1) / (1 + z_qF);
                dv_kms *= ORC_SPEED_LIGHT;
                if (dv_kms < P->zerr_cut_kms) continue;
            }

            double r_par, r_trans;
            if (P->ang_correlation) { /* cf.py:350-354 */
                r_par = r_comov1[i] / r_comov2[j];
                if (!P->x_correlation && r_par < 1.0) r_par = 1.0 / r_par;
                r_trans = ang;
            } else { /* cf.py:356-362 */
                r_par = (r_comov1[i] - r_comov2[j]) * cos_half;
                r_trans = (dist_m1[i] + dist_m2[j]) * sin_half;
                if (P->rmu_binning) {
                    r_trans = sqrt(r_trans * r_trans + r_par * r_par);
                    r_par /= r_trans;
                }
                if (!P->x_correlation) r_par = fabs(r_par);
            }

            if (r_par >= r_par_max || r_trans >= r_trans_max || r_par < r_par_min) /* cf.py:364 */
                continue;

            double delta_times_weight1 = delta1[i] * weights1[i];
            double delta_times_weight2 = delta2[j] * weights2[j];
            double delta_times_weight12 = delta_times_weight1 * delta_times_weight2;
            double weights12 = weights1[i] * weights2[j];

            double bins_r_par =
                floor((r_par - r_par_min) / (r_par_max - r_par_min) * num_bins_r_par); /* :372 */
            double bins_r_trans = floor(r_trans / r_trans_max * num_bins_r_trans);     /* :375 */
            int64_t bins = (int64_t)(bins_r_trans + num_bins_r_trans * bins_r_par);    /* :376 */

            if (P->remove_same_half_plate_close_pairs && same_half_plate) { /* cf.py:378-380 */
                if (fabs(r_par) < (r_par_max - r_par_min) / num_bins_r_par) continue;
            }

            rebin_xi[bins] += delta_times_weight12; /* cf.py:382-387 */
            rebin_weight[bins] += weights12;
            rebin_r_par[bins] += r_par * weights12;
            rebin_r_trans[bins] += r_trans * weights12;
            rebin_z[bins] += z * weights12;
            rebin_num_pairs[bins] += 1;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * xcf.compute_xi_forest_pairs_fast, xcf.py:223-322  (forest x list of objects, ang is a vector)
 * ---------------------------------------------------------------------------------------- */
void orc_xi_cross_forest(const orc_params *P, int64_t n1, const double *z1, const double *r_comov1,
                         const double *dist_m1, const double *weights1, const double *delta1,
                         int64_t n2, const double *z2, const double *r_comov2,
                         const double *dist_m2, const double *weights2, const double *ang,
                         double *rebin_weight, double *rebin_xi, double *rebin_r_par,
                         double *rebin_r_trans, double *rebin_z, int64_t *rebin_num_pairs)
{
    const double r_par_max = P->r_par_max, r_par_min = P->r_par_min, r_trans_max = P->r_trans_max;
    const int num_bins_r_par = P->num_bins_r_par, num_bins_r_trans = P->num_bins_r_trans;
    for (int64_t i = 0; i < n1; i++) {
        if (weights1[i] == 0) continue; /* xcf.py:279 */
        for (int64_t j = 0; j < n2; j++) {
            if (weights2[j] == 0) continue; /* xcf.py:283 */

            double z = (z1[i] + z2[j]) / 2;

            if ((P->has_z_min_pairs && z < P->z_min_pairs) ||
                (P->has_z_max_pairs && z > P->z_max_pairs)) /* xcf.py:288-291 */
                continue;

            double r_par, r_trans;
            if (P->ang_correlation) { /* xcf.py:293-295 */
                r_par = r_comov1[i] / r_comov2[j];
                r_trans = ang[j];
            } else { /* xcf.py:297-298 */
                r_par = (r_comov1[i] - r_comov2[j]) * cos(ang[j] / 2);
                r_trans = (dist_m1[i] + dist_m2[j]) * sin(ang[j] / 2);
            }
            if (P->rmu_binning) { /* xcf.py:300-302 */
                r_trans = sqrt(r_trans * r_trans + r_par * r_par);
                r_par /= r_trans;
            }

            if (r_par >= r_par_max || r_trans >= r_trans_max || r_par <= r_par_min) /* :304 */
                continue;

            double delta_times_weight = delta1[i] * weights1[i] * weights2[j]; /* xcf.py:308 */
            double weights12 = weights1[i] * weights2[j];

            double bins_r_par =
                floor((r_par - r_par_min) / (r_par_max - r_par_min) * num_bins_r_par);
            double bins_r_trans = floor(r_trans / r_trans_max * num_bins_r_trans);
            int64_t bins = (int64_t)(bins_r_trans + num_bins_r_trans * bins_r_par);

            rebin_xi[bins] += delta_times_weight; /* xcf.py:317-322 */
            rebin_weight[bins] += weights12;
            rebin_r_par[bins] += r_par * weights12;
            rebin_r_trans[bins] += r_trans * weights12;
            rebin_z[bins] += z * weights12;
            rebin_num_pairs[bins] += 1;
        }
    }
}

static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* np.unique of the first n entries: sorted, duplicates removed; returns the new length. */
static int64_t unique_i32(int32_t *v, int64_t n)
{
    if (n == 0) return 0;
    qsort(v, (size_t)n, sizeof(int32_t), cmp_i32);
    int64_t m = 1;
    for (int64_t k = 1; k < n; k++)
        if (v[k] != v[m - 1]) v[m++] = v[k];
    return m;
}

/* ------------------------------------------------------------------------------------------
 * cf.compute_dmat_forest_pairs_fast, cf.py:520-887
 * Returns 0, or -1 for the reference's IndexError("negative bin index") (cf.py:855-856).
 *
 * Q8: with remove_same_half_plate_close_pairs the reference's pass 0 does not count
 * close same-half-plate pairs but pass 1 still records their model bin, writing past the end of
 * all_model_bins; np.unique then only sees the first num_pairs entries.  Restated here as
 * "record only while counter_of_pairs < num_pairs".
 * ---------------------------------------------------------------------------------------- */
int orc_dmat_auto_pair(const orc_params *P, int64_t n1, const double *log_lambda1,
                       const double *r_comov1, const double *dist_m1, const double *z1,
                       const double *weights1, double z_qso_1, int32_t order1, int64_t n2,
                       const double *log_lambda2, const double *r_comov2, const double *dist_m2,
                       const double *z2, const double *weights2, double z_qso_2, int32_t order2,
                       double ang, int32_t same_half_plate, double *weights_dmat, double *dmat,
                       double *r_par_eff, double *r_trans_eff, double *z_eff, double *weight_eff)
{
    const double r_par_max = P->r_par_max, r_par_min = P->r_par_min, r_trans_max = P->r_trans_max;
    const int num_bins_r_par = P->num_bins_r_par, num_bins_r_trans = P->num_bins_r_trans;
    const int num_model_bins_r_par = P->num_model_bins_r_par;
    const int num_model_bins_r_trans = P->num_model_bins_r_trans;
    const int64_t nbm = (int64_t)num_model_bins_r_par * num_model_bins_r_trans;
    const double z_ref = P->z_ref, alpha = P->alpha, alpha2 = P->alpha2;
    const double cos_half = cos(ang / 2), sin_half = sin(ang / 2); /* hoisted, same values */

    /* pass 0: count relevant pixel pairs, cf.py:547-571 */
    int64_t num_pairs = 0;
    for (int64_t i = 0; i < n1; i++) {
        if (weights1[i] == 0) continue;
        for (int64_t j = 0; j < n2; j++) {
            if (weights2[j] == 0) continue;
            double r_par = (r_comov1[i] - r_comov2[j]) * cos_half;
            double r_trans = (dist_m1[i] + dist_m2[j]) * sin_half;
            if (P->rmu_binning) {
                r_trans = sqrt(r_trans * r_trans + r_par * r_par);
                r_par /= r_trans;
            }
            if (!P->x_correlation) r_par = fabs(r_par);
            if (r_par >= r_par_max || r_trans >= r_trans_max || r_par < r_par_min) continue;
            if (P->remove_same_half_plate_close_pairs && same_half_plate) {
                if (fabs(r_par) < (r_par_max - r_par_min) / num_bins_r_par) continue;
            }
            num_pairs += 1;
        }
    }
    if (num_pairs == 0) return 0;

    /* cf.py:577-594 (numba's .sum() is a sequential left-to-right loop) */
    double sum_weights1 = 0, sum_weights2 = 0;
    for (int64_t i = 0; i < n1; i++) sum_weights1 += weights1[i];
    for (int64_t j = 0; j < n2; j++) sum_weights2 += weights2[j];
    double mean_log_lambda1 = 0, mean_log_lambda2 = 0;
    for (int64_t i = 0; i < n1; i++) mean_log_lambda1 += log_lambda1[i] * weights1[i];
    for (int64_t j = 0; j < n2; j++) mean_log_lambda2 += log_lambda2[j] * weights2[j];
    mean_log_lambda1 /= sum_weights1;
    mean_log_lambda2 /= sum_weights2;
    double *dll1 = (double *)malloc(sizeof(double) * (size_t)(n1 > 0 ? n1 : 1));
    double *dll2 = (double *)malloc(sizeof(double) * (size_t)(n2 > 0 ? n2 : 1));
    for (int64_t i = 0; i < n1; i++) dll1[i] = log_lambda1[i] - mean_log_lambda1;
    for (int64_t j = 0; j < n2; j++) dll2[j] = log_lambda2[j] - mean_log_lambda2;
    double swsll1 = 0, swsll2 = 0;
    for (int64_t i = 0; i < n1; i++) swsll1 += weights1[i] * (dll1[i] * dll1[i]);
    for (int64_t j = 0; j < n2; j++) swsll2 += weights2[j] * (dll2[j] * dll2[j]);

    const int64_t num_pixels1 = n1, num_pixels2 = n2;
    double *eta1 = (double *)calloc((size_t)(nbm * num_pixels1), sizeof(double)); /* cf.py:600 */
    double *eta2 = (double *)calloc((size_t)(nbm * num_pixels2), sizeof(double));
    double *eta3 = (double *)calloc((size_t)(nbm * num_pixels1), sizeof(double));
    double *eta4 = (double *)calloc((size_t)(nbm * num_pixels2), sizeof(double));
    double *eta5 = (double *)calloc((size_t)nbm, sizeof(double));
    double *eta6 = (double *)calloc((size_t)nbm, sizeof(double));
    double *eta7 = (double *)calloc((size_t)nbm, sizeof(double));
    double *eta8 = (double *)calloc((size_t)nbm, sizeof(double));

    int32_t *all_selected_data_bins = (int32_t *)malloc(sizeof(int32_t) * (size_t)num_pairs);
    int32_t *all_selected_model_bins = (int32_t *)calloc((size_t)num_pairs, sizeof(int32_t));
    int32_t *all_selected_i = (int32_t *)calloc((size_t)num_pairs, sizeof(int32_t));
    int32_t *all_selected_j = (int32_t *)calloc((size_t)num_pairs, sizeof(int32_t));
    int32_t *all_model_bins = (int32_t *)calloc((size_t)num_pairs, sizeof(int32_t));
    for (int64_t k = 0; k < num_pairs; k++) all_selected_data_bins[k] = -1; /* cf.py:612 */

    int64_t counter_of_selected_pairs = 0, counter_of_pairs = 0;
    for (int64_t i = 0; i < n1; i++) { /* pass 1, cf.py:623-843 */
        if (weights1[i] == 0) continue;
        int i_selected = 1;
        if (P->has_zerr_cut && (ang < P->zerr_cut_deg * M_PI / 180.0)) { /* cf.py:629-636 */
            double z_qF = 0.5 * (z1[i] + z_qso_2);
            double dv_kms = fabs(z1[i] - z_qso_2) / (1 + z_qF);
            dv_kms *= ORC_SPEED_LIGHT;
            if (dv_kms < P->zerr_cut_kms) i_selected = 0;
        }
        for (int64_t j = 0; j < n2; j++) {
            if (weights2[j] == 0) continue;
            double z = (z1[i] + z2[j]) / 2;
            int j_selected = 1;
            if ((P->has_z_min_pairs && z < P->z_min_pairs) ||
                (P->has_z_max_pairs && z > P->z_max_pairs))
                j_selected = 0; /* cf.py:646-649 */
            if (P->has_zerr_cut && (ang < P->zerr_cut_deg * M_PI / 180.0)) { /* cf.py:651-658 */
                double z_qF = 0.5 * (z2[j] + z_qso_1);
                double dv_kms = fabs(z2[j] - z_qso_1) / (1 + z_qF);
                dv_kms *= ORC_SPEED_LIGHT;
                if (dv_kms < P->zerr_cut_kms) j_selected = 0;
            }
            double r_par = (r_comov1[i] - r_comov2[j]) * cos_half;
            double r_trans = (dist_m1[i] + dist_m2[j]) * sin_half;
            if (P->rmu_binning) {
                r_trans = sqrt(r_trans * r_trans + r_par * r_par);
                r_par /= r_trans;
            }
            if (!P->x_correlation) r_par = fabs(r_par);
            if (r_par >= r_par_max || r_trans >= r_trans_max || r_par < r_par_min) continue;
            if (P->remove_same_half_plate_close_pairs && same_half_plate) { /* cf.py:669-671 */
                if (fabs(r_par) < (r_par_max - r_par_min) / num_bins_r_par) j_selected = 0;
            }

            double weights12 = weights1[i] * weights2[j];
            double zfac;
            if (P->redshift_evolution_in_distortion_matrix) /* cf.py:680-685 */
                zfac = pow((1 + z1[i]) / (1 + z_ref), alpha - 1) *
                       pow((1 + z2[j]) / (1 + z_ref), alpha2 - 1);
            else
                zfac = 1.0;

            double bins_r_par =
                floor((r_par - r_par_min) / (r_par_max - r_par_min) * num_bins_r_par);
            double bins_r_trans = floor(r_trans / r_trans_max * num_bins_r_trans);
            int32_t bins = (int32_t)(bins_r_trans + num_bins_r_trans * bins_r_par);
            double model_bins_r_par =
                floor((r_par - r_par_min) / (r_par_max - r_par_min) * num_model_bins_r_par);
            double model_bins_r_trans = floor(r_trans / r_trans_max * num_model_bins_r_trans);
            int32_t model_bins =
                (int32_t)(model_bins_r_trans + num_model_bins_r_trans * model_bins_r_par);

            if (counter_of_pairs < num_pairs) /* Q8: see header comment */
                all_model_bins[counter_of_pairs] = model_bins; /* cf.py:702 */
            counter_of_pairs += 1;

            if (i_selected && j_selected) { /* cf.py:705-718 */
                all_selected_model_bins[counter_of_selected_pairs] = model_bins;
                all_selected_data_bins[counter_of_selected_pairs] = bins;
                all_selected_i[counter_of_selected_pairs] = (int32_t)i;
                all_selected_j[counter_of_selected_pairs] = (int32_t)j;
                counter_of_selected_pairs += 1;

                r_par_eff[model_bins] += weights12 * r_par;
                r_trans_eff[model_bins] += weights12 * r_trans;
                z_eff[model_bins] += weights12 * z;
                weight_eff[model_bins] += weights12;
                weights_dmat[bins] += weights12;
            }

            eta1[i + num_pixels1 * model_bins] += zfac * weights2[j] / sum_weights2; /* :767 */
            eta2[j + num_pixels2 * model_bins] += zfac * weights1[i] / sum_weights1; /* :771 */
            eta5[model_bins] += zfac * weights12 / sum_weights1 / sum_weights2;      /* :775 */

            if (order2 == 1) { /* cf.py:777-802 */
                eta3[i + num_pixels1 * model_bins] += zfac * weights2[j] * dll2[j] / swsll2;
                eta6[model_bins] +=
                    zfac * weights1[i] / sum_weights1 * (weights2[j] * dll2[j] / swsll2);
            }
            if (order1 == 1) { /* cf.py:803-843 */
                eta4[j + num_pixels2 * model_bins] += zfac * weights1[i] * dll1[i] / swsll1;
                eta7[model_bins] +=
                    zfac * weights2[j] / sum_weights2 * (weights1[i] * dll1[i] / swsll1);
                if (order2 == 1)
                    eta8[model_bins] +=
                        zfac * weights1[i] * dll1[i] * weights2[j] * dll2[j] / swsll1 / swsll2;
            }
        }
    }

    /* cf.py:846-848 */
    int64_t num_unique = unique_i32(all_model_bins, num_pairs);

    int status = 0;
    for (int64_t pair = 0; pair < counter_of_selected_pairs; pair++) { /* cf.py:851-887 */
        int64_t i = all_selected_i[pair];
        int64_t j = all_selected_j[pair];
        int64_t bins = all_selected_data_bins[pair];
        if (bins < 0) {
            status = -1;
            break;
        }
        int64_t model_bins = all_selected_model_bins[pair];
        double weights12 = weights1[i] * weights2[j];
        int64_t dmat_bin = model_bins + nbm * bins;
        double zfac;
        if (P->redshift_evolution_in_distortion_matrix)
            zfac = pow((1 + z1[i]) / (1 + z_ref), alpha - 1) *
                   pow((1 + z2[j]) / (1 + z_ref), alpha2 - 1);
        else
            zfac = 1;
        dmat[dmat_bin] += weights12 * zfac;

        for (int64_t u = 0; u < num_unique; u++) {
            int64_t k = all_model_bins[u];
            dmat_bin = k + nbm * bins;
            dmat[dmat_bin] +=
                weights12 * (eta5[k] + eta6[k] * dll2[j] + eta7[k] * dll1[i] +
                             eta8[k] * dll1[i] * dll2[j] - eta1[i + num_pixels1 * k] -
                             eta2[j + num_pixels2 * k] - eta3[i + num_pixels1 * k] * dll2[j] -
                             eta4[j + num_pixels2 * k] * dll1[i]);
        }
    }

    free(dll1); free(dll2);
    free(eta1); free(eta2); free(eta3); free(eta4);
    free(eta5); free(eta6); free(eta7); free(eta8);
    free(all_selected_data_bins); free(all_selected_model_bins);
    free(all_selected_i); free(all_selected_j); free(all_model_bins);
    return status;
}

/* ------------------------------------------------------------------------------------------
 * xcf.compute_dmat_forest_pairs_fast, xcf.py:427-674 (one forest x its kept objects)
 * Pass 0 rejects r_par <= r_par_min (xcf.py:463-464) while pass 1 rejects r_par < r_par_min
 * (xcf.py:534-535); a pair with r_par == r_par_min exactly would overflow the reference's
 * scratch arrays -- restated with the same guard as Q8 (record only while in bounds).
 * ---------------------------------------------------------------------------------------- */
int orc_dmat_cross_forest(const orc_params *P, int64_t n1, const double *log_lambda1,
                          const double *r_comov1, const double *dist_m1, const double *z1,
                          const double *weights1, int32_t order1, int64_t n2,
                          const double *r_comov2, const double *dist_m2, const double *z2,
                          const double *weights2, const double *ang, double *weights_dmat,
                          double *dmat, double *r_par_eff, double *r_trans_eff, double *z_eff,
                          double *weight_eff)
{
    const double r_par_max = P->r_par_max, r_par_min = P->r_par_min, r_trans_max = P->r_trans_max;
    const int num_bins_r_par = P->num_bins_r_par, num_bins_r_trans = P->num_bins_r_trans;
    const int num_model_bins_r_par = P->num_model_bins_r_par;
    const int num_model_bins_r_trans = P->num_model_bins_r_trans;
    const int64_t nbm = (int64_t)num_model_bins_r_par * num_model_bins_r_trans;
    const double z_ref = P->z_ref, alpha = P->alpha, alpha_obj = P->alpha2;

    int64_t num_pairs = 0; /* xcf.py:449-469 */
    for (int64_t i = 0; i < n1; i++) {
        if (weights1[i] == 0) continue;
        for (int64_t j = 0; j < n2; j++) {
            if (weights2[j] == 0) continue;
            double r_par = (r_comov1[i] - r_comov2[j]) * cos(ang[j] / 2);
            double r_trans = (dist_m1[i] + dist_m2[j]) * sin(ang[j] / 2);
            if (P->rmu_binning) {
                r_trans = sqrt(r_trans * r_trans + r_par * r_par);
                r_par /= r_trans;
            }
            if (r_par >= r_par_max || r_trans >= r_trans_max || r_par <= r_par_min) continue;
            num_pairs += 1;
        }
    }
    if (num_pairs == 0) return 0;

    double sum_weights1 = 0; /* xcf.py:475-486 */
    for (int64_t i = 0; i < n1; i++) sum_weights1 += weights1[i];
    double mean_log_lambda1 = 0;
    for (int64_t i = 0; i < n1; i++) mean_log_lambda1 += log_lambda1[i] * weights1[i];
    mean_log_lambda1 /= sum_weights1;
    double *dll1 = (double *)malloc(sizeof(double) * (size_t)(n1 > 0 ? n1 : 1));
    for (int64_t i = 0; i < n1; i++) dll1[i] = log_lambda1[i] - mean_log_lambda1;
    double swsll1 = 0;
    for (int64_t i = 0; i < n1; i++) swsll1 += weights1[i] * (dll1[i] * dll1[i]);

    const int64_t num_pixels2 = n2;
    double *eta2 = (double *)calloc((size_t)(nbm * num_pixels2), sizeof(double));
    double *eta4 = (double *)calloc((size_t)(nbm * num_pixels2), sizeof(double));
    int32_t *all_selected_data_bins = (int32_t *)malloc(sizeof(int32_t) * (size_t)num_pairs);
    int32_t *all_selected_model_bins = (int32_t *)calloc((size_t)num_pairs, sizeof(int32_t));
    int32_t *all_selected_i = (int32_t *)calloc((size_t)num_pairs, sizeof(int32_t));
    int32_t *all_selected_j = (int32_t *)calloc((size_t)num_pairs, sizeof(int32_t));
    int32_t *all_model_bins = (int32_t *)calloc((size_t)num_pairs, sizeof(int32_t));
    for (int64_t k = 0; k < num_pairs; k++) all_selected_data_bins[k] = -1;

    int64_t counter_of_selected_pairs = 0, counter_of_pairs = 0;
    for (int64_t i = 0; i < n1; i++) { /* xcf.py:509-636 */
        if (weights1[i] == 0) continue;
        for (int64_t j = 0; j < n2; j++) {
            if (weights2[j] == 0) continue;
            double z = (z1[i] + z2[j]) / 2;
            int j_selected = 1;
            if ((P->has_z_min_pairs && z < P->z_min_pairs) ||
                (P->has_z_max_pairs && z > P->z_max_pairs))
                j_selected = 0;
            double r_par = (r_comov1[i] - r_comov2[j]) * cos(ang[j] / 2);
            double r_trans = (dist_m1[i] + dist_m2[j]) * sin(ang[j] / 2);
            if (P->rmu_binning) {
                r_trans = sqrt(r_trans * r_trans + r_par * r_par);
                r_par /= r_trans;
            }
            if (r_par >= r_par_max || r_trans >= r_trans_max || r_par < r_par_min) continue;

            double weights12 = weights1[i] * weights2[j];
            double zfac;
            if (P->redshift_evolution_in_distortion_matrix) /* xcf.py:544-549 */
                zfac = pow((1 + z1[i]) / (1 + z_ref), alpha - 1) *
                       pow((1 + z2[j]) / (1 + z_ref), alpha_obj - 1);
            else
                zfac = 1;

            double bins_r_par =
                floor((r_par - r_par_min) / (r_par_max - r_par_min) * num_bins_r_par);
            double bins_r_trans = floor(r_trans / r_trans_max * num_bins_r_trans);
            int32_t bins = (int32_t)(bins_r_trans + num_bins_r_trans * bins_r_par);
            double model_bins_r_par =
                floor((r_par - r_par_min) / (r_par_max - r_par_min) * num_model_bins_r_par);
            double model_bins_r_trans = floor(r_trans / r_trans_max * num_model_bins_r_trans);
            int32_t model_bins =
                (int32_t)(model_bins_r_trans + num_model_bins_r_trans * model_bins_r_par);

            if (counter_of_pairs < num_pairs) all_model_bins[counter_of_pairs] = model_bins;
            counter_of_pairs += 1;

            if (j_selected && counter_of_selected_pairs < num_pairs) { /* xcf.py:569-582 */
                all_selected_model_bins[counter_of_selected_pairs] = model_bins;
                all_selected_data_bins[counter_of_selected_pairs] = bins;
                all_selected_i[counter_of_selected_pairs] = (int32_t)i;
                all_selected_j[counter_of_selected_pairs] = (int32_t)j;
                counter_of_selected_pairs += 1;

                r_par_eff[model_bins] += weights12 * r_par;
                r_trans_eff[model_bins] += weights12 * r_trans;
                z_eff[model_bins] += weights12 * z;
                weight_eff[model_bins] += weights12;
                weights_dmat[bins] += weights12;
            }

            eta2[j + num_pixels2 * model_bins] += zfac * weights1[i] / sum_weights1; /* :625 */
            if (order1 == 1) /* xcf.py:627-636 */
                eta4[j + num_pixels2 * model_bins] += zfac * (weights1[i] * dll1[i] / swsll1);
        }
    }

    int64_t num_unique = unique_i32(all_model_bins, num_pairs); /* xcf.py:639-641 */

    int status = 0;
    for (int64_t pair = 0; pair < counter_of_selected_pairs; pair++) { /* xcf.py:644-674 */
        int64_t i = all_selected_i[pair];
        int64_t j = all_selected_j[pair];
        int64_t bins = all_selected_data_bins[pair];
        if (bins < 0) {
            status = -1;
            break;
        }
        int64_t model_bins = all_selected_model_bins[pair];
        double weights12 = weights1[i] * weights2[j];
        int64_t dmat_bin = model_bins + nbm * bins;
        double zfac;
        if (P->redshift_evolution_in_distortion_matrix)
            zfac = pow((1 + z1[i]) / (1 + z_ref), alpha - 1) *
                   pow((1 + z2[j]) / (1 + z_ref), alpha_obj - 1);
        else
            zfac = 1;
        dmat[dmat_bin] += zfac * weights12;
        for (int64_t u = 0; u < num_unique; u++) {
            int64_t k = all_model_bins[u];
            dmat_bin = k + nbm * bins;
            dmat[dmat_bin] +=
                weights12 * (-eta2[j + num_pixels2 * k] - eta4[j + num_pixels2 * k] * dll1[i]);
        }
    }
    free(dll1); free(eta2); free(eta4);
    free(all_selected_data_bins); free(all_selected_model_bins);
    free(all_selected_i); free(all_selected_j); free(all_model_bins);
    return status;
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers over a packed catalogue (same SoA/CSR layout the product uses):
 *   pixel arrays z, r_comov, dist_m, weights, delta of length offset[n_forest];
 *   forest f owns pixels [offset[f], offset[f+1]).
 * The neighbour list is CSR: the k-th listed forest f1_index[k] has neighbours
 *   nb_index[nb_offset[k] .. nb_offset[k+1]) with angles nb_ang[...] (and flags).
 * Its output row is out_row[k] (its HEALPix row; listed forests are grouped by ascending row);
 * rows are un-normalised sums laid out [n_rows][6][nb]: weight, xi, r_par, r_trans, z (double)
 * and num_pairs (int64 bits).  num_threads = 1 reproduces the reference's summation order; more
 * threads (pthreads) split the forests dynamically (timing baseline, cf. the reference's Pool.map
 * over HEALPix pixels, picca_cf.py:454-457).
 * ---------------------------------------------------------------------------------------- */
#include <pthread.h>

typedef struct {
    const orc_params *P;
    const int64_t *offset1; const double *z1, *rc1, *dm1, *w1, *d1, *zq1;
    const int64_t *offset2; const double *z2, *rc2, *dm2, *w2, *d2, *zq2;
    const int64_t *f1_index, *nb_offset, *nb_index; const double *nb_ang;
    const int32_t *nb_same_half_plate;
    const int64_t *out_row; int64_t n_f1; int64_t n_rows; double *out;
    int64_t next_row; int cross;
} orc_batch;

static void orc_batch_forest(orc_batch *B, int64_t k, double *base)
{
    const orc_params *P = B->P;
    const int64_t nb = (int64_t)P->num_bins_r_par * P->num_bins_r_trans;
    int64_t f1 = B->f1_index[k];
    int64_t a = B->offset1[f1], n1 = B->offset1[f1 + 1] - a;
    if (!B->cross) {
        for (int64_t e = B->nb_offset[k]; e < B->nb_offset[k + 1]; e++) {
            int64_t f2 = B->nb_index[e];
            int64_t b = B->offset2[f2], n2 = B->offset2[f2 + 1] - b;
            orc_xi_auto_pair(P, n1, B->z1 + a, B->rc1 + a, B->dm1 + a, B->w1 + a, B->d1 + a,
                             B->zq1[f1], n2, B->z2 + b, B->rc2 + b, B->dm2 + b, B->w2 + b,
                             B->d2 + b, B->zq2[f2], B->nb_ang[e],
                             B->nb_same_half_plate ? B->nb_same_half_plate[e] : 0,
                             base + 0 * nb, base + 1 * nb, base + 2 * nb, base + 3 * nb,
                             base + 4 * nb, (int64_t *)(base + 5 * nb));
        }
    } else {
        int64_t m = B->nb_offset[k + 1] - B->nb_offset[k];
        if (m == 0) return; /* xcf.py:157 */
        double *gz = (double *)malloc(sizeof(double) * (size_t)m * 4);
        double *grc = gz + m, *gdm = gz + 2 * m, *gw = gz + 3 * m;
        for (int64_t e = 0; e < m; e++) { /* the gathers of xcf.py:159-185 */
            int64_t q = B->nb_index[B->nb_offset[k] + e];
            gz[e] = B->z2[q]; grc[e] = B->rc2[q]; gdm[e] = B->dm2[q]; gw[e] = B->w2[q];
        }
        orc_xi_cross_forest(P, n1, B->z1 + a, B->rc1 + a, B->dm1 + a, B->w1 + a, B->d1 + a, m,
                            gz, grc, gdm, gw, B->nb_ang + B->nb_offset[k], base + 0 * nb,
                            base + 1 * nb, base + 2 * nb, base + 3 * nb, base + 4 * nb,
                            (int64_t *)(base + 5 * nb));
        free(gz);
    }
}

/* One thread: rows in order, forests in order, accumulating straight into the output -- the
 * reference's sequential summation order (used by the parity tests). */
static void orc_batch_serial(orc_batch *B, int64_t n_f1)
{
    const int64_t nb = (int64_t)B->P->num_bins_r_par * B->P->num_bins_r_trans;
    for (int64_t k = 0; k < n_f1; k++) orc_batch_forest(B, k, B->out + B->out_row[k] * 6 * nb);
}

/* Several threads: forests are handed out dynamically (the reference forks workers over HEALPix
 * pixels; finer units keep every core busy on a bounded sample); each thread accumulates into a
 * private block and merges it under a lock when its row changes.  Counts are exact; the order of
 * fp64 additions differs from the serial run at the 1e-16 level.  Timing baseline only. */
static pthread_mutex_t orc_merge_lock = PTHREAD_MUTEX_INITIALIZER;

static void orc_merge(orc_batch *B, int64_t row, double *priv)
{
    const int64_t nb = (int64_t)B->P->num_bins_r_par * B->P->num_bins_r_trans;
    double *dst = B->out + row * 6 * nb;
    pthread_mutex_lock(&orc_merge_lock);
    for (int64_t x = 0; x < 5 * nb; x++) dst[x] += priv[x];
    int64_t *dc = (int64_t *)(dst + 5 * nb), *pc = (int64_t *)(priv + 5 * nb);
    for (int64_t x = 0; x < nb; x++) dc[x] += pc[x];
    pthread_mutex_unlock(&orc_merge_lock);
    memset(priv, 0, sizeof(double) * (size_t)(6 * nb));
}

static void *orc_batch_worker(void *arg)
{
    orc_batch *B = (orc_batch *)arg;
    const int64_t nb = (int64_t)B->P->num_bins_r_par * B->P->num_bins_r_trans;
    double *priv = (double *)calloc((size_t)(6 * nb), sizeof(double));
    int64_t cur = -1;
    for (;;) {
        int64_t k = __atomic_fetch_add(&B->next_row, 1, __ATOMIC_RELAXED);
        if (k >= B->n_f1) break;
        if (B->out_row[k] != cur) {
            if (cur >= 0) orc_merge(B, cur, priv);
            cur = B->out_row[k];
        }
        orc_batch_forest(B, k, priv);
    }
    if (cur >= 0) orc_merge(B, cur, priv);
    free(priv);
    return NULL;
}

static void orc_batch_run(orc_batch *B, int64_t n_f1, const int64_t *out_row, int32_t num_threads)
{
    B->out_row = out_row;
    B->n_f1 = n_f1;
    B->next_row = 0;
    if (num_threads <= 1) {
        orc_batch_serial(B, n_f1);
        return;
    }
    if (num_threads > 256) num_threads = 256;
    pthread_t th[256];
    for (int t = 1; t < num_threads; t++) pthread_create(&th[t], NULL, orc_batch_worker, B);
    orc_batch_worker(B);
    for (int t = 1; t < num_threads; t++) pthread_join(th[t], NULL);
}

void orc_xi_auto_batch(const orc_params *P, const int64_t *offset1, const double *z1,
                       const double *rc1, const double *dm1, const double *w1, const double *d1,
                       const double *zq1, const int64_t *offset2, const double *z2,
                       const double *rc2, const double *dm2, const double *w2, const double *d2,
                       const double *zq2, int64_t n_f1, const int64_t *f1_index,
                       const int64_t *out_row, const int64_t *nb_offset, const int64_t *nb_index,
                       const double *nb_ang, const int32_t *nb_same_half_plate, int64_t n_rows,
                       double *out, int32_t num_threads)
{
    orc_batch B;
    memset(&B, 0, sizeof(B));
    B.P = P; B.offset1 = offset1; B.z1 = z1; B.rc1 = rc1; B.dm1 = dm1; B.w1 = w1; B.d1 = d1;
    B.zq1 = zq1; B.offset2 = offset2; B.z2 = z2; B.rc2 = rc2; B.dm2 = dm2; B.w2 = w2; B.d2 = d2;
    B.zq2 = zq2; B.f1_index = f1_index; B.nb_offset = nb_offset; B.nb_index = nb_index;
    B.nb_ang = nb_ang; B.nb_same_half_plate = nb_same_half_plate; B.n_rows = n_rows; B.out = out;
    B.cross = 0;
    orc_batch_run(&B, n_f1, out_row, num_threads);
}

/* objects: z2 = z_qso, rc2 = r_comov, dm2 = dist_m, w2 = weights, one entry per object */
void orc_xi_cross_batch(const orc_params *P, const int64_t *offset1, const double *z1,
                        const double *rc1, const double *dm1, const double *w1, const double *d1,
                        const double *obj_z, const double *obj_rc, const double *obj_dm,
                        const double *obj_w, int64_t n_f1, const int64_t *f1_index,
                        const int64_t *out_row, const int64_t *nb_offset,
                        const int64_t *nb_index, const double *nb_ang, int64_t n_rows,
                        double *out, int32_t num_threads)
{
    orc_batch B;
    memset(&B, 0, sizeof(B));
    B.P = P; B.offset1 = offset1; B.z1 = z1; B.rc1 = rc1; B.dm1 = dm1; B.w1 = w1; B.d1 = d1;
    B.z2 = obj_z; B.rc2 = obj_rc; B.dm2 = obj_dm; B.w2 = obj_w; B.f1_index = f1_index;
    B.nb_offset = nb_offset; B.nb_index = nb_index; B.nb_ang = nb_ang; B.n_rows = n_rows;
    B.out = out; B.cross = 1;
    orc_batch_run(&B, n_f1, out_row, num_threads);
}

/* ------------------------------------------------------------------------------------------
 * cf.compute_wickT123_pairs, cf.py:1497-1626 (statement by statement; z_weight_evol with libm
 * pow as Numba's `**` lowers to).  weighted_xi_1d_1 is [n1][n1], weighted_xi_1d_2 [n2][n2]
 * (cf.py:1413-1421, built by the caller with NumPy as the reference does); t1, t2, t3 [nb][nb].
 * ---------------------------------------------------------------------------------------- */
void orc_wick_t123_pair(const orc_params *P, int64_t num_pixels1, const double *r_comov1,
                        int64_t num_pixels2, const double *r_comov2, double ang,
                        const double *weights1, const double *weights2, const double *z1,
                        const double *z2, const double *weighted_xi_1d_1,
                        const double *weighted_xi_1d_2, double *weights_wick,
                        int64_t *num_pairs_wick, double *t1, double *t2, double *t3)
{
    const double r_par_max = P->r_par_max, r_par_min = P->r_par_min, r_trans_max = P->r_trans_max;
    const int num_bins_r_par = P->num_bins_r_par, num_bins_r_trans = P->num_bins_r_trans;
    const int64_t nb = (int64_t)num_bins_r_par * num_bins_r_trans;
    double *z_weight_evol1 = (double *)malloc(sizeof(double) * (size_t)(num_pixels1 + 1));
    double *z_weight_evol2 = (double *)malloc(sizeof(double) * (size_t)(num_pixels2 + 1));
    for (int64_t i = 0; i < num_pixels1; i++) /* cf.py:1556 */
        z_weight_evol1[i] = pow((1 + z1[i]) / (1 + P->z_ref), P->alpha - 1);
    for (int64_t j = 0; j < num_pixels2; j++) /* cf.py:1557 */
        z_weight_evol2[j] = pow((1 + z2[j]) / (1 + P->z_ref), P->alpha2 - 1);
    const double cos_half = cos(ang / 2), sin_half = sin(ang / 2);

    int64_t wsum = 0; /* cf.py:1560-1572 */
    for (int64_t ind2 = 0; ind2 < num_pixels2; ind2++)
        for (int64_t ind1 = 0; ind1 < num_pixels1; ind1++) {
            double r_par = (r_comov1[ind1] - r_comov2[ind2]) * cos_half;
            if (!P->x_correlation) r_par = fabs(r_par);
            double r_trans = (r_comov1[ind1] + r_comov2[ind2]) * sin_half;
            if ((r_par < r_par_max) && (r_trans < r_trans_max) && (r_par >= r_par_min)) wsum += 1;
        }
    if (wsum == 0) { /* cf.py:1573-1574 */
        free(z_weight_evol1);
        free(z_weight_evol2);
        return;
    }
    int64_t *bins = (int64_t *)malloc(sizeof(int64_t) * (size_t)wsum);
    int64_t *bins_forest = (int64_t *)malloc(sizeof(int64_t) * (size_t)wsum);
    double *weights12 = (double *)malloc(sizeof(double) * (size_t)wsum);
    double *weight1 = (double *)malloc(sizeof(double) * (size_t)wsum);
    double *weight2 = (double *)malloc(sizeof(double) * (size_t)wsum);
    double *z_weight_evol = (double *)malloc(sizeof(double) * (size_t)wsum);
    int64_t ind = 0; /* cf.py:1583-1596 */
    for (int64_t ind2 = 0; ind2 < num_pixels2; ind2++)
        for (int64_t ind1 = 0; ind1 < num_pixels1; ind1++) {
            double r_par = (r_comov1[ind1] - r_comov2[ind2]) * cos_half;
            if (!P->x_correlation) r_par = fabs(r_par);
            double r_trans = (r_comov1[ind1] + r_comov2[ind2]) * sin_half;
            if (!((r_par < r_par_max) && (r_trans < r_trans_max) && (r_par >= r_par_min))) continue;
            bins[ind] = ind1 + num_pixels1 * ind2;
            int64_t bin_r_par =
                (int64_t)((r_par - r_par_min) / (r_par_max - r_par_min) * num_bins_r_par);
            int64_t bin_r_trans = (int64_t)(r_trans / r_trans_max * num_bins_r_trans);
            bins_forest[ind] = bin_r_trans + num_bins_r_trans * bin_r_par;
            weights12[ind] = weights1[ind1] * weights2[ind2];
            weight1[ind] = weights1[ind1];
            weight2[ind] = weights2[ind2];
            z_weight_evol[ind] = z_weight_evol1[ind1] * z_weight_evol2[ind2];
            ind += 1;
        }
    for (int64_t index1 = 0; index1 < wsum; index1++) { /* cf.py:1598-1624 */
        const int64_t p1 = bins_forest[index1];
        const int64_t i1 = bins[index1] % num_pixels1;
        const int64_t j1 = (bins[index1] - i1) / num_pixels1;
        weights_wick[p1] += weights12[index1];
        num_pairs_wick[p1] += 1;
        t1[p1 * nb + p1] += weights12[index1] * z_weight_evol[index1];
        for (int64_t index2 = index1 + 1; index2 < wsum; index2++) {
            const int64_t p2 = bins_forest[index2];
            const int64_t i2 = bins[index2] % num_pixels1;
            const int64_t j2 = (bins[index2] - i2) / num_pixels1;
            if (i1 == i2) {
                const double prod =
                    weighted_xi_1d_2[j1 * num_pixels2 + j2] * weight1[index1] * z_weight_evol1[i1];
                t2[p1 * nb + p2] += prod;
                t2[p2 * nb + p1] += prod;
            } else if (j1 == j2) {
                const double prod =
                    weighted_xi_1d_1[i1 * num_pixels1 + i2] * weight2[index2] * z_weight_evol2[j1];
                t2[p1 * nb + p2] += prod;
                t2[p2 * nb + p1] += prod;
            } else {
                const double prod = weighted_xi_1d_1[i1 * num_pixels1 + i2] *
                                    weighted_xi_1d_2[j1 * num_pixels2 + j2];
                t3[p1 * nb + p2] += prod;
                t3[p2 * nb + p1] += prod;
            }
        }
    }
    free(bins); free(bins_forest); free(weights12); free(weight1); free(weight2);
    free(z_weight_evol); free(z_weight_evol1); free(z_weight_evol2);
}

/* ------------------------------------------------------------------------------------------
 * xcf.compute_wickT1234_pairs, xcf.py:1219-1351: one forest against its neighbouring objects
 * (ang, r_comov2, z2, weights2 are arrays over the objects).  t1..t4 [nb][nb].
 * ---------------------------------------------------------------------------------------- */
void orc_wick_t1234_forest(const orc_params *P, int64_t num_pixels1, const double *r_comov1,
                           const double *z1, const double *weights1,
                           const double *weighted_xi_1d_1, int64_t num_pixels2, const double *ang,
                           const double *r_comov2, const double *z2, const double *weights2,
                           double *weights_wick, int64_t *num_pairs_wick, double *t1, double *t2,
                           double *t3, double *t4)
{
    const double r_par_max = P->r_par_max, r_par_min = P->r_par_min, r_trans_max = P->r_trans_max;
    const int num_bins_r_par = P->num_bins_r_par, num_bins_r_trans = P->num_bins_r_trans;
    const int64_t nb = (int64_t)num_bins_r_par * num_bins_r_trans;
    double *z_weight_evol1 = (double *)malloc(sizeof(double) * (size_t)(num_pixels1 + 1));
    double *z_weight_evol2 = (double *)malloc(sizeof(double) * (size_t)(num_pixels2 + 1));
    for (int64_t i = 0; i < num_pixels1; i++) /* xcf.py:1271 */
        z_weight_evol1[i] = pow((1 + z1[i]) / (1 + P->z_ref), P->alpha - 1);
    for (int64_t j = 0; j < num_pixels2; j++) /* xcf.py:1272: alpha_obj */
        z_weight_evol2[j] = pow((1 + z2[j]) / (1 + P->z_ref), P->alpha2 - 1);

    int64_t wsum = 0; /* xcf.py:1275-1286 */
    for (int64_t ind2 = 0; ind2 < num_pixels2; ind2++)
        for (int64_t ind1 = 0; ind1 < num_pixels1; ind1++) {
            double r_par = (r_comov1[ind1] - r_comov2[ind2]) * cos(ang[ind2] / 2);
            double r_trans = (r_comov1[ind1] + r_comov2[ind2]) * sin(ang[ind2] / 2);
            if ((r_par < r_par_max) && (r_trans < r_trans_max) && (r_par >= r_par_min)) wsum += 1;
        }
    if (wsum == 0) {
        free(z_weight_evol1);
        free(z_weight_evol2);
        return;
    }
    int64_t *bins_forest = (int64_t *)malloc(sizeof(int64_t) * (size_t)wsum);
    double *weights12 = (double *)malloc(sizeof(double) * (size_t)wsum);
    double *weight1 = (double *)malloc(sizeof(double) * (size_t)wsum);
    int64_t *index_obj = (int64_t *)malloc(sizeof(int64_t) * (size_t)wsum);
    int64_t *index_delta = (int64_t *)malloc(sizeof(int64_t) * (size_t)wsum);
    int64_t ind = 0; /* xcf.py:1295-1311 */
    for (int64_t ind2 = 0; ind2 < num_pixels2; ind2++)
        for (int64_t ind1 = 0; ind1 < num_pixels1; ind1++) {
            double r_par = (r_comov1[ind1] - r_comov2[ind2]) * cos(ang[ind2] / 2);
            double r_trans = (r_comov1[ind1] + r_comov2[ind2]) * sin(ang[ind2] / 2);
            if (!((r_par < r_par_max) && (r_trans < r_trans_max) && (r_par >= r_par_min))) continue;
            int64_t bin_r_par =
                (int64_t)((r_par - r_par_min) / (r_par_max - r_par_min) * num_bins_r_par);
            int64_t bin_r_trans = (int64_t)(r_trans / r_trans_max * num_bins_r_trans);
            bins_forest[ind] = bin_r_trans + num_bins_r_trans * bin_r_par;
            weights12[ind] = weights1[ind1] * weights2[ind2];
            weight1[ind] = weights1[ind1];
            index_delta[ind] = ind1;
            index_obj[ind] = ind2;
            ind += 1;
        }
    for (int64_t index1 = 0; index1 < wsum; index1++) { /* xcf.py:1313-1342 */
        const int64_t p1 = bins_forest[index1];
        const int64_t i1 = index_delta[index1], j1 = index_obj[index1];
        weights_wick[p1] += weights12[index1];
        num_pairs_wick[p1] += 1;
        t1[p1 * nb + p1] +=
            weights12[index1] * weights12[index1] / weight1[index1] * z_weight_evol1[i1];
        for (int64_t index2 = index1 + 1; index2 < wsum; index2++) {
            const int64_t p2 = bins_forest[index2];
            const int64_t i2 = index_delta[index2], j2 = index_obj[index2];
            if (j1 == j2) {
                const double prod = weighted_xi_1d_1[i1 * num_pixels1 + i2] *
                                    (z_weight_evol2[j1] * z_weight_evol2[j1]);
                t2[p1 * nb + p2] += prod;
                t2[p2 * nb + p1] += prod;
            } else if (i1 == i2) {
                const double prod =
                    weights12[index1] * weights12[index2] / weight1[index1] * z_weight_evol1[i1];
                t3[p1 * nb + p2] += prod;
                t3[p2 * nb + p1] += prod;
            } else {
                const double prod = weighted_xi_1d_1[i1 * num_pixels1 + i2] * z_weight_evol2[j1] *
                                    z_weight_evol2[j2];
                t4[p1 * nb + p2] += prod;
                t4[p2 * nb + p1] += prod;
            }
        }
    }
    free(bins_forest); free(weights12); free(weight1); free(index_obj); free(index_delta);
    free(z_weight_evol1); free(z_weight_evol2);
}

int32_t orc_sizeof_params(void) { return (int32_t)sizeof(orc_params); }
