// Helpers shared by the distortion-matrix kernels (pb2_dmat.cu: register-tiled contraction over
// a dense per-CTA scratch, kept for rmu binning; pb2_dmat_run.cu: the product kernel for the
// standard binning): exact pixel-pair geometry and bins (reference py/picca/cf.py:660-700),
// conservative row / column windows, per-launch work description.
#pragma once
#include "pb2_common.cuh"

#define DM_THREADS 256
#define DM_TILE 64
#define DM_KCH 32
#define DM_CAP 512  // compact bins handled per chunk (padded to DM_TILE)
#define DM_MAXCH 256  // K chunks of rows with a column range (more: no tile skipping)

struct DmatGeom {
    bool in;       // inside the model range (cf.py:667)
    bool close;    // same-half-plate close pair (cf.py:669-671)
    int A, B;      // data bin, model bin
    double rp, rt;
};

// bin constants n / range * (1 -+ 2^-40) of the data and model grids, for the division-free proof
// of a pixel pair's bins (same sandwich as the xi kernels: both round-down products against
// 2^52 + 2^51 must agree, otherwise the reference expression below decides)
struct DmatFast {
    double kp_lo, kp_hi, kt_lo, kt_hi;  // data grid
    double mp_lo, mp_hi, mt_lo, mt_hi;  // model grid
    int same;                           // model grid == data grid (coefficient 1)
};
#define DM_MAGIC 6755399441055744.0  // 2^52 + 2^51

// evaluation of one pixel pair for the distortion matrix (cf.py:660-700 / xcf.py:528-564): bins
// proven by the sandwich, else the reference expression with IEEE divisions
__device__ __forceinline__ DmatGeom dmat_pair(const pb2_params &P, const DmatFast &F, double rc1,
                                              double dm1, double rc2, double dm2, double ch,
                                              double sh, bool cross_obj, bool shp)
{
    DmatGeom g;
    g.in = false;
    g.close = false;
    g.A = g.B = -1;
    double r_par = mul_rn(sub_rn(rc1, rc2), ch);
    double r_trans = mul_rn(add_rn(dm1, dm2), sh);
    if (P.rmu_binning) {
        r_trans = sqrt(add_rn(mul_rn(r_trans, r_trans), mul_rn(r_par, r_par)));
        r_par = div_rn(r_par, r_trans);
    }
    if (!cross_obj && !P.x_correlation) r_par = fabs(r_par);
    g.rp = r_par;
    g.rt = r_trans;
    if (r_par >= P.r_par_max || r_trans >= P.r_trans_max || r_par < P.r_par_min) return g;
    const double span = sub_rn(P.r_par_max, P.r_par_min);
    if (shp && fabs(r_par) < div_rn(span, (double)P.num_bins_r_par)) g.close = true;
    {
        const double x = sub_rn(r_par, P.r_par_min);
        const int bpl = __double2loint(__fma_rd(x, F.kp_lo, DM_MAGIC));
        const int bph = __double2loint(__fma_rd(x, F.kp_hi, DM_MAGIC));
        const int btl = __double2loint(__fma_rd(r_trans, F.kt_lo, DM_MAGIC));
        const int bth = __double2loint(__fma_rd(r_trans, F.kt_hi, DM_MAGIC));
        int mpl = bpl, mph = bph, mtl = btl, mth = bth;
        if (!F.same) {
            mpl = __double2loint(__fma_rd(x, F.mp_lo, DM_MAGIC));
            mph = __double2loint(__fma_rd(x, F.mp_hi, DM_MAGIC));
            mtl = __double2loint(__fma_rd(r_trans, F.mt_lo, DM_MAGIC));
            mth = __double2loint(__fma_rd(r_trans, F.mt_hi, DM_MAGIC));
        }
        if (bpl == bph && btl == bth && mpl == mph && mtl == mth &&
            (unsigned)bpl < (unsigned)P.num_bins_r_par && (unsigned)btl < (unsigned)P.num_bins_r_trans &&
            (unsigned)mpl < (unsigned)P.num_model_bins_r_par &&
            (unsigned)mtl < (unsigned)P.num_model_bins_r_trans) {
            g.in = true;
            g.A = btl + P.num_bins_r_trans * bpl;
            g.B = mtl + P.num_model_bins_r_trans * mpl;
            return g;
        }
    }
    const double fp = div_rn(sub_rn(r_par, P.r_par_min), span);
    const double ft = div_rn(r_trans, P.r_trans_max);
    const double bp = floor(mul_rn(fp, (double)P.num_bins_r_par));
    const double bt = floor(mul_rn(ft, (double)P.num_bins_r_trans));
    const double mp = floor(mul_rn(fp, (double)P.num_model_bins_r_par));
    const double mt = floor(mul_rn(ft, (double)P.num_model_bins_r_trans));
    const long long A = (long long)add_rn(bt, mul_rn((double)P.num_bins_r_trans, bp));
    const long long B = (long long)add_rn(mt, mul_rn((double)P.num_model_bins_r_trans, mp));
    const long long nb = (long long)P.num_bins_r_par * P.num_bins_r_trans;
    const long long nbm = (long long)P.num_model_bins_r_par * P.num_model_bins_r_trans;
    if (A < 0 || A >= nb || B < 0 || B >= nbm) return g;  // the reference would index out of bounds
    g.in = true;
    g.A = (int)A;
    g.B = (int)B;
    return g;
}

__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}

__device__ __forceinline__ double block_sum(double v, double *red)
{
    __syncthreads();
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.;
    for (int w = 0; w < DM_THREADS / 32; w++) t += red[w];
    return t;
}

// conservative column window of a row (sorted forests, standard binning); full range otherwise
__device__ __forceinline__ void row_window(const pb2_params &P, bool windows, double rc_i,
                                           double dm_i, const double *rc2, const double *dm2,
                                           int n2, double ch, double sh, bool signed_rp, int &lo,
                                           int &hi)
{
    lo = 0;
    hi = n2;
    if (!windows) return;
    const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
    const double dmax = P.r_par_max * inv_c * (1. + 1e-9) + 1e-9;
    const double dmin = P.r_par_min * inv_c;
    const double dlow = signed_rp ? (dmin - fabs(dmin) * 1e-9 - 1e-9) : -dmax;
    // rc2 > rc_i - dmax, rc2 < rc_i - dlow, dm2 < r_trans_max/sh - dm_i
    int a = 0, b = n2;
    const double v0 = rc_i - dmax;
    while (a < b) {
        const int m = (a + b) >> 1;
        if (rc2[m] <= v0) a = m + 1; else b = m;
    }
    lo = a;
    a = lo;
    b = n2;
    const double v1 = rc_i - dlow;
    while (a < b) {
        const int m = (a + b) >> 1;
        if (rc2[m] < v1) a = m + 1; else b = m;
    }
    hi = a;
    const double tsum = P.r_trans_max * inv_s * (1. + 1e-9) + 1e-9;
    if (isfinite(tsum)) {
        a = lo;
        b = hi;
        const double v2 = tsum - dm_i;
        while (a < b) {
            const int m = (a + b) >> 1;
            if (dm2[m] < v2) a = m + 1; else b = m;
        }
        hi = a;
    }
}

// conservative row window of a column: rows i with rc1[i] - rc_j in [dlow, dmax] and
// dm1[i] + dm_j below the r_trans limit (the mirror image of row_window)
__device__ __forceinline__ void col_window(const pb2_params &P, bool windows, double rc_j,
                                           double dm_j, const double *rc1, const double *dm1,
                                           int n1, double ch, double sh, bool signed_rp, int &lo,
                                           int &hi)
{
    lo = 0;
    hi = n1;
    if (!windows) return;
    const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
    const double dmax = P.r_par_max * inv_c * (1. + 1e-9) + 1e-9;
    const double dmin = P.r_par_min * inv_c;
    const double dlow = signed_rp ? (dmin - fabs(dmin) * 1e-9 - 1e-9) : -dmax;
    // rc1 >= rc_j + dlow, rc1 <= rc_j + dmax, dm1 < r_trans_max/sh - dm_j
    int a = 0, b = n1;
    const double v0 = rc_j + dlow;
    while (a < b) {
        const int m = (a + b) >> 1;
        if (rc1[m] < v0) a = m + 1; else b = m;
    }
    lo = a;
    b = n1;
    const double v1 = rc_j + dmax;
    while (a < b) {
        const int m = (a + b) >> 1;
        if (rc1[m] <= v1) a = m + 1; else b = m;
    }
    hi = a;
    const double tsum = P.r_trans_max * inv_s * (1. + 1e-9) + 1e-9;
    if (isfinite(tsum)) {
        a = lo;
        b = hi;
        const double v2 = tsum - dm_j;
        while (a < b) {
            const int m = (a + b) >> 1;
            if (dm1[m] < v2) a = m + 1; else b = m;
        }
        hi = a;
    }
}

struct DmatWork {
    long long *kept;            // kept pair indices
    unsigned long long *count;  // [0] number of kept pairs, [1] claim counter
    double *stats;              // [0] as-written FP64 ops of the reference algorithm (SURVEY 8d:
                                // N_sel (15 U + 4) + 40 N_inrange per forest pair), [1] sum of U,
                                // [2] in-range pixel pairs -- measurement only
    char *cta_base;             // per-CTA scratch
    long long cta_stride;
    int rows_max;               // 2*max_pix1 + 2*max_pix2 + 4
    int cap;                    // DM_CAP
    DmatFast fast;
    // per pixel / per line of sight constants of cf.py:577-594, 680-685 (filled by dmat_prologue)
    double *fz1, *dl1, *fz2, *dl2;  // ((1+z)/(1+z_ref))^(alpha-1), log_lambda - <log_lambda>_w
    double2 *fs1, *fs2;             // (sum w, sum w dll^2) per line of sight
};

