// Shared device/host helpers for libpicca_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "picca_b200.h"

#define PB2_SPEED_LIGHT 299792.458  // reference py/picca/constants.py:18 (km/s)
#define PB2_SMALL_ANGLE_CUT_OFF (2. / 3600. * 3.141592653589793 / 180.)  // constants.py:16
#define PB2_PI 3.141592653589793

void pb2_set_error(const char *fmt, ...);
int32_t pb2_check_launch(const char *what);
void pb2_count_launch(int n);
void pb2_timing_begin(cudaStream_t s);
void pb2_timing_end(cudaStream_t s);

#define PB2_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t err__ = (call);                                               \
        if (err__ != cudaSuccess) {                                               \
            pb2_set_error("%s failed: %s", #call, cudaGetErrorString(err__));     \
            return (int32_t)err__;                                                \
        }                                                                         \
    } while (0)

// ---- IEEE helpers that forbid FMA contraction: the reference (Numba/LLVM, no fastmath) rounds
// every product and sum separately, and bin assignment must be bit-exact.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ void atomic_add_f64(double *addr, double v) { atomicAdd(addr, v); }
__device__ __forceinline__ void atomic_add_i64(double *slot, long long v)
{
    atomicAdd(reinterpret_cast<unsigned long long *>(slot), (unsigned long long)v);
}

// Exact reference geometry + bin of one pixel pair (cf.py:350-380 / xcf.py:293-315).
// Returns the flat bin (>= 0) or -1 when the pair is rejected.  `cross_obj` selects the xcf
// variant (no abs, `r_par <= r_par_min` rejection, xcf.py:304-305).
struct PairGeom {
    double r_par, r_trans;
    int bin;
};

__device__ __forceinline__ PairGeom pb2_pair_exact(const pb2_params &P, double rc1, double dm1,
                                                   double rc2, double dm2, double ang,
                                                   double cos_half, double sin_half,
                                                   bool cross_obj, bool same_half_plate)
{
    PairGeom g;
    double r_par, r_trans;
    if (P.ang_correlation) {
        r_par = div_rn(rc1, rc2);
        if (!cross_obj && !P.x_correlation && r_par < 1.0) r_par = div_rn(1.0, r_par);
        r_trans = ang;
        if (cross_obj && P.rmu_binning) {  // xcf.py:300-302 applies rmu after either branch
            r_trans = sqrt(add_rn(mul_rn(r_trans, r_trans), mul_rn(r_par, r_par)));
            r_par = div_rn(r_par, r_trans);
        }
    } else {
        r_par = mul_rn(sub_rn(rc1, rc2), cos_half);
        r_trans = mul_rn(add_rn(dm1, dm2), sin_half);
        if (P.rmu_binning) {
            r_trans = sqrt(add_rn(mul_rn(r_trans, r_trans), mul_rn(r_par, r_par)));
            r_par = div_rn(r_par, r_trans);
        }
        if (!cross_obj && !P.x_correlation) r_par = fabs(r_par);
    }
    g.r_par = r_par;
    g.r_trans = r_trans;
    g.bin = -1;
    bool rej = (r_par >= P.r_par_max) || (r_trans >= P.r_trans_max) ||
               (cross_obj ? (r_par <= P.r_par_min) : (r_par < P.r_par_min));
    if (rej) return g;
    double span = sub_rn(P.r_par_max, P.r_par_min);
    double bp = floor(mul_rn(div_rn(sub_rn(r_par, P.r_par_min), span), (double)P.num_bins_r_par));
    double bt = floor(mul_rn(div_rn(r_trans, P.r_trans_max), (double)P.num_bins_r_trans));
    long long bins = (long long)add_rn(bt, mul_rn((double)P.num_bins_r_trans, bp));
    if (!cross_obj && P.remove_same_half_plate_close_pairs && same_half_plate) {
        if (fabs(r_par) < div_rn(span, (double)P.num_bins_r_par)) return g;  // cf.py:378-380
    }
    long long nb = (long long)P.num_bins_r_par * P.num_bins_r_trans;
    // the reference would write out of bounds here (measure-zero rounding case); drop instead
    if (bins < 0 || bins >= nb) return g;
    g.bin = (int)bins;
    return g;
}

// cf.py:321-328 / :341-348: is pixel redshift z_pix within zerr_cut_kms of the other quasar?
__device__ __forceinline__ bool pb2_zerr_close(const pb2_params &P, double z_pix, double z_qso)
{
    double z_qF = mul_rn(0.5, add_rn(z_pix, z_qso));
    double dv = div_rn(fabs(sub_rn(z_pix, z_qso)), add_rn(1.0, z_qF));
    dv = mul_rn(dv, PB2_SPEED_LIGHT);
    return dv < P.zerr_cut_kms;
}

__device__ __forceinline__ bool pb2_same_half_plate(const pb2_catalog &c1, const pb2_catalog &c2,
                                                    int f1, int f2)
{
    // cf.py:180-183
    long long fa = c1.fiberid[f1], fb = c2.fiberid[f2];
    return (c1.plate[f1] == c2.plate[f2]) && ((fa <= 500 && fb <= 500) || (fa > 500 && fb > 500));
}
