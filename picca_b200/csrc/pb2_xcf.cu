// Forest x object (quasar) pixel-pair histogram: replaces xcf.compute_xi's loop and
// xcf.compute_xi_forest_pairs_fast (reference py/picca/xcf.py:149-213, 223-322).
//
// One warp per forest; lane l owns one neighbouring object (its r_comov, dist_m, z, weight and
// cos/sin of half the separation stay in registers) and the warp sweeps the forest's pixels, every
// lane reading the same pixel (broadcast loads).  For a fixed object the pixels walk monotonically
// through the (r_par, r_trans) bins, so each lane accumulates a run of same-bin pixels in
// registers, in factored form (object constants applied at flush time), and flushes it with
// native fp64 global reductions when its bin changes.  Bins are the reference's, bit for bit:
// same sandwich test as the auto kernel, exact re-evaluation with true divisions when in doubt.
#include "pb2_common.cuh"

#define PB2_MAGIC 6755399441055744.0  // 2^52 + 2^51

struct XcfFast {
    double kp_lo, kp_hi, kt_lo, kt_hi, magic;
    int fast;
};

struct XRun {
    double sw, sdw, srp, srt, szw;  // sums over pixels of w1, delta1*w1, r_par*w1, r_trans*w1, z1*w1
    int cnt;
    int key;
};

__device__ __forceinline__ void xrun_flush(XRun &r, double wq, double zq, double *__restrict__ orow,
                                           int nb)
{
    if (r.key >= 0) {
        const double we = wq * r.sw;
        atomic_add_f64(orow + 0 * (size_t)nb + r.key, we);                       // xcf.py:318
        atomic_add_f64(orow + 1 * (size_t)nb + r.key, wq * r.sdw);               // xcf.py:308,317
        atomic_add_f64(orow + 2 * (size_t)nb + r.key, wq * r.srp);               // xcf.py:319
        atomic_add_f64(orow + 3 * (size_t)nb + r.key, wq * r.srt);               // xcf.py:320
        atomic_add_f64(orow + 4 * (size_t)nb + r.key, 0.5 * (wq * r.szw + zq * we));  // :286,:321
        atomic_add_i64(orow + 5 * (size_t)nb + r.key, (long long)r.cnt);
    }
    r.sw = r.sdw = r.srp = r.srt = r.szw = 0.;
    r.cnt = 0;
    r.key = -1;
}

// first index in the non-decreasing a[0..n) with a[i] > v (strict) / a[i] >= v
__device__ __forceinline__ int lane_upper_bound(const double *__restrict__ a, int n, double v,
                                                bool strict)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double x = __ldg(a + mid);
        const bool left = strict ? (x <= v) : (x < v);
        if (left) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

template <bool FAST>
__global__ void __launch_bounds__(256)
pb2_xi_cross_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, XcfFast F,
                    const int32_t *__restrict__ out_row, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const unsigned np_u = (unsigned)P.num_bins_r_par, nt_u = (unsigned)P.num_bins_r_trans;
    const bool zcut = P.has_z_min_pairs || P.has_z_max_pairs;
    const double magic = F.magic;

    for (long long k = warp; k < pr.n_f1; k += nwarps) {
        const int f1 = pr.f1_index[k];
        const long long a = c1.offset[f1];
        const int n1 = (int)(c1.offset[f1 + 1] - a);
        const long long e0 = pr.nb_offset[k], e1 = pr.nb_offset[k + 1];
        if (n1 == 0 || e1 == e0) continue;  // xcf.py:157
        double *__restrict__ orow = out + (size_t)out_row[k] * 6 * nb;
        const double *__restrict__ p_rc1 = c1.r_comov + a;
        const double *__restrict__ p_dm1 = c1.dist_m + a;
        const double *__restrict__ p_z1 = c1.z + a;
        const double *__restrict__ p_w1 = c1.weights + a;
        const double *__restrict__ p_dw1 = c1.delta_w + a;

        for (long long eb = e0; eb < e1; eb += 32) {
            const long long e = eb + lane;
            const bool have = e < e1;
            double rcq = 0., dmq = 0., zq = 0., wq = 0., ang = 0., ch = 1., sh = 0.;
            if (have) {
                const int f2 = pr.nb_f2[e];
                const long long q = c2.offset[f2];
                rcq = c2.r_comov[q];
                dmq = c2.dist_m[q];
                zq = c2.z[q];
                wq = c2.weights[q];
                ang = pr.nb_ang[e];
                ch = pr.nb_cos[e];
                sh = pr.nb_sin[e];
            }
            const bool vq = have && (wq != 0.);  // xcf.py:283

            // ---- per-lane pixel window (superset), warp sweeps the union
            int ilo = 0, ihi = n1;
            if (FAST) {
                // r_par_min < (rc1 - rcq) ch < r_par_max  and  (dm1 + dmq) sh < r_trans_max
                const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
                const double hi_rc = rcq + P.r_par_max * inv_c;
                const double lo_rc = rcq + P.r_par_min * inv_c;
                ilo = lane_upper_bound(p_rc1, n1, lo_rc - fabs(lo_rc) * 1e-9 - 1e-9, false);
                ihi = lane_upper_bound(p_rc1, n1, hi_rc + fabs(hi_rc) * 1e-9 + 1e-9, true);
                const double tsum = P.r_trans_max * inv_s;
                if (isfinite(tsum))
                    ihi = min(ihi, lane_upper_bound(p_dm1, n1, (tsum - dmq) * (1. + 1e-9) + 1e-9, true));
            }
            if (!vq) {
                ilo = n1;
                ihi = 0;
            }
            int ILO = ilo, IHI = ihi;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                ILO = min(ILO, __shfl_xor_sync(0xffffffffu, ILO, m));
                IHI = max(IHI, __shfl_xor_sync(0xffffffffu, IHI, m));
            }

            XRun run;
            run.key = -1;
            xrun_flush(run, 0., 0., orow, nb);
            for (int i = ILO; i < IHI; i++) {
                const double rc1 = __ldg(p_rc1 + i), dm1 = __ldg(p_dm1 + i);
                const double w1 = __ldg(p_w1 + i), dw1 = __ldg(p_dw1 + i), z1 = __ldg(p_z1 + i);
                const bool both = vq && (w1 != 0.) && (i >= ilo) && (i < ihi);  // xcf.py:279
                bool in = false;
                int bin = -1;
                double rp = 0., rt = 0.;
                if (FAST) {
                    rp = mul_rn(sub_rn(rc1, rcq), ch);
                    rt = mul_rn(add_rn(dm1, dmq), sh);
                    const double x = sub_rn(rp, P.r_par_min);
                    const int bpl = __double2loint(__fma_rd(x, F.kp_lo, magic));
                    const int bph = __double2loint(__fma_rd(x, F.kp_hi, magic));
                    const int btl = __double2loint(__fma_rd(rt, F.kt_lo, magic));
                    const int bth = __double2loint(__fma_rd(rt, F.kt_hi, magic));
                    // x == 0 exactly is rejected by the reference (r_par <= r_par_min, xcf.py:305)
                    const bool sure = (bpl == bph) && (btl == bth) && (x != 0.);
                    in = both && sure && ((unsigned)bpl < np_u) && ((unsigned)btl < nt_u);
                    bin = btl + (int)nt_u * bpl;
                    if (both && !sure) {
                        PairGeom g = pb2_pair_exact(P, rc1, dm1, rcq, dmq, ang, ch, sh, true, false);
                        in = g.bin >= 0;
                        bin = g.bin;
                    }
                } else if (both) {
                    PairGeom g = pb2_pair_exact(P, rc1, dm1, rcq, dmq, ang, ch, sh, true, false);
                    in = g.bin >= 0;
                    bin = g.bin;
                    rp = g.r_par;
                    rt = g.r_trans;
                }
                if (zcut && in) {
                    const double zm = div_rn(add_rn(z1, zq), 2.);  // xcf.py:286-291
                    if (P.has_z_min_pairs && zm < P.z_min_pairs) in = false;
                    if (P.has_z_max_pairs && zm > P.z_max_pairs) in = false;
                }
                if (in) {
                    if (bin != run.key) {
                        xrun_flush(run, wq, zq, orow, nb);
                        run.key = bin;
                    }
                    run.sw += w1;
                    run.sdw += dw1;
                    run.szw = fma(z1, w1, run.szw);
                    run.srp = fma(rp, w1, run.srp);
                    run.srt = fma(rt, w1, run.srt);
                    run.cnt += 1;
                }
            }
            xrun_flush(run, wq, zq, orow, nb);
        }
    }
}

extern "C" {

int32_t pb2_xi_cross(const pb2_catalog *cat1, const pb2_catalog *objs, const pb2_params *par,
                     const pb2_pairs *pairs, const int32_t *d_out_row, int64_t n_rows,
                     double *d_out, int32_t variant, void *stream)
{
    if (!cat1 || !objs || !par || !pairs || !d_out_row || !d_out) {
        pb2_set_error("pb2_xi_cross: null pointer argument");
        return PB2_EINVAL;
    }
    (void)n_rows;
    if (pairs->n_pairs <= 0 || pairs->n_f1 <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    XcfFast F;
    const double kp = (double)par->num_bins_r_par / (par->r_par_max - par->r_par_min);
    const double kt = (double)par->num_bins_r_trans / par->r_trans_max;
    const double eps = 9.094947017729282e-13;  // 2^-40
    F.kp_lo = kp * (1. - eps);
    F.kp_hi = kp * (1. + eps);
    F.kt_lo = kt * (1. - eps);
    F.kt_hi = kt * (1. + eps);
    F.magic = PB2_MAGIC;
    // variant 1 = validation: every pair through the reference expression, no windows
    F.fast = (variant == 0 && !par->rmu_binning && !par->ang_correlation && cat1->sorted &&
              par->num_bins_r_par <= 4096 && par->num_bins_r_trans <= 4096 &&
              par->r_par_max > par->r_par_min && par->r_trans_max > 0.) ? 1 : 0;
    long long blocks = (pairs->n_f1 + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    pb2_timing_begin(s);
    if (F.fast)
        pb2_xi_cross_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(*cat1, *objs, *par, *pairs, F,
                                                                    d_out_row, d_out);
    else
        pb2_xi_cross_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(*cat1, *objs, *par, *pairs, F,
                                                                     d_out_row, d_out);
    pb2_count_launch(1);
    int32_t rc = pb2_check_launch("pb2_xi_cross_kernel");
    pb2_timing_end(s);
    return rc;
}

}  // extern "C"
