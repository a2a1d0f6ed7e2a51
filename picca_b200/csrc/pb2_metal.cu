// Metal distortion matrices: replace the pair loops of
// cf.compute_metal_dmat (reference py/picca/cf.py:890-1232, forest x forest, below) and
// xcf.compute_metal_dmat (py/picca/xcf.py:677-835, forest x object, further down).
//
// For every kept forest pair and every pixel pair whose pixels are consistent with the quasar
// redshift (z_abs < z_qso, cf.py:953-960, :985-992) the reference computes a DATA bin from the
// Lyman-alpha distances (cf.py:995-1013) and a MODEL bin from the distances the pixels would have
// if the absorption came from the metal transitions (cf.py:1021-1055), then scatters
// weights12 * z_weight_evol into dmat[data bin][model bin] and four effective-coordinate sums
// into the model bin -- NumPy bincounts per forest pair.  Here: one warp per kept forest pair
// (claimed from a device-wide counter), lane = column, rows in turn; the column window of a row
// comes from binary searches on the sorted distances (a superset of the pairs whose data bin is
// valid: outside it nothing is added anywhere); every bin is evaluated with the reference's
// IEEE operations in its order (floor for r_par bins, truncation for r_trans bins), so the bins
// are bit-exact; the sums are native red.global.add.f64 in any order (1e-9 tolerance), aggregated
// over runs of neighbouring columns with equal bins by a segmented warp scan.
// L2-reduction bound: six reductions per run of contributing pixel pairs.
#include "pb2_common.cuh"

struct MetalArgs {
    const double *z1, *rc1, *dm1, *pw1;  // per pixel of catalogue 1: absorber of forest 1
    const double *z2, *rc2, *dm2, *pw2;  // per pixel of catalogue 2: absorber of forest 2
    double evol_den;                     // (1 + z_ref)^(alpha_abs1 + alpha_abs2 - 2)
};

__device__ __forceinline__ int mt_lower(const double *__restrict__ a, int n, double v)
{
    int lo = 0, hi = n;  // first index with a[idx] >= v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// inclusive sum over the lanes [start, lane] of the caller's run
__device__ __forceinline__ double mt_seg_sum(double v, int start, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane - d >= start) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(256)
pb2_metal_dmat_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, MetalArgs A,
                      double *__restrict__ weights_dmat, double *__restrict__ dmat,
                      double *__restrict__ r_par_eff, double *__restrict__ r_trans_eff,
                      double *__restrict__ z_eff, double *__restrict__ weight_eff,
                      unsigned long long *__restrict__ work_ctr)
{
    const int lane = threadIdx.x & 31;
    const int np_i = P.num_bins_r_par, nt_i = P.num_bins_r_trans;
    const int npm = P.num_model_bins_r_par, ntm = P.num_model_bins_r_trans;
    const long long nbm = (long long)npm * ntm;
    const double span = sub_rn(P.r_par_max, P.r_par_min);
    const double close_cut = div_rn(span, (double)np_i);
    const bool windows = c1.sorted && c2.sorted;
    for (;;) {
        unsigned long long u = 0;
        if (lane == 0) u = atomicAdd(work_ctr, 1ull);
        const long long e = (long long)__shfl_sync(0xffffffffu, u, 0);
        if (e >= pr.n_pairs) break;
        if (pr.nb_keep && !pr.nb_keep[e]) continue;
        const int f1 = pr.f1_index[pr.nb_f1[e]];
        const int f2 = pr.nb_f2[e];
        const long long a1 = c1.offset[f1], a2 = c2.offset[f2];
        const int n1 = (int)(c1.offset[f1 + 1] - a1), n2 = (int)(c2.offset[f2 + 1] - a2);
        const double ch = pr.nb_cos[e], sh = pr.nb_sin[e];
        const double zq1 = c1.z_qso[f1], zq2 = c2.z_qso[f2];
        const bool same_hp = P.remove_same_half_plate_close_pairs &&
                             pb2_same_half_plate(c1, c2, f1, f2);
        const double *__restrict__ rc2 = c2.r_comov + a2;
        // conservative column window of a row: |r_comov1 - r_comov2| * cos < reach
        const double reach = (fmax(fabs(P.r_par_max), fabs(P.r_par_min)) / ch) * (1. + 1e-9) + 1e-9;
        for (int i = 0; i < n1; ++i) {
            const double z1 = A.z1[a1 + i];
            if (!(z1 < zq1)) continue;  // w = z1_abs1 < delta1.z_qso (cf.py:953)
            const double r1 = c1.r_comov[a1 + i], d1 = c1.dist_m[a1 + i], w1 = c1.weights[a1 + i];
            const double r1m = A.rc1[a1 + i], d1m = A.dm1[a1 + i], p1 = A.pw1[a1 + i];
            int j0 = 0, j1 = n2;
            if (windows) {
                j0 = mt_lower(rc2, n2, r1 - reach);
                j1 = mt_lower(rc2, n2, r1 + reach);
            }
            // 32 consecutive columns per step, all lanes in step (the shuffles below need the whole
            // warp).  Neighbouring columns mostly share their bins (r_par moves ~0.5 Mpc/h per
            // pixel), so the reductions are aggregated over runs of equal keys first: a segmented
            // warp scan, then ONE red.global.add per run and sum instead of one per pixel pair.
            for (int jb = j0; jb < j1; jb += 32) {
                const int j = jb + lane;
                bool dvalid = false, mvalid = false;
                int bin = 0, mbin = 0;
                double w12 = 0., wz = 0., s_rp = 0., s_rt = 0., s_z = 0.;
                if (j < j1) {
                    const double z2 = A.z2[a2 + j];
                    if (z2 < zq2) {  // cf.py:985
                        double r_par = mul_rn(sub_rn(r1, rc2[j]), ch);  // cf.py:995-999
                        if (!P.x_correlation) r_par = fabs(r_par);
                        const double r_trans = mul_rn(add_rn(d1, c2.dist_m[a2 + j]), sh);
                        w12 = mul_rn(w1, c2.weights[a2 + j]);
                        const double bp =
                            floor(mul_rn(div_rn(sub_rn(r_par, P.r_par_min), span), (double)np_i));
                        const double btf = mul_rn(div_rn(r_trans, P.r_trans_max), (double)nt_i);
                        if (bp >= 0. && bp < (double)np_i && btf < (double)nt_i) {  // cf.py:1016-1020
                            bin = (int)btf + nt_i * (int)bp;  // .astype(int) truncates (cf.py:1006)
                            if (same_hp && fabs(r_par) < close_cut) w12 = 0.;  // cf.py:1010-1013
                            dvalid = w12 != 0.;  // adding a zero weight changes nothing
                            double r_par_m = mul_rn(sub_rn(r1m, A.rc2[a2 + j]), ch);  // cf.py:1024-1032
                            if (!P.x_correlation) r_par_m = fabs(r_par_m);
                            const double r_trans_m = mul_rn(add_rn(d1m, A.dm2[a2 + j]), sh);
                            const double mbp = floor(
                                mul_rn(div_rn(sub_rn(r_par_m, P.r_par_min), span), (double)npm));
                            const double mbtf = mul_rn(div_rn(r_trans_m, P.r_trans_max), (double)ntm);
                            if (dvalid && mbp >= 0. && mbp < (double)npm && mbtf < (double)ntm) {
                                mvalid = true;  // cf.py:1051-1055
                                mbin = (int)mbtf + ntm * (int)mbp;
                                // z_weight_evol (cf.py:1033-1037): (a * b) / c
                                const double zwe = div_rn(mul_rn(p1, A.pw2[a2 + j]), A.evol_den);
                                wz = mul_rn(w12, zwe);                                   // cf.py:1056-1063
                                s_rp = mul_rn(mul_rn(r_par_m, w12), zwe);                // cf.py:1064-1068
                                s_rt = mul_rn(mul_rn(r_trans_m, w12), zwe);              // cf.py:1069-1073
                                s_z = mul_rn(mul_rn(div_rn(add_rn(z1, z2), 2.), w12), zwe);  // :1074-1083
                            }
                        }
                    }
                }
                if (!__any_sync(0xffffffffu, dvalid)) continue;
                // ---- data-bin runs: weights_dmat (cf.py:1021-1022)
                {
                    const int key = dvalid ? bin : -1;
                    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
                    const unsigned heads =
                        __ballot_sync(0xffffffffu, lane == 0 || prev != key || key < 0);
                    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
                    const double tot = mt_seg_sum(dvalid ? w12 : 0., start, lane);
                    if (dvalid && (lane == 31 || ((heads >> (lane + 1)) & 1u)))
                        atomicAdd(weights_dmat + bin, tot);
                }
                if (!__any_sync(0xffffffffu, mvalid)) continue;
                // ---- (data bin, model bin) runs: dmat and the four model-bin sums
                {
                    const long long key = mvalid ? (long long)bin * nbm + mbin : -1;
                    const long long prev = __shfl_up_sync(0xffffffffu, key, 1);
                    const unsigned heads =
                        __ballot_sync(0xffffffffu, lane == 0 || prev != key || key < 0);
                    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
                    const double t_wz = mt_seg_sum(wz, start, lane);
                    const double t_rp = mt_seg_sum(s_rp, start, lane);
                    const double t_rt = mt_seg_sum(s_rt, start, lane);
                    const double t_z = mt_seg_sum(s_z, start, lane);
                    if (mvalid && (lane == 31 || ((heads >> (lane + 1)) & 1u))) {
                        atomicAdd(dmat + key, t_wz);
                        atomicAdd(r_par_eff + mbin, t_rp);
                        atomicAdd(r_trans_eff + mbin, t_rt);
                        atomicAdd(z_eff + mbin, t_z);
                        atomicAdd(weight_eff + mbin, t_wz);  // cf.py:1084-1087
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
pb2_metal_dmat_cross_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr,
                            const double *__restrict__ mz1, const double *__restrict__ mrc1,
                            const double *__restrict__ mdm1, const double *__restrict__ mpw1,
                            double *__restrict__ weights_dmat, double *__restrict__ dmat,
                            double *__restrict__ r_par_eff, double *__restrict__ r_trans_eff,
                            double *__restrict__ z_eff, double *__restrict__ weight_eff)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    const int np_i = P.num_bins_r_par, nt_i = P.num_bins_r_trans;
    const int npm = P.num_model_bins_r_par, ntm = P.num_model_bins_r_trans;
    const long long nbm = (long long)npm * ntm;
    const double span = sub_rn(P.r_par_max, P.r_par_min);
    for (long long e = warp; e < pr.n_pairs; e += n_warps) {
        if (pr.nb_keep && !pr.nb_keep[e]) continue;
        const int f1 = pr.f1_index[pr.nb_f1[e]];
        const int q = pr.nb_f2[e];
        const long long a1 = c1.offset[f1];
        const int n1 = (int)(c1.offset[f1 + 1] - a1);
        const double ch = pr.nb_cos[e], sh = pr.nb_sin[e];
        const double zq1 = c1.z_qso[f1];
        const long long aq = c2.offset[q];
        const double r2 = c2.r_comov[aq], d2 = c2.dist_m[aq], w2 = c2.weights[aq], z2 = c2.z_qso[q];
        for (int i = lane; i < n1; i += 32) {
            const double z1 = mz1[a1 + i];
            if (!(z1 < zq1)) continue;  // xcf.py:735
            const double r_par = mul_rn(sub_rn(c1.r_comov[a1 + i], r2), ch);  // xcf.py:755-757
            const double r_trans = mul_rn(add_rn(c1.dist_m[a1 + i], d2), sh);
            const double w12 = mul_rn(c1.weights[a1 + i], w2);
            if (!(r_par > P.r_par_min && r_par < P.r_par_max && r_trans < P.r_trans_max)) continue;
            const int bp = (int)mul_rn(div_rn(sub_rn(r_par, P.r_par_min), span), (double)np_i);
            const int bt = (int)mul_rn(div_rn(r_trans, P.r_trans_max), (double)nt_i);
            if (bp >= np_i || bt >= nt_i) continue;  // rounding onto the upper edge: no such bin
            const int bin = bt + nt_i * bp;
            if (w12 != 0.) atomicAdd(weights_dmat + bin, w12);  // xcf.py:764-765
            const double r_par_m = mul_rn(sub_rn(mrc1[a1 + i], r2), ch);  // xcf.py:767-768
            const double r_trans_m = mul_rn(add_rn(mdm1[a1 + i], d2), sh);
            if (!(r_par_m > P.r_par_min && r_par_m < P.r_par_max && r_trans_m < P.r_trans_max))
                continue;  // xcf.py:785-789
            const int mbp = (int)mul_rn(div_rn(sub_rn(r_par_m, P.r_par_min), span), (double)npm);
            const int mbt = (int)mul_rn(div_rn(r_trans_m, P.r_trans_max), (double)ntm);
            if (mbp >= npm || mbt >= ntm || w12 == 0.) continue;
            const int mbin = mbt + ntm * mbp;
            const double zwe = mpw1[a1 + i];
            const double wz = mul_rn(w12, zwe);
            atomicAdd(dmat + (long long)bin * nbm + mbin, wz);                               // :791-798
            atomicAdd(r_par_eff + mbin, mul_rn(mul_rn(r_par_m, w12), zwe));                  // :800-804
            atomicAdd(r_trans_eff + mbin, mul_rn(mul_rn(r_trans_m, w12), zwe));              // :805-809
            atomicAdd(z_eff + mbin, mul_rn(mul_rn(div_rn(add_rn(z1, z2), 2.), w12), zwe));   // :810-814
            atomicAdd(weight_eff + mbin, wz);                                                // :815-818
        }
    }
}

extern "C" {

int32_t pb2_metal_dmat_cross(const pb2_catalog *cat1, const pb2_catalog *objs, const pb2_params *par,
                             const pb2_pairs *pairs, const double *d_z1, const double *d_rc1,
                             const double *d_dm1, const double *d_pw1, double *d_weights_dmat,
                             double *d_dmat, double *d_r_par_eff, double *d_r_trans_eff,
                             double *d_z_eff, double *d_weight_eff, void *stream)
{
    if (!cat1 || !objs || !par || !pairs || !d_z1 || !d_rc1 || !d_dm1 || !d_pw1 || !d_weights_dmat ||
        !d_dmat || !d_r_par_eff || !d_r_trans_eff || !d_z_eff || !d_weight_eff) {
        pb2_set_error("pb2_metal_dmat_cross: null pointer argument");
        return PB2_EINVAL;
    }
    if (par->rmu_binning || par->ang_correlation) {
        pb2_set_error("pb2_metal_dmat_cross: the reference has no rmu / angular metal matrix");
        return PB2_ECONFIG;
    }
    if (pairs->n_pairs <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    long long blocks = (pairs->n_pairs + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pb2_timing_begin(s);
    pb2_metal_dmat_cross_kernel<<<(unsigned)blocks, 256, 0, s>>>(
        *cat1, *objs, *par, *pairs, d_z1, d_rc1, d_dm1, d_pw1, d_weights_dmat, d_dmat, d_r_par_eff,
        d_r_trans_eff, d_z_eff, d_weight_eff);
    pb2_count_launch(1);
    int32_t rc = pb2_check_launch("pb2_metal_dmat_cross_kernel");
    pb2_timing_end(s);
    return rc;
}

int32_t pb2_metal_dmat_auto(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                            const pb2_pairs *pairs, const double *d_z1, const double *d_rc1,
                            const double *d_dm1, const double *d_pw1, const double *d_z2,
                            const double *d_rc2, const double *d_dm2, const double *d_pw2,
                            double evol_den, double *d_weights_dmat, double *d_dmat,
                            double *d_r_par_eff, double *d_r_trans_eff, double *d_z_eff,
                            double *d_weight_eff, void *stream)
{
    if (!cat1 || !cat2 || !par || !pairs || !d_z1 || !d_rc1 || !d_dm1 || !d_pw1 || !d_z2 || !d_rc2 ||
        !d_dm2 || !d_pw2 || !d_weights_dmat || !d_dmat || !d_r_par_eff || !d_r_trans_eff ||
        !d_z_eff || !d_weight_eff) {
        pb2_set_error("pb2_metal_dmat_auto: null pointer argument");
        return PB2_EINVAL;
    }
    if (par->rmu_binning || par->ang_correlation) {
        pb2_set_error("pb2_metal_dmat_auto: the reference has no rmu / angular metal matrix");
        return PB2_ECONFIG;
    }
    if (pairs->n_pairs <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    // the work counter lives for this call only, allocated in stream order: concurrent calls on
    // other streams / threads / devices each get their own
    unsigned long long *d_ctr = nullptr;
    int dev = 0, sms = 0;
    PB2_CUDA(cudaGetDevice(&dev));
    PB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PB2_CUDA(cudaMallocAsync((void **)&d_ctr, sizeof(unsigned long long), s));
    PB2_CUDA(cudaMemsetAsync(d_ctr, 0, sizeof(unsigned long long), s));
    MetalArgs A;
    A.z1 = d_z1, A.rc1 = d_rc1, A.dm1 = d_dm1, A.pw1 = d_pw1;
    A.z2 = d_z2, A.rc2 = d_rc2, A.dm2 = d_dm2, A.pw2 = d_pw2;
    A.evol_den = evol_den;
    pb2_timing_begin(s);
    pb2_metal_dmat_kernel<<<sms * 8, 256, 0, s>>>(*cat1, *cat2, *par, *pairs, A, d_weights_dmat,
                                                  d_dmat, d_r_par_eff, d_r_trans_eff, d_z_eff,
                                                  d_weight_eff, d_ctr);
    pb2_count_launch(1);
    int32_t rc = pb2_check_launch("pb2_metal_dmat_kernel");
    pb2_timing_end(s);
    cudaFreeAsync(d_ctr, s);
    return rc;
}

}  // extern "C"
