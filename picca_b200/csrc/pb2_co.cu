// Object x object pair counting (first half): replaces the pair loop of
// co.compute_xi and co.compute_xi_forest_pairs (reference py/picca/co.py:77-132, :135-202).
// The neighbour list is the CSR pair list of pb2_neigh (mode 1: thingid != and ang < ang_max,
// co.py:43-69); the mean-redshift cut of co.py:70-74 is applied here, per pair.  One thread per
// (object 1, object 2) pair; the bin is the reference's expression (truncation, co.py:185-189)
// evaluated with its IEEE operations in order -> num_pairs bit-exact; the sums are native
// red.global.add in any order (1e-9).  Trivially latency-bound: 5e5 objects x ~100 neighbours.
#include "pb2_common.cuh"

__global__ void pb2_co_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr,
                              int take_abs, int has_z_cut, double z_cut_min, double z_cut_max,
                              const int32_t *__restrict__ out_row, double *__restrict__ out)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= pr.n_pairs) return;
    const int k1 = pr.nb_f1[e];
    const int q1 = pr.f1_index[k1], q2 = pr.nb_f2[e];
    const double z1 = c1.z_qso[q1], z2 = c2.z_qso[q2];
    if (has_z_cut) {  // co.py:70-74
        const double zm = div_rn(add_rn(z2, z1), 2.);
        if (!(zm >= z_cut_min && zm < z_cut_max)) return;
    }
    const long long a1 = c1.offset[q1], a2 = c2.offset[q2];
    double r_par = mul_rn(sub_rn(c1.r_comov[a1], c2.r_comov[a2]), pr.nb_cos[e]);  // co.py:171
    if (take_abs) r_par = fabs(r_par);                                             // co.py:172-173
    const double r_trans = mul_rn(add_rn(c1.dist_m[a1], c2.dist_m[a2]), pr.nb_sin[e]);
    const double z = div_rn(add_rn(z1, z2), 2.);
    const double w12 = mul_rn(c1.weights[a1], c2.weights[a2]);
    if (!(r_par >= P.r_par_min && r_par < P.r_par_max && r_trans < P.r_trans_max && w12 > 0.))
        return;  // co.py:178-179
    const int bp = (int)mul_rn(div_rn(sub_rn(r_par, P.r_par_min), sub_rn(P.r_par_max, P.r_par_min)),
                               (double)P.num_bins_r_par);
    const int bt = (int)mul_rn(div_rn(r_trans, P.r_trans_max), (double)P.num_bins_r_trans);
    if (bp >= P.num_bins_r_par || bt >= P.num_bins_r_trans) return;  // rounding onto the upper edge
    const int bin = bt + P.num_bins_r_trans * bp;
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    double *o = out + (long long)out_row[k1] * 5 * nb + bin;
    atomic_add_f64(o, w12);                               // co.py:192
    atomic_add_f64(o + nb, mul_rn(r_par, w12));           // co.py:193
    atomic_add_f64(o + 2 * (long long)nb, mul_rn(r_trans, w12));
    atomic_add_f64(o + 3 * (long long)nb, mul_rn(z, w12));
    atomic_add_i64(o + 4 * (long long)nb, 1);             // co.py:196
}

extern "C" {

/* d_out is [n_rows][5][np*nt]: sums of weight, r_par*w, r_trans*w, z*w (fp64) and num_pairs (int64
 * in the same 8-byte slots), accumulated into; d_out_row[k] = output row of f1_index[k]. */
int32_t pb2_co_pairs(const pb2_catalog *objs1, const pb2_catalog *objs2, const pb2_params *par,
                     const pb2_pairs *pairs, int32_t take_abs, int32_t has_z_cut, double z_cut_min,
                     double z_cut_max, const int32_t *d_out_row, int64_t n_rows, double *d_out,
                     void *stream)
{
    if (!objs1 || !objs2 || !par || !pairs || !d_out_row || !d_out || n_rows <= 0) {
        pb2_set_error("pb2_co_pairs: bad argument");
        return PB2_EINVAL;
    }
    if (pairs->n_pairs <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    pb2_timing_begin(s);
    pb2_co_kernel<<<(unsigned)((pairs->n_pairs + 255) / 256), 256, 0, s>>>(
        *objs1, *objs2, *par, *pairs, take_abs, has_z_cut, z_cut_min, z_cut_max, d_out_row, d_out);
    pb2_count_launch(1);
    int32_t rc = pb2_check_launch("pb2_co_kernel");
    pb2_timing_end(s);
    return rc;
}

}  // extern "C"
