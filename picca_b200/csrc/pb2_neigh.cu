// Device-side neighbour search: replaces cf.fill_neighs (reference py/picca/cf.py:82-135) and
// xcf.fill_neighs (py/picca/xcf.py:71-123) including QSO.get_angle_between (data.py:106-162).
//
// One warp per line of sight.  The reference asks healpy.query_disc for candidate pixels and then
// applies the exact `ang < ang_max` filter; any superset of candidate pixels is equivalent, so
// here the candidate pixels are those whose bounding cap (centre + radius of the actual members,
// built at pack time) can intersect the disc.  Members are visited in catalogue order (ascending
// HEALPix id, then list order) and compacted with warp ballots, which preserves the reference's
// neighbour order -- the --rej RNG contract of compute_dmat depends on it.
#include "pb2_common.cuh"

struct NeighHit {
    bool keep;
    double ang;
};

// get_angle_between(self=f1, other=f2): data.py:126-141 / :150-161
__device__ __forceinline__ double pb2_angle(const pb2_catalog &c1, int f1, const pb2_catalog &c2,
                                            int f2)
{
    double cosv = add_rn(add_rn(mul_rn(c2.x_cart[f2], c1.x_cart[f1]), mul_rn(c2.y_cart[f2], c1.y_cart[f1])),
                       mul_rn(c2.z_cart[f2], c1.z_cart[f1]));
    if (cosv >= 1.) cosv = 1.;
    else if (cosv <= -1.) cosv = -1.;
    double ang = acos(cosv);
    double dra = sub_rn(c2.ra[f2], c1.ra[f1]);
    double ddec = sub_rn(c2.dec[f2], c1.dec[f1]);
    if (fabs(dra) < PB2_SMALL_ANGLE_CUT_OFF && fabs(ddec) < PB2_SMALL_ANGLE_CUT_OFF) {
        double t = mul_rn(c1.cos_dec[f1], dra);
        ang = sqrt(add_rn(mul_rn(ddec, ddec), mul_rn(t, t)));
    }
    return ang;
}

__device__ __forceinline__ NeighHit pb2_neigh_test(const pb2_catalog &c1, int f1,
                                                   const pb2_catalog &c2, int f2,
                                                   const pb2_params &P, int mode)
{
    NeighHit h;
    h.ang = pb2_angle(c1, f1, c2, f2);
    bool keep = (c1.thingid[f1] != c2.thingid[f2]) && (h.ang < P.ang_max);
    if (mode == 0) keep = keep && (c1.ra[f1] > c2.ra[f2]);  // cf.py:129-135
    if (mode == 2 && keep) {
        if (P.has_zerr_cut) {  // xcf.py:102-115
            double ang_deg = mul_rn(div_rn(180.0, PB2_PI), h.ang);
            double zq1 = c1.z_qso[f1], zq2 = c2.z_qso[f2];
            double z_qq = mul_rn(0.5, add_rn(zq1, zq2));
            double dv = div_rn(fabs(sub_rn(zq1, zq2)), add_rn(1.0, z_qq));
            dv = mul_rn(dv, PB2_SPEED_LIGHT);
            if (ang_deg < P.zerr_cut_deg && dv < P.zerr_cut_kms) keep = false;
        }
        if (!P.ang_correlation) {  // xcf.py:117-121
            double f = P.rmu_binning ? P.r_trans_max : 1.0;
            double ch = cos(div_rn(h.ang, 2.));
            int64_t a = c1.offset[f1], b = c1.offset[f1 + 1];
            double rq = c2.r_comov[c2.offset[f2]];
            if (b > a) {
                keep = keep && (mul_rn(sub_rn(c1.r_comov[a], rq), ch) < mul_rn(P.r_par_max, f));
                keep = keep && (mul_rn(sub_rn(c1.r_comov[b - 1], rq), ch) > mul_rn(P.r_par_min, f));
            } else {
                keep = false;  // the reference would raise IndexError on an empty forest
            }
        }
    }
    h.keep = keep;
    return h;
}

template <bool FILL>
__global__ void __launch_bounds__(256)
pb2_neigh_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, int mode, int64_t n_f1,
                 const int32_t *__restrict__ f1_index, int32_t *__restrict__ count,
                 const int64_t *__restrict__ nb_offset, int32_t *__restrict__ nb_f1,
                 int32_t *__restrict__ nb_f2, double *__restrict__ nb_ang,
                 double *__restrict__ nb_cos, double *__restrict__ nb_sin)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const unsigned lt_mask = (1u << lane) - 1u;

    for (int64_t k = warp; k < n_f1; k += nwarps) {
        const int f1 = f1_index[k];
        const double x1 = c1.x_cart[f1], y1 = c1.y_cart[f1], z1 = c1.z_cart[f1];
        int64_t pos = FILL ? nb_offset[k] : 0;
        int total = 0;
        for (int h0 = 0; h0 < c2.n_hp; h0 += 32) {
            const int h = h0 + lane;
            bool cand = false;
            if (h < c2.n_hp) {
                double d = x1 * c2.cap_x[h] + y1 * c2.cap_y[h] + z1 * c2.cap_z[h];
                d = fmin(1.0, fmax(-1.0, d));
                // conservative: 1e-6 rad covers acos conditioning near 0 and cap-radius rounding
                cand = acos(d) <= P.ang_max + c2.cap_rad[h] + 1e-6;
            }
            unsigned hits = __ballot_sync(0xffffffffu, cand);
            while (hits) {
                const int hb = __ffs(hits) - 1;
                hits &= hits - 1;
                const int first = c2.hp_first[h0 + hb], last = c2.hp_first[h0 + hb + 1];
                for (int m0 = first; m0 < last; m0 += 32) {
                    const int f2 = m0 + lane;
                    NeighHit nh;
                    nh.keep = false;
                    nh.ang = 0.;
                    if (f2 < last) nh = pb2_neigh_test(c1, f1, c2, f2, P, mode);
                    const unsigned kept = __ballot_sync(0xffffffffu, nh.keep);
                    if (FILL && nh.keep) {
                        const int64_t e = pos + __popc(kept & lt_mask);
                        nb_f1[e] = (int32_t)k;
                        nb_f2[e] = f2;
                        nb_ang[e] = nh.ang;
                        const double half = div_rn(nh.ang, 2.);
                        nb_cos[e] = cos(half);
                        nb_sin[e] = sin(half);
                    }
                    pos += __popc(kept);
                    total += __popc(kept);
                }
            }
        }
        if (!FILL && lane == 0) count[k] = total;
    }
}

static int32_t check_catalogs(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par)
{
    if (!c1 || !c2 || !par) {
        pb2_set_error("null catalogue/params pointer");
        return PB2_EINVAL;
    }
    return 0;
}

extern "C" {

int32_t pb2_neigh_count(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                        int32_t mode, int64_t n_f1, const int32_t *d_f1_index, int32_t *d_count,
                        void *stream)
{
    if (int32_t e = check_catalogs(cat1, cat2, par)) return e;
    if (n_f1 <= 0) return 0;
    const int threads = 256, wpb = threads / 32;
    int64_t blocks = (n_f1 + wpb - 1) / wpb;
    if (blocks > 148 * 64) blocks = 148 * 64;
    pb2_neigh_kernel<false><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        *cat1, *cat2, *par, mode, n_f1, d_f1_index, d_count, nullptr, nullptr, nullptr, nullptr,
        nullptr, nullptr);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_neigh_count");
}

int32_t pb2_neigh_fill(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                       int32_t mode, int64_t n_f1, const int32_t *d_f1_index,
                       const int64_t *d_nb_offset, int32_t *d_nb_f1, int32_t *d_nb_f2,
                       double *d_nb_ang, double *d_nb_cos, double *d_nb_sin, void *stream)
{
    if (int32_t e = check_catalogs(cat1, cat2, par)) return e;
    if (n_f1 <= 0) return 0;
    const int threads = 256, wpb = threads / 32;
    int64_t blocks = (n_f1 + wpb - 1) / wpb;
    if (blocks > 148 * 64) blocks = 148 * 64;
    pb2_neigh_kernel<true><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        *cat1, *cat2, *par, mode, n_f1, d_f1_index, nullptr, d_nb_offset, d_nb_f1, d_nb_f2,
        d_nb_ang, d_nb_cos, d_nb_sin);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_neigh_fill");
}

}  // extern "C"
