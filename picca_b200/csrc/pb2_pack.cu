// Packed record copies of the diagonal-lane xi kernel, written in HBM from the SoA arrays
// (layout: include/picca_b200.h, pb2_catalog; specification in NumPy: catalog.diag_records_host).
// Pure data movement -- one read of five SoA arrays, two 48-byte record writes per kept pixel --
// that took 74 % of the host-side packing time as NumPy fancy indexing; on the device it is
// HBM-bound and takes milliseconds, and the host->device copy of a catalogue shrinks from
// 152 to 56 bytes per pixel.
#include "pb2_common.cuh"

#define PK_DUMMY_COL 1e300  // distance of the dummy pixels around a line of sight (interleaved copy)
#define PK_DUMMY_ROW 1e299  // ... and after it in the natural-order copy

// every record := (dummy, dummy, 0, 0, 0, 0)
__global__ void pb2_pack_fill_kernel(double *__restrict__ rec, long long n_rec, double dummy)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    double2 *p = reinterpret_cast<double2 *>(rec + 6 * r);
    p[0] = make_double2(dummy, dummy);
    p[1] = make_double2(0., 0.);
    p[2] = make_double2(0., 0.);
}

// one warp per line of sight: compact the pixels with weight != 0 (the reference never counts
// the others, cf.py:318,331) by ballot prefix and write each as a record into both copies
__global__ void pb2_pack_diag_kernel(pb2_catalog c, double *__restrict__ dg_rec,
                                     double *__restrict__ il_rec)
{
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= c.n_los) return;
    const int lane = threadIdx.x & 31;
    const long long a = c.offset[f];
    const int n = (int)(c.offset[f + 1] - a);
    const long long dg0 = c.dg_offset[f], il0 = c.il_offset[f];
    int rank0 = 0;
    for (int p0 = 0; p0 < n; p0 += 32) {
        const int p = p0 + lane;
        double w = 0.;
        if (p < n) w = c.weights[a + p];
        const bool keep = p < n && w != 0.;
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int rank = rank0 + __popc(mask & ((1u << lane) - 1u));
            const double2 r0 = make_double2(c.r_comov[a + p], c.dist_m[a + p]);
            const double2 r1 = make_double2(w, c.delta_w[a + p]);
            const double2 r2 = make_double2(mul_rn(0.5, c.z[a + p]), 0.);
            double2 *d = reinterpret_cast<double2 *>(dg_rec + 6 * (dg0 + rank));
            d[0] = r0, d[1] = r1, d[2] = r2;
            const int jp = rank + PB2_DIAG_PAD;
            double2 *q = reinterpret_cast<double2 *>(
                il_rec + 6 * ((long long)(jp % PB2_DIAG_LANES) * c.il_total + il0 + jp / PB2_DIAG_LANES));
            q[0] = r0, q[1] = r1, q[2] = r2;
        }
        rank0 += __popc(mask);
    }
}

// delta_w = delta * weights (0 where the weight is 0: a NaN delta there must not leak, and the
// reference never visits such pixels, cf.py:318,331) and z_w = z * weights, rounded like NumPy's
__global__ void pb2_derive_products_kernel(long long n, const double *__restrict__ w,
                                           const double *__restrict__ delta,
                                           const double *__restrict__ z, double *__restrict__ delta_w,
                                           double *__restrict__ z_w)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double wi = w[i];
    delta_w[i] = wi != 0. ? mul_rn(delta[i], wi) : 0.;
    z_w[i] = mul_rn(z[i], wi);
}

// One warp per forest: count of non-zero-weight pixels; flags[0] != 0 when a kept pixel carries a
// non-finite r_comov / dist_m / weight / delta*weight / z, flags[1] != 0 when r_comov or dist_m
// decreases inside a forest or is not finite anywhere; flags[2] = bit pattern of
// max(|r_comov|, |dist_m|) over the kept pixels (non-negative doubles order like integers).
__global__ void pb2_catalog_stats_kernel(long long n_los, const long long *__restrict__ offset,
                                         const double *__restrict__ w, const double *__restrict__ rc,
                                         const double *__restrict__ dm, const double *__restrict__ z,
                                         const double *__restrict__ dw, int *__restrict__ count,
                                         unsigned long long *__restrict__ flags)
{
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= n_los) return;
    const int lane = threadIdx.x & 31;
    const long long a = offset[f];
    const int n = (int)(offset[f + 1] - a);
    int cnt = 0;
    bool bad = false, unsorted = false;
    double reach = 0.;
    for (int p = lane; p < n; p += 32) {
        const double wi = w[a + p], r = rc[a + p], d = dm[a + p];
        if (!isfinite(r) || !isfinite(d)) unsorted = true;
        if (p + 1 < n && (rc[a + p + 1] < r || dm[a + p + 1] < d)) unsorted = true;
        if (wi != 0.) {
            cnt++;
            if (!isfinite(r) || !isfinite(d) || !isfinite(wi) || !isfinite(dw[a + p]) ||
                !isfinite(z[a + p])) bad = true;
            reach = fmax(reach, fmax(fabs(r), fabs(d)));
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
        reach = fmax(reach, __shfl_xor_sync(0xffffffffu, reach, m));
    }
    bad = __any_sync(0xffffffffu, bad);
    unsorted = __any_sync(0xffffffffu, unsorted);
    if (lane == 0) {
        count[f] = cnt;
        if (bad) atomicOr(flags, 1ull);
        if (unsorted) atomicOr(flags + 1, 1ull);
        if (isfinite(reach)) atomicMax(flags + 2, (unsigned long long)__double_as_longlong(reach));
    }
}

extern "C" {

int32_t pb2_catalog_stats(int64_t n_los, const int64_t *d_offset, const double *d_weights,
                          const double *d_r_comov, const double *d_dist_m, const double *d_z,
                          const double *d_delta_w, int32_t *d_count, int64_t *d_flags, void *stream)
{
    if (n_los < 0 || (n_los > 0 && (!d_offset || !d_weights || !d_r_comov || !d_dist_m || !d_z ||
                                    !d_delta_w || !d_count || !d_flags))) {
        pb2_set_error("pb2_catalog_stats: bad argument");
        return PB2_EINVAL;
    }
    if (n_los == 0) return 0;
    pb2_catalog_stats_kernel<<<(unsigned)((n_los + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        n_los, (const long long *)d_offset, d_weights, d_r_comov, d_dist_m, d_z, d_delta_w, d_count,
        (unsigned long long *)d_flags);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_catalog_stats_kernel");
}

int32_t pb2_derive_products(int64_t n_pix, const double *d_weights, const double *d_delta,
                            const double *d_z, double *d_delta_w, double *d_z_w, void *stream)
{
    if (n_pix < 0 || (n_pix > 0 && (!d_weights || !d_delta || !d_z || !d_delta_w || !d_z_w))) {
        pb2_set_error("pb2_derive_products: bad argument");
        return PB2_EINVAL;
    }
    if (n_pix == 0) return 0;
    pb2_derive_products_kernel<<<(unsigned)((n_pix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n_pix, d_weights, d_delta, d_z, d_delta_w, d_z_w);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_derive_products_kernel");
}

/* Fill cat->dg_rec (dg_total records) and cat->il_rec (PB2_DIAG_LANES * il_total records) from the
 * SoA arrays of the catalogue, whose dg_offset / dg_count / il_offset / il_total describe the
 * layout (computed by the host from the weights, catalog._diag_metadata). */
int32_t pb2_pack_diag(const pb2_catalog *cat, int64_t dg_total, void *stream)
{
    if (!cat || !cat->dg_rec || !cat->il_rec || !cat->dg_offset || !cat->il_offset || dg_total <= 0) {
        pb2_set_error("pb2_pack_diag: bad argument");
        return PB2_EINVAL;
    }
    if (cat->dg_lanes != PB2_DIAG_LANES) {
        pb2_set_error("pb2_pack_diag: catalogue laid out for %d lanes, library built for %d",
                      cat->dg_lanes, PB2_DIAG_LANES);
        return PB2_ECONFIG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    double *dg = const_cast<double *>(cat->dg_rec), *il = const_cast<double *>(cat->il_rec);
    const long long il_rec_n = (long long)PB2_DIAG_LANES * cat->il_total;
    pb2_pack_fill_kernel<<<(unsigned)((dg_total + 255) / 256), 256, 0, s>>>(dg, dg_total, PK_DUMMY_ROW);
    pb2_pack_fill_kernel<<<(unsigned)((il_rec_n + 255) / 256), 256, 0, s>>>(il, il_rec_n, PK_DUMMY_COL);
    pb2_count_launch(2);
    int32_t rc = pb2_check_launch("pb2_pack_fill_kernel");
    if (rc || cat->n_los <= 0) return rc;
    pb2_pack_diag_kernel<<<(unsigned)((cat->n_los + 7) / 8), 256, 0, s>>>(*cat, dg, il);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_pack_diag_kernel");
}

}  // extern "C"
