// Wick expansion of the covariance, diagrams T1-T3 of the auto-correlation
// (cf.compute_wick_terms / compute_wickT123_pairs, reference py/picca/cf.py:1326-1626) and T1-T4
// of the forest x object cross-correlation (xcf.compute_wick_terms / compute_wickT1234_pairs,
// py/picca/xcf.py:838-1351).  SURVEY.md 8f rank 4.
//
// For one forest pair the reference lists the in-range pixel pairs L = {(i, j)} (j outer, i inner,
// cf.py:1562-1596) and then visits every pair {a < b} of list entries (:1598-1624): O(|L|^2)
// products of the weighted 1-D correlations  A[i,i'] = w_i w_i' xi1d(|ll_i - ll_i'|) sqrt(v_i v_i')
// of forest 1 and B[j,j'] of forest 2, scattered into t2 / t3 [bin(a), bin(b)] and its transpose.
//
// Here: one CTA per kept forest pair (claimed from a device-wide counter).  Phase 0 tabulates A, B
// and the bin matrix P[j][i] (-1 = out of range; the reference's IEEE expression with int()
// truncation, cf.py:1585-1589) in the CTA's slab of scratch (L2-resident) and adds the O(|L|)
// terms (weights_wick, num_pairs_wick, t1).  Phase 1: every warp takes list entries a in turn and
// sweeps the entries b > a row by row (j' >= j), lane = 32 consecutive i': P and the row A[i][.]
// are coalesced loads, B[j][j'] is one broadcast load per row.  Along i' the bin changes every few
// pixels, so products are summed over runs of equal bin with a segmented warp scan and each run
// costs two native red.global.add.f64 (the [p_a, p_b] entry and its transpose) -- the same
// aggregation as the metal matrices (pb2_metal.cu).  Sums are re-associated w.r.t. the reference
// (1e-9 tolerance); which (bin, bin) cells are touched, weights_wick's terms and num_pairs_wick
// are exact.  L2-reduction / issue bound; no tensor cores (scatter, data-dependent bins).
#include "pb2_common.cuh"

#define WK_THREADS 512
#define WK_WARPS (WK_THREADS / 32)

struct WickTable {      // scipy interp1d(kind="nearest", fill_value="extrapolate") as a table:
    int n;              // y[searchsorted(xb, x, side="left").clip(0, n-1)], xb = the n-1 midpoints
    const double *xb;   // scipy forms (x[1:]/2 + x[:-1]/2), computed on the host
    const double *y;
};

struct WickWork {
    unsigned long long *counter;
    char *slab_base;
    long long slab_stride;
    long long p_cap;     // int32 entries of P per slab
    long long a_cap;     // doubles of A per slab
    long long b_cap;     // doubles of B per slab (auto) / 0
    int rows_cap;        // rows of P per slab (max pixels of forest 2 / max neighbours)
};

__device__ __forceinline__ double wk_nearest(const WickTable &T, double x)
{
    int lo = 0, hi = T.n - 1;  // first bound >= x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (T.xb[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    return T.y[lo];
}

// inclusive sum over the lanes [start, lane] of the caller's run
__device__ __forceinline__ double wk_seg_sum(double v, int start, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane - d >= start) v += t;
    }
    return v;
}

// all lanes of the warp: add v into T[pa][key] and T[key][pa] (cf.py:1612-1613: both orders),
// aggregated over runs of equal key (< 0: nothing to add)
__device__ __forceinline__ void wk_flush(double *__restrict__ T, long long nb, int pa, int key,
                                         double v, int lane)
{
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != key || key < 0);
    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    const double tot = wk_seg_sum(key >= 0 ? v : 0., start, lane);
    if (key >= 0 && (lane == 31 || ((heads >> (lane + 1)) & 1u)) && tot != 0.) {
        atomicAdd(T + (long long)pa * nb + key, tot);
        atomicAdd(T + (long long)key * nb + pa, tot);
    }
}

// the reference's selection and bin of one pixel pair (cf.py:1564-1571, :1583-1590; xcf.py the
// same without abs): int() truncation; -1 when out of range
__device__ __forceinline__ int wk_bin(const pb2_params &P, double rc1, double rc2, double ch,
                                      double sh, bool take_abs, double span)
{
    double r_par = mul_rn(sub_rn(rc1, rc2), ch);
    if (take_abs) r_par = fabs(r_par);
    const double r_trans = mul_rn(add_rn(rc1, rc2), sh);   // r_comov, not dist_m (cf.py:1567)
    if (!((r_par < P.r_par_max) && (r_trans < P.r_trans_max) && (r_par >= P.r_par_min))) return -1;
    const int bp = (int)mul_rn(div_rn(sub_rn(r_par, P.r_par_min), span), (double)P.num_bins_r_par);
    const int bt = (int)mul_rn(div_rn(r_trans, P.r_trans_max), (double)P.num_bins_r_trans);
    // rounding onto the upper edge has no bin (the reference would index out of bounds): drop
    if (bp >= P.num_bins_r_par || bt >= P.num_bins_r_trans) return -1;
    return bt + P.num_bins_r_trans * bp;
}

// A[r][c] = (w[c] w[r]) xi1d(|ll[c] - ll[r]|) sqrt(v[c] v[r])   (cf.py:1413-1421)
__device__ __forceinline__ void wk_fill_wxi(double *__restrict__ A, int n,
                                            const double *__restrict__ w,
                                            const double *__restrict__ ll,
                                            const double *__restrict__ var, const WickTable &T)
{
    for (long long idx = threadIdx.x; idx < (long long)n * n; idx += blockDim.x) {
        const int r = (int)(idx / n), c = (int)(idx - (long long)r * n);
        const double x = fabs(sub_rn(ll[c], ll[r]));
        A[idx] = mul_rn(mul_rn(mul_rn(w[c], w[r]), wk_nearest(T, x)), sqrt(mul_rn(var[c], var[r])));
    }
}

// per row j of P: [lo, hi) bounding the valid entries (lo = hi = 0 when none)
__device__ __forceinline__ void wk_row_ranges(const int *__restrict__ Pm, int n_rows, int n1,
                                              int *__restrict__ lo, int *__restrict__ hi)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = warp; j < n_rows; j += WK_WARPS) {
        int mn = n1, mx = -1;
        for (int i = lane; i < n1; i += 32)
            if (Pm[(long long)j * n1 + i] >= 0) {
                mn = min(mn, i);
                mx = max(mx, i);
            }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        }
        if (lane == 0) {
            lo[j] = mx < 0 ? 0 : mn;
            hi[j] = mx < 0 ? 0 : mx + 1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// forest x forest: T1, T2, T3
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WK_THREADS)
pb2_wick_auto_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr,
                     const double *__restrict__ var1, const double *__restrict__ ze1,
                     const double *__restrict__ var2, const double *__restrict__ ze2,
                     WickTable X1, WickTable X2, WickWork W, double *__restrict__ weights_wick,
                     unsigned long long *__restrict__ num_pairs_wick, double *__restrict__ t1,
                     double *__restrict__ t2, double *__restrict__ t3)
{
    __shared__ long long s_e;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nb = (long long)P.num_bins_r_par * P.num_bins_r_trans;
    const double span = sub_rn(P.r_par_max, P.r_par_min);
    char *slab = W.slab_base + (long long)blockIdx.x * W.slab_stride;
    double *A = (double *)slab;
    double *B = A + W.a_cap;
    int *Pm = (int *)(B + W.b_cap);
    int *lo = Pm + W.p_cap;
    int *hi = lo + W.rows_cap;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            long long e;
            for (;;) {  // next kept forest pair
                e = (long long)atomicAdd(W.counter, 1ull);
                if (e >= pr.n_pairs || !pr.nb_keep || pr.nb_keep[e]) break;
            }
            s_e = e;
        }
        __syncthreads();
        const long long e = s_e;
        if (e >= pr.n_pairs) break;
        const int f1 = pr.f1_index[pr.nb_f1[e]], f2 = pr.nb_f2[e];
        const long long a1 = c1.offset[f1], a2 = c2.offset[f2];
        const int n1 = (int)(c1.offset[f1 + 1] - a1), n2 = (int)(c2.offset[f2 + 1] - a2);
        if (n1 <= 0 || n2 <= 0) continue;
        const double ch = pr.nb_cos[e], sh = pr.nb_sin[e];
        const double *__restrict__ rc1 = c1.r_comov + a1, *__restrict__ rc2 = c2.r_comov + a2;
        const double *__restrict__ w1 = c1.weights + a1, *__restrict__ w2 = c2.weights + a2;
        // ---- phase 0: bins + the O(|L|) terms
        for (long long idx = threadIdx.x; idx < (long long)n2 * n1; idx += blockDim.x) {
            const int j = (int)(idx / n1), i = (int)(idx - (long long)j * n1);
            const int p = wk_bin(P, rc1[i], rc2[j], ch, sh, !P.x_correlation, span);
            Pm[idx] = p;
            if (p >= 0) {
                const double w12 = mul_rn(w1[i], w2[j]);
                atomicAdd(weights_wick + p, w12);                                 // cf.py:1602
                atomicAdd(num_pairs_wick + p, 1ull);                              // :1603
                atomicAdd(t1 + p * nb + p, mul_rn(w12, mul_rn(ze1[a1 + i], ze2[a2 + j])));  // :1604
            }
        }
        wk_fill_wxi(A, n1, w1, c1.log_lambda + a1, var1 + a1, X1);
        wk_fill_wxi(B, n2, w2, c2.log_lambda + a2, var2 + a2, X2);
        __syncthreads();
        wk_row_ranges(Pm, n2, n1, lo, hi);
        __syncthreads();
        // ---- phase 1: pairs of list entries a < b (list order: j outer, i inner)
        for (long long ia = warp; ia < (long long)n2 * n1; ia += WK_WARPS) {
            const int pa = Pm[ia];
            if (pa < 0) continue;
            const int j = (int)(ia / n1), i = (int)(ia - (long long)j * n1);
            const double *__restrict__ Arow = A + (long long)i * n1;
            const double *__restrict__ Brow = B + (long long)j * n2;
            // same forest-2 pixel, later forest-1 pixels: t2 (cf.py:1614-1617)
            {
                const int *__restrict__ Prow = Pm + (long long)j * n1;
                for (int ib = i + 1; ib < hi[j]; ib += 32) {
                    const int i2 = ib + lane;
                    int key = -1;
                    double v = 0.;
                    if (i2 < hi[j]) {
                        key = Prow[i2];
                        // weight2[index2] is w2[j] here (same forest-2 pixel)
                        v = mul_rn(mul_rn(Arow[i2], w2[j]), ze2[a2 + j]);
                    }
                    wk_flush(t2, nb, pa, key, v, lane);
                }
            }
            for (int j2 = j + 1; j2 < n2; ++j2) {
                const int l0 = lo[j2], h0 = hi[j2];
                if (h0 <= l0) continue;
                const double bjj = Brow[j2];
                const int *__restrict__ Prow = Pm + (long long)j2 * n1;
                for (int ib = l0; ib < h0; ib += 32) {
                    const int i2 = ib + lane;
                    int key = -1;
                    double v = 0.;
                    if (i2 < h0) {
                        key = Prow[i2];
                        if (i2 == i) {
                            // same forest-1 pixel: t2 (cf.py:1610-1613), one entry per row
                            if (key >= 0) {
                                const double v2 = mul_rn(mul_rn(bjj, w1[i]), ze1[a1 + i]);
                                atomicAdd(t2 + (long long)pa * nb + key, v2);
                                atomicAdd(t2 + (long long)key * nb + pa, v2);
                            }
                        } else {
                            v = mul_rn(Arow[i2], bjj);                        // cf.py:1619-1621
                        }
                    }
                    wk_flush(t3, nb, pa, key, v, lane);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// forest x objects: T1, T2, T3, T4
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WK_THREADS)
pb2_wick_cross_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr,
                      const uint8_t *__restrict__ keep_f1, const double *__restrict__ var1,
                      const double *__restrict__ ze1, const double *__restrict__ ze_obj,
                      WickTable X1, WickWork W, double *__restrict__ weights_wick,
                      unsigned long long *__restrict__ num_pairs_wick, double *__restrict__ t1,
                      double *__restrict__ t2, double *__restrict__ t3, double *__restrict__ t4)
{
    __shared__ long long s_k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nb = (long long)P.num_bins_r_par * P.num_bins_r_trans;
    const double span = sub_rn(P.r_par_max, P.r_par_min);
    char *slab = W.slab_base + (long long)blockIdx.x * W.slab_stride;
    double *A = (double *)slab;
    int *Pm = (int *)(A + W.a_cap);
    int *lo = Pm + W.p_cap;
    int *hi = lo + W.rows_cap;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            long long k;
            for (;;) {  // next kept forest with neighbours
                k = (long long)atomicAdd(W.counter, 1ull);
                if (k >= pr.n_f1 || (keep_f1[k] && pr.nb_offset[k + 1] > pr.nb_offset[k])) break;
            }
            s_k = k;
        }
        __syncthreads();
        const long long k = s_k;
        if (k >= pr.n_f1) break;
        const int f1 = pr.f1_index[k];
        const long long a1 = c1.offset[f1];
        const int n1 = (int)(c1.offset[f1 + 1] - a1);
        const long long e0 = pr.nb_offset[k];
        const int n2 = (int)(pr.nb_offset[k + 1] - e0);   // neighbouring objects
        if (n1 <= 0) continue;
        const double *__restrict__ rc1 = c1.r_comov + a1, *__restrict__ w1 = c1.weights + a1;
        // ---- phase 0
        for (long long idx = threadIdx.x; idx < (long long)n2 * n1; idx += blockDim.x) {
            const int j = (int)(idx / n1), i = (int)(idx - (long long)j * n1);
            const int q = pr.nb_f2[e0 + j];
            const long long aq = c2.offset[q];
            const int p = wk_bin(P, rc1[i], c2.r_comov[aq], pr.nb_cos[e0 + j], pr.nb_sin[e0 + j],
                                 false, span);
            Pm[idx] = p;
            if (p >= 0) {
                const double w12 = mul_rn(w1[i], c2.weights[aq]);
                atomicAdd(weights_wick + p, w12);                                // xcf.py:1314
                atomicAdd(num_pairs_wick + p, 1ull);
                // weights12**2 / weight1 * z_weight_evol1 (xcf.py:1316)
                atomicAdd(t1 + p * nb + p, mul_rn(div_rn(mul_rn(w12, w12), w1[i]), ze1[a1 + i]));
            }
        }
        wk_fill_wxi(A, n1, w1, c1.log_lambda + a1, var1 + a1, X1);
        __syncthreads();
        wk_row_ranges(Pm, n2, n1, lo, hi);
        __syncthreads();
        // ---- phase 1
        for (long long ia = warp; ia < (long long)n2 * n1; ia += WK_WARPS) {
            const int pa = Pm[ia];
            if (pa < 0) continue;
            const int j = (int)(ia / n1), i = (int)(ia - (long long)j * n1);
            const double *__restrict__ Arow = A + (long long)i * n1;
            const int q = pr.nb_f2[e0 + j];
            const double zq = ze_obj[q];
            const double w12a = mul_rn(w1[i], c2.weights[c2.offset[q]]);
            // same object, later pixels: t2 (xcf.py:1322-1325)
            {
                const double zz = mul_rn(zq, zq);
                const int *__restrict__ Prow = Pm + (long long)j * n1;
                for (int ib = i + 1; ib < hi[j]; ib += 32) {
                    const int i2 = ib + lane;
                    int key = -1;
                    double v = 0.;
                    if (i2 < hi[j]) {
                        key = Prow[i2];
                        v = mul_rn(Arow[i2], zz);
                    }
                    wk_flush(t2, nb, pa, key, v, lane);
                }
            }
            for (int j2 = j + 1; j2 < n2; ++j2) {
                const int l0 = lo[j2], h0 = hi[j2];
                if (h0 <= l0) continue;
                const int q2 = pr.nb_f2[e0 + j2];
                const double zq2 = ze_obj[q2];
                const int *__restrict__ Prow = Pm + (long long)j2 * n1;
                for (int ib = l0; ib < h0; ib += 32) {
                    const int i2 = ib + lane;
                    int key = -1;
                    double v = 0.;
                    if (i2 < h0) {
                        key = Prow[i2];
                        if (i2 == i) {
                            // same pixel, another object: t3 (xcf.py:1326-1334)
                            if (key >= 0) {
                                const double w12b = mul_rn(w1[i], c2.weights[c2.offset[q2]]);
                                const double v3 =
                                    mul_rn(div_rn(mul_rn(w12a, w12b), w1[i]), ze1[a1 + i]);
                                atomicAdd(t3 + (long long)pa * nb + key, v3);
                                atomicAdd(t3 + (long long)key * nb + pa, v3);
                            }
                        } else {
                            v = mul_rn(mul_rn(Arow[i2], zq), zq2);            // xcf.py:1336-1340
                        }
                    }
                    wk_flush(t4, nb, pa, key, v, lane);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
static const int WK_BLOCKS = 148 * 2;

static long long wick_slab_bytes(long long n1_max, long long rows_max, bool cross)
{
    long long b = n1_max * n1_max * 8;                 // A
    if (!cross) b += rows_max * rows_max * 8;          // B
    b += rows_max * n1_max * 4;                        // P
    b += 2 * rows_max * 4;                             // lo, hi
    return (b + 255) / 256 * 256;
}

extern "C" {

int64_t pb2_wick_scratch_bytes(int64_t max_pix1, int64_t max_rows2, int32_t cross)
{
    if (max_pix1 < 1) max_pix1 = 1;
    if (max_rows2 < 1) max_rows2 = 1;
    return 256 + (long long)WK_BLOCKS * wick_slab_bytes(max_pix1, max_rows2, cross != 0);
}

static void wick_work(WickWork *W, void *d_scratch, long long n1_max, long long rows_max, bool cross)
{
    W->counter = (unsigned long long *)d_scratch;
    W->slab_base = (char *)d_scratch + 256;
    W->slab_stride = wick_slab_bytes(n1_max, rows_max, cross);
    W->a_cap = n1_max * n1_max;
    W->b_cap = cross ? 0 : rows_max * rows_max;
    W->p_cap = rows_max * n1_max;
    W->rows_cap = (int)rows_max;
}

int32_t pb2_wick_auto(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                      const pb2_pairs *pairs, const double *d_var1, const double *d_ze1,
                      const double *d_var2, const double *d_ze2, int32_t n_x1,
                      const double *d_xb1, const double *d_xy1, int32_t n_x2,
                      const double *d_xb2, const double *d_xy2, double *d_weights_wick,
                      int64_t *d_num_pairs_wick, double *d_t1, double *d_t2, double *d_t3,
                      void *d_scratch, int64_t scratch_bytes, void *stream)
{
    if (!cat1 || !cat2 || !par || !pairs || !d_var1 || !d_ze1 || !d_var2 || !d_ze2 || !d_xy1 ||
        !d_xy2 || !d_weights_wick || !d_num_pairs_wick || !d_t1 || !d_t2 || !d_t3 || !d_scratch) {
        pb2_set_error("pb2_wick_auto: null pointer argument");
        return PB2_EINVAL;
    }
    if (par->rmu_binning || par->ang_correlation) {
        pb2_set_error("pb2_wick_auto: the reference has no rmu / angular Wick expansion");
        return PB2_ECONFIG;
    }
    if (!cat1->log_lambda || !cat2->log_lambda || n_x1 < 1 || n_x2 < 1 ||
        (n_x1 > 1 && !d_xb1) || (n_x2 > 1 && !d_xb2)) {
        pb2_set_error("pb2_wick_auto: missing log_lambda or 1-D correlation table");
        return PB2_EINVAL;
    }
    if (pairs->n_pairs <= 0) return 0;
    const long long need = pb2_wick_scratch_bytes(cat1->max_pix, cat2->max_pix, 0);
    if (scratch_bytes < need) {
        pb2_set_error("pb2_wick_auto: scratch too small (%lld < %lld bytes)",
                      (long long)scratch_bytes, need);
        return PB2_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    WickWork W;
    wick_work(&W, d_scratch, cat1->max_pix > 0 ? cat1->max_pix : 1,
              cat2->max_pix > 0 ? cat2->max_pix : 1, false);
    WickTable X1 = {n_x1, d_xb1, d_xy1}, X2 = {n_x2, d_xb2, d_xy2};
    PB2_CUDA(cudaMemsetAsync(d_scratch, 0, 256, s));
    pb2_timing_begin(s);
    pb2_wick_auto_kernel<<<WK_BLOCKS, WK_THREADS, 0, s>>>(
        *cat1, *cat2, *par, *pairs, d_var1, d_ze1, d_var2, d_ze2, X1, X2, W, d_weights_wick,
        (unsigned long long *)d_num_pairs_wick, d_t1, d_t2, d_t3);
    pb2_count_launch(1);
    int32_t rc = pb2_check_launch("pb2_wick_auto_kernel");
    pb2_timing_end(s);
    return rc;
}

int32_t pb2_wick_cross(const pb2_catalog *cat1, const pb2_catalog *objs, const pb2_params *par,
                       const pb2_pairs *pairs, const uint8_t *d_keep_f1, int64_t max_neighbours,
                       const double *d_var1, const double *d_ze1, const double *d_ze_obj,
                       int32_t n_x1, const double *d_xb1, const double *d_xy1,
                       double *d_weights_wick, int64_t *d_num_pairs_wick, double *d_t1,
                       double *d_t2, double *d_t3, double *d_t4, void *d_scratch,
                       int64_t scratch_bytes, void *stream)
{
    if (!cat1 || !objs || !par || !pairs || !d_keep_f1 || !d_var1 || !d_ze1 || !d_ze_obj ||
        !d_xy1 || !d_weights_wick || !d_num_pairs_wick || !d_t1 || !d_t2 || !d_t3 || !d_t4 ||
        !d_scratch) {
        pb2_set_error("pb2_wick_cross: null pointer argument");
        return PB2_EINVAL;
    }
    if (par->rmu_binning || par->ang_correlation) {
        pb2_set_error("pb2_wick_cross: the reference has no rmu / angular Wick expansion");
        return PB2_ECONFIG;
    }
    if (!cat1->log_lambda || n_x1 < 1 || (n_x1 > 1 && !d_xb1)) {
        pb2_set_error("pb2_wick_cross: missing log_lambda or 1-D correlation table");
        return PB2_EINVAL;
    }
    if (pairs->n_pairs <= 0 || pairs->n_f1 <= 0) return 0;
    const long long need = pb2_wick_scratch_bytes(cat1->max_pix, max_neighbours, 1);
    if (scratch_bytes < need) {
        pb2_set_error("pb2_wick_cross: scratch too small (%lld < %lld bytes)",
                      (long long)scratch_bytes, need);
        return PB2_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    WickWork W;
    wick_work(&W, d_scratch, cat1->max_pix > 0 ? cat1->max_pix : 1,
              max_neighbours > 0 ? max_neighbours : 1, true);
    WickTable X1 = {n_x1, d_xb1, d_xy1};
    PB2_CUDA(cudaMemsetAsync(d_scratch, 0, 256, s));
    pb2_timing_begin(s);
    pb2_wick_cross_kernel<<<WK_BLOCKS, WK_THREADS, 0, s>>>(
        *cat1, *objs, *par, *pairs, d_keep_f1, d_var1, d_ze1, d_ze_obj, X1, W, d_weights_wick,
        (unsigned long long *)d_num_pairs_wick, d_t1, d_t2, d_t3, d_t4);
    pb2_count_launch(1);
    int32_t rc = pb2_check_launch("pb2_wick_cross_kernel");
    pb2_timing_end(s);
    return rc;
}

}  // extern "C"
