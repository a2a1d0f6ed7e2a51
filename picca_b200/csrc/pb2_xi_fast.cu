// Auto / delta x delta pixel-pair histogram -- the specialised product kernel for the standard
// (r_par, r_trans) binning without per-pair cuts (what picca_cf.py runs by default).
// Replaces cf.compute_xi's pair loop + cf.compute_xi_forest_pairs_fast
// (reference py/picca/cf.py:161-240, 250-387).  Same algorithm as pb2_xi_auto_tiled in pb2_xi.cu
// (diagonal sweep, sandwiched bins, two warp-uniform run slots per row set, transposed flush);
// here the hot loop is written on named scalars for exactly two row sets, with the validity of a
// row / column folded into integer masks, so that one step of a row set costs ~14 FP64 and ~12
// integer / predicate instructions.  Everything that is not the default mode (rmu / angular
// binning, z cuts, zerr cut, half-plate removal, unsorted forests) goes to pb2_xi_auto_tiled.
#include "pb2_common.cuh"

#define XF_THREADS 384
#define XF_CHUNK 8
#define XF_MAGIC 6755399441055744.0  // 2^52 + 2^51

struct XfConst {
    double kp_lo, kp_hi, kt_lo, kt_hi, magic;
    int tmax;
};

__device__ __forceinline__ int xf_lower_bound(const double *__restrict__ a, int n, double v,
                                              bool strict, int lane)
{
    // number of elements of the non-decreasing a[0..n) that are < v (strict) or <= v
    int lo = 0, hi = n;
    while (hi - lo > 32) {
        const int len = hi - lo;
        const int p = lo + (int)(((long long)(lane + 1) * len) / 33);
        const double x = __ldg(a + p);
        const bool below = strict ? (x < v) : (x <= v);
        const int c = __popc(__ballot_sync(0xffffffffu, below));
        const int p_prev = lo + (int)(((long long)c * len) / 33);
        const int p_next = lo + (int)(((long long)(c + 1) * len) / 33);
        if (c > 0) lo = p_prev + 1;
        if (c < 32) hi = p_next;
    }
    const int p = lo + lane;
    bool below = false;
    if (p < hi) {
        const double x = __ldg(a + p);
        below = strict ? (x < v) : (x <= v);
    }
    return lo + __popc(__ballot_sync(0xffffffffu, below));
}

__device__ __forceinline__ double xf_shfl_xor(double v, int m)
{
    return __shfl_xor_sync(0xffffffffu, v, m);
}

// Sum a0..a4 over the warp; lane 4*v (v = 0..4) ends up with the total of a_v.
__device__ __forceinline__ double xf_reduce5(double a0, double a1, double a2, double a3, double a4,
                                             int lane)
{
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    double k0 = b4 ? a4 : a0, k1 = b4 ? 0. : a1, k2 = b4 ? 0. : a2, k3 = b4 ? 0. : a3;
    const double s0 = b4 ? a0 : a4, s1 = b4 ? a1 : 0., s2 = b4 ? a2 : 0., s3 = b4 ? a3 : 0.;
    k0 += xf_shfl_xor(s0, 16);
    k1 += xf_shfl_xor(s1, 16);
    k2 += xf_shfl_xor(s2, 16);
    k3 += xf_shfl_xor(s3, 16);
    double u0 = b3 ? k2 : k0, u1 = b3 ? k3 : k1;
    const double t0 = b3 ? k0 : k2, t1 = b3 ? k1 : k3;
    u0 += xf_shfl_xor(t0, 8);
    u1 += xf_shfl_xor(t1, 8);
    double w = b2 ? u1 : u0;
    const double x = b2 ? u0 : u1;
    w += xf_shfl_xor(x, 4);
    w += xf_shfl_xor(w, 2);
    w += xf_shfl_xor(w, 1);
    return w;
}

// flush one slot (factored sums, see pb2_xi.cu) into bin `key` of the output row
__device__ __noinline__ void xf_flush(double sw, double sdw, double srp, double srt, double szw,
                                      int cnt, int key, double w1, double dw1, double z1,
                                      double *__restrict__ orow, int nb)
{
    const int lane = threadIdx.x & 31;
    const double we = w1 * sw;
    const double tot = xf_reduce5(we, dw1 * sdw, w1 * srp, w1 * srt, 0.5 * (z1 * we + w1 * szw), lane);
    const int c = __reduce_add_sync(0xffffffffu, cnt);
    if (lane < 20 && (lane & 3) == 0) atomic_add_f64(orow + (size_t)(lane >> 2) * nb + key, tot);
    else if (lane == 20) atomic_add_i64(orow + (size_t)5 * nb + key, (long long)c);
}

// ---- per row set state on named scalars (r = 0, 1)
#define XF_DECL(r)                                                                       \
    double rc1_##r, dm1_##r, z1_##r, w1_##r, dw1_##r;                                    \
    int m1_##r;              /* -1 if the row exists and has weight, else 0 */           \
    int jl_##r, jh_##r;      /* column window of the row set */                          \
    int ka_##r = -1, kb_##r = -1;                                                        \
    double asw_##r = 0., asdw_##r = 0., asrp_##r = 0., asrt_##r = 0., aszw_##r = 0.;     \
    double bsw_##r = 0., bsdw_##r = 0., bsrp_##r = 0., bsrt_##r = 0., bszw_##r = 0.;     \
    int acnt_##r = 0, bcnt_##r = 0;

#define XF_LOAD_ROW(r)                                                                   \
    {                                                                                    \
        const int i = i0 + 32 * r + lane;                                                \
        const bool ok = i < n1;                                                          \
        const long long p = a + (ok ? i : 0);                                            \
        rc1_##r = __ldg(c1.r_comov + p);                                                 \
        dm1_##r = __ldg(c1.dist_m + p);                                                  \
        z1_##r = __ldg(c1.z + p);                                                        \
        w1_##r = __ldg(c1.weights + p);                                                  \
        dw1_##r = __ldg(c1.delta_w + p);                                                 \
        m1_##r = (ok && (w1_##r != 0.)) ? -1 : 0; /* cf.py:318 */                        \
    }

#define XF_WINDOW(r)                                                                     \
    {                                                                                    \
        const int ifirst = i0 + 32 * r;                                                  \
        jl_##r = n2;                                                                     \
        jh_##r = 0;                                                                      \
        if (ifirst < n1) {                                                               \
            const int ilast = min(ifirst + 31, n1 - 1);                                  \
            const double rc_first = __ldg(c1.r_comov + a + ifirst);                      \
            const double rc_last = __ldg(c1.r_comov + a + ilast);                        \
            const double dm_first = __ldg(c1.dist_m + a + ifirst);                       \
            const int lo = xf_lower_bound(p_rc2, n2, rc_first - dmax, false, lane);      \
            int hi = xf_lower_bound(p_rc2, n2, rc_last - dlow, true, lane);              \
            if (isfinite(tsum))                                                          \
                hi = min(hi, xf_lower_bound(p_dm2, n2, tsum - dm_first, true, lane));    \
            jl_##r = lo;                                                                 \
            jh_##r = hi;                                                                 \
            if (hi > lo) {                                                               \
                JL = min(JL, lo);                                                        \
                JH = max(JH, hi);                                                        \
            }                                                                            \
        }                                                                                \
    }

#define XF_FLUSH_A(r)                                                                    \
    xf_flush(asw_##r, asdw_##r, asrp_##r, asrt_##r, aszw_##r, acnt_##r, ka_##r, w1_##r,  \
             dw1_##r, z1_##r, orow, nb);                                                 \
    asw_##r = asdw_##r = asrp_##r = asrt_##r = aszw_##r = 0.;                            \
    acnt_##r = 0;

#define XF_FLUSH_B(r)                                                                    \
    xf_flush(bsw_##r, bsdw_##r, bsrp_##r, bsrt_##r, bszw_##r, bcnt_##r, kb_##r, w1_##r,  \
             dw1_##r, z1_##r, orow, nb);                                                 \
    bsw_##r = bsdw_##r = bsrp_##r = bsrt_##r = bszw_##r = 0.;                            \
    bcnt_##r = 0;

// one column against one row set.  `m2` = -1 if the column exists and has weight.
// Common case: every lane's pair falls in slot A or slot B (the warp sits on one bin, or straddles
// one bin edge); only a bin that matches neither slot, or a borderline bin, takes the slow branch.
#define XF_PAIR(r)                                                                       \
    if (!(j0 + 31 < jl_##r || j0 >= jh_##r)) {                                           \
        double rp = mul_rn(sub_rn(rc1_##r, rc2), ch);                                    \
        if (!XCORR) rp = fabs(rp);                                                       \
        const double rt = mul_rn(add_rn(dm1_##r, dm2), sh);                              \
        const double x = sub_rn(rp, rpmin);                                              \
        const int bpl = __double2loint(__fma_rd(x, K.kp_lo, magic));                     \
        const int bph = __double2loint(__fma_rd(x, K.kp_hi, magic));                     \
        const int btl = __double2loint(__fma_rd(rt, K.kt_lo, magic));                    \
        const int bth = __double2loint(__fma_rd(rt, K.kt_hi, magic));                    \
        const int both = m1_##r & m2;                                                    \
        const bool sure = (bpl == bph) && (btl == bth);                                  \
        bool in = sure && ((unsigned)(bpl | ~both) < np_u) && ((unsigned)btl < nt_u);    \
        int bin = btl + (int)nt_u * bpl;                                                 \
        bool is_a = in && (bin == ka_##r);                                               \
        bool is_b = in && (bin == kb_##r);                                               \
        if (__any_sync(0xffffffffu, (in && !is_a && !is_b) || (both && !sure))) {        \
            if (both && !sure) { /* reference expression, true divisions */              \
                const PairGeom g = pb2_pair_exact(P, rc1_##r, dm1_##r, rc2, dm2, ang, ch, \
                                                  sh, false, false);                     \
                in = g.bin >= 0;                                                         \
                bin = g.bin;                                                             \
            }                                                                            \
            for (;;) {                                                                   \
                is_a = in && (bin == ka_##r);                                            \
                is_b = in && (bin == kb_##r);                                            \
                const unsigned other = __ballot_sync(0xffffffffu, in && !is_a && !is_b); \
                if (!other) break;                                                       \
                const int key = __shfl_sync(0xffffffffu, bin, __ffs(other) - 1);         \
                const unsigned ma = __ballot_sync(0xffffffffu, is_a);                    \
                const unsigned mb = __ballot_sync(0xffffffffu, is_b);                    \
                if (ka_##r < 0) {                                                        \
                    ka_##r = key;                                                        \
                } else if (kb_##r < 0) {                                                 \
                    kb_##r = key;                                                        \
                } else if (!ma) {                                                        \
                    XF_FLUSH_A(r)                                                        \
                    ka_##r = key;                                                        \
                } else if (!mb) {                                                        \
                    XF_FLUSH_B(r)                                                        \
                    kb_##r = key;                                                        \
                } else if (in && bin == key) { /* three bins live: add the pair directly */ \
                    const double w12 = mul_rn(w1_##r, w2);                               \
                    atomic_add_f64(orow + 0 * (size_t)nb + bin, w12);                    \
                    atomic_add_f64(orow + 1 * (size_t)nb + bin, mul_rn(dw1_##r, dw2));   \
                    atomic_add_f64(orow + 2 * (size_t)nb + bin, mul_rn(rp, w12));        \
                    atomic_add_f64(orow + 3 * (size_t)nb + bin, mul_rn(rt, w12));        \
                    atomic_add_f64(orow + 4 * (size_t)nb + bin,                          \
                                   0.5 * (z1_##r * w12 + w1_##r * zw2));                 \
                    atomic_add_i64(orow + 5 * (size_t)nb + bin, 1);                      \
                    in = false;                                                          \
                }                                                                        \
            }                                                                            \
        }                                                                                \
        if (is_a) {                                                                      \
            asw_##r += w2;                                                               \
            asdw_##r += dw2;                                                             \
            aszw_##r += zw2;                                                             \
            asrp_##r = fma(rp, w2, asrp_##r);                                            \
            asrt_##r = fma(rt, w2, asrt_##r);                                            \
            acnt_##r += 1;                                                               \
        }                                                                                \
        if (is_b) {                                                                      \
            bsw_##r += w2;                                                               \
            bsdw_##r += dw2;                                                             \
            bszw_##r += zw2;                                                             \
            bsrp_##r = fma(rp, w2, bsrp_##r);                                            \
            bsrt_##r = fma(rt, w2, bsrt_##r);                                            \
            bcnt_##r += 1;                                                               \
        }                                                                                \
    }

template <bool XCORR>
__global__ void __launch_bounds__(XF_THREADS, 1)
pb2_xi_auto_fast(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, XfConst K,
                 const int32_t *__restrict__ out_row, double *__restrict__ out)
{
    __shared__ unsigned s_ctr;
    if (threadIdx.x == 0) s_ctr = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const unsigned np_u = (unsigned)P.num_bins_r_par, nt_u = (unsigned)P.num_bins_r_trans;
    const unsigned tmax = (unsigned)K.tmax;
    const unsigned units_per_chunk = XF_CHUNK * tmax;
    const double magic = K.magic;
    const double rpmin = P.r_par_min;

    for (;;) {
        unsigned u = 0;
        if (lane == 0) u = atomicAdd(&s_ctr, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        const long long chunk = (long long)blockIdx.x + (long long)(u / units_per_chunk) * gridDim.x;
        if (chunk * XF_CHUNK >= pr.n_pairs) break;
        const unsigned local = u % units_per_chunk;
        const long long e = chunk * XF_CHUNK + local / tmax;
        if (e >= pr.n_pairs) continue;
        const int tile = (int)(local % tmax);

        const int k = pr.nb_f1[e];
        const int f1 = pr.f1_index[k];
        const int f2 = pr.nb_f2[e];
        const long long a = c1.offset[f1];
        const int n1 = (int)(c1.offset[f1 + 1] - a);
        const int i0 = tile * 64;
        if (i0 >= n1) continue;
        const long long b = c2.offset[f2];
        const int n2 = (int)(c2.offset[f2 + 1] - b);
        if (n2 == 0) continue;
        const double ang = pr.nb_ang[e];
        const double ch = pr.nb_cos[e], sh = pr.nb_sin[e];
        double *__restrict__ orow = out + (size_t)out_row[k] * 6 * nb;
        const double *__restrict__ p_rc2 = c2.r_comov + b;
        const double *__restrict__ p_dm2 = c2.dist_m + b;
        const double *__restrict__ p_w2 = c2.weights + b;
        const double *__restrict__ p_dw2 = c2.delta_w + b;
        const double *__restrict__ p_zw2 = c2.z_w + b;

        XF_DECL(0)
        XF_DECL(1)
        XF_LOAD_ROW(0)
        XF_LOAD_ROW(1)

        // column windows (conservative supersets; the exact test still runs on every pair)
        int JL = n2, JH = 0;
        {
            const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
            const double dmax = P.r_par_max * inv_c * (1. + 1e-9) + 1e-9;
            const double dmin = P.r_par_min * inv_c;
            const double dlow = XCORR ? (dmin - fabs(dmin) * 1e-9 - 1e-9) : -dmax;
            const double tsum = P.r_trans_max * inv_s * (1. + 1e-9) + 1e-9;
            XF_WINDOW(0)
            XF_WINDOW(1)
        }
        if (JH <= JL) continue;

        const int nsteps = JH - JL + 31;
        int j = JL - 31 + lane;
        for (int s = 0; s < nsteps; s++, j++) {
            const int j0 = j - lane;  // column of lane 0 (warp-uniform)
            const int jc = min(max(j, 0), n2 - 1);
            const double rc2 = __ldg(p_rc2 + jc), dm2 = __ldg(p_dm2 + jc);
            const double w2 = __ldg(p_w2 + jc), dw2 = __ldg(p_dw2 + jc), zw2 = __ldg(p_zw2 + jc);
            const int m2 = ((j == jc) && (w2 != 0.)) ? -1 : 0;  // cf.py:331
            XF_PAIR(0)
            XF_PAIR(1)
        }
        if (ka_0 >= 0) { XF_FLUSH_A(0) }
        if (kb_0 >= 0) { XF_FLUSH_B(0) }
        if (ka_1 >= 0) { XF_FLUSH_A(1) }
        if (kb_1 >= 0) { XF_FLUSH_B(1) }
    }
}

int32_t pb2_launch_xi_fast(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                           const pb2_pairs *pairs, const int32_t *d_out_row, double *d_out,
                           cudaStream_t s)
{
    XfConst K;
    const double kp = (double)par->num_bins_r_par / (par->r_par_max - par->r_par_min);
    const double kt = (double)par->num_bins_r_trans / par->r_trans_max;
    const double eps = 9.094947017729282e-13;  // 2^-40
    K.kp_lo = kp * (1. - eps);
    K.kp_hi = kp * (1. + eps);
    K.kt_lo = kt * (1. - eps);
    K.kt_hi = kt * (1. + eps);
    K.magic = XF_MAGIC;
    K.tmax = (c1->max_pix + 63) / 64;
    if (K.tmax < 1) K.tmax = 1;
    int dev = 0, sms = 0;
    PB2_CUDA(cudaGetDevice(&dev));
    PB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long want = (pairs->n_pairs + XF_CHUNK - 1) / XF_CHUNK;
    int blocks = (int)(want < sms ? want : sms);
    if (blocks < 1) blocks = 1;
    if (par->x_correlation)
        pb2_xi_auto_fast<true><<<blocks, XF_THREADS, 0, s>>>(*c1, *c2, *par, *pairs, K, d_out_row, d_out);
    else
        pb2_xi_auto_fast<false><<<blocks, XF_THREADS, 0, s>>>(*c1, *c2, *par, *pairs, K, d_out_row, d_out);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_xi_auto_fast");
}
