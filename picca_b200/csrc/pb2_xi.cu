// Forest x forest pixel-pair histogram: replaces cf.compute_xi's pair loop and
// cf.compute_xi_forest_pairs_fast (reference py/picca/cf.py:161-240, 250-387).
//
// Product kernel (variant 0): "diagonal sweep".
//   work unit  = (forest pair e, tile of 32*R rows of forest 1), claimed by one warp;
//   lane l     = rows i0 + 32 r + l (r < R) of forest 1, held in registers;
//   step s     = every lane reads column j = JL - 31 + l + s of forest 2 (coalesced), so that at a
//                given step all lanes sit on (nearly) the same diagonal i - j: r_par and r_trans,
//                hence the (r_par, r_trans) bin, are almost always warp-uniform and change every
//                ~6 steps.  Each lane accumulates a run of same-bin pairs in registers; finished
//                runs are parked and flushed by a warp-cooperative transposed reduction that ends
//                in ONE red.global.add.f64 instruction (5 lanes, 5 sums) + one u64 atomic (count).
//   window     = r_comov / dist_m are sorted inside a forest, and fp subtraction/multiplication by
//                a positive constant are monotone, so the in-range columns of a row form one
//                interval; [JL, JH) is a conservative superset found by a warp-wide search and the
//                exact test still runs on every visited pair.
//   exact bins = the reference bins with floor((r-min)/(max-min)*np): two IEEE divisions per pair.
//                Here floor(x*K) is evaluated with DFMA.RM against 2^52+2^51 for K*(1-2^-40) and
//                K*(1+2^-40); when both agree the reference's value is sandwiched and the bin is
//                proven identical, otherwise (|frac| < ~1e-12, about one pair in 1e11) the lane
//                re-evaluates the reference expression with true divisions.
// Shared-memory fp64 atomics are CAS loops on sm_100a (ATOMS.CAST.SPIN), whereas global fp64
// reductions are native (REDG.E.ADD.F64), so the per-HEALPix histograms live in L2 and receive
// only warp-aggregated runs.
//
// Validation kernel (variant 1): brute force, one exact evaluation + 6 atomics per pair.
#include "pb2_common.cuh"

#define PB2_MAGIC 6755399441055744.0  // 2^52 + 2^51

struct XiFast {
    double kp_lo, kp_hi, kt_lo, kt_hi;  // bin scale factors bracketing np/(max-min), nt/rt_max
    int fast;                           // 1: windows + sandwiched bins are valid for this call
    int tmax;                           // tiles per forest pair (longest forest 1)
};

// ------------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------------
// number of elements of the non-decreasing array a[0..n) that are <  v (strict=1) or <= v (0)
__device__ __forceinline__ int warp_lower_bound(const double *__restrict__ a, int n, double v,
                                                bool strict, int lane)
{
    int lo = 0, hi = n;
    while (hi - lo > 32) {
        const int len = hi - lo;
        const int p = lo + (int)(((long long)(lane + 1) * len) / 33);
        const double x = __ldg(a + p);
        const bool below = strict ? (x < v) : (x <= v);
        const unsigned m = __ballot_sync(0xffffffffu, below);
        const int c = __popc(m);
        const int p_prev = lo + (int)(((long long)c * len) / 33);        // probe of lane c-1
        const int p_next = lo + (int)(((long long)(c + 1) * len) / 33);  // probe of lane c
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

__device__ __forceinline__ double shfl_xor_f64(double v, int m)
{
    return __shfl_xor_sync(0xffffffffu, v, m);
}

// Sum each of a0..a4 over the 32 lanes.  On return lane 4*v (v = 0..4) holds the total of a_v
// (all four lanes of quad v do).  9 shuffled doubles instead of 25 for five butterflies.
__device__ __forceinline__ double transpose_reduce5(double a0, double a1, double a2, double a3,
                                                    double a4, int lane)
{
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    // level 1 (xor 16): bit4 = 0 keeps slots (a0..a3), bit4 = 1 keeps (a4,0,0,0)
    double k0 = b4 ? a4 : a0, k1 = b4 ? 0. : a1, k2 = b4 ? 0. : a2, k3 = b4 ? 0. : a3;
    double s0 = b4 ? a0 : a4, s1 = b4 ? a1 : 0., s2 = b4 ? a2 : 0., s3 = b4 ? a3 : 0.;
    k0 += shfl_xor_f64(s0, 16);
    k1 += shfl_xor_f64(s1, 16);
    k2 += shfl_xor_f64(s2, 16);
    k3 += shfl_xor_f64(s3, 16);
    // level 2 (xor 8): bit3 = 0 keeps (k0,k1), bit3 = 1 keeps (k2,k3)
    double u0 = b3 ? k2 : k0, u1 = b3 ? k3 : k1;
    double t0 = b3 ? k0 : k2, t1 = b3 ? k1 : k3;
    u0 += shfl_xor_f64(t0, 8);
    u1 += shfl_xor_f64(t1, 8);
    // level 3 (xor 4): bit2 = 0 keeps u0, bit2 = 1 keeps u1
    double w = b2 ? u1 : u0;
    double x = b2 ? u0 : u1;
    w += shfl_xor_f64(x, 4);
    w += shfl_xor_f64(w, 2);
    w += shfl_xor_f64(w, 1);
    return w;
}

struct Run {
    double we, xi, rp, rt, z;
    int cnt;
    int key;  // flat bin, -1 = empty
};

__device__ __forceinline__ void run_clear(Run &r)
{
    r.we = r.xi = r.rp = r.rt = r.z = 0.;
    r.cnt = 0;
    r.key = -1;
}

// Flush the runs held by the lanes of a warp into the output row (one red.f64 per distinct key).
// out_row_ptr points at [6][nb].  Clears the runs.
__device__ __forceinline__ void flush_runs(Run &r, double *__restrict__ out_row_ptr, int nb,
                                           int lane)
{
    unsigned pending = __ballot_sync(0xffffffffu, r.key >= 0);
    while (pending) {
        const int leader = __ffs(pending) - 1;
        const int key = __shfl_sync(0xffffffffu, r.key, leader);
        const bool mine = (r.key == key);
        // r.z accumulates (z1 + z2) * w12; the reference's z = (z1 + z2) / 2 (cf.py:334)
        const double tot = transpose_reduce5(mine ? r.we : 0., mine ? r.xi : 0., mine ? r.rp : 0.,
                                             mine ? r.rt : 0., mine ? 0.5 * r.z : 0., lane);
        const int cnt = __reduce_add_sync(0xffffffffu, mine ? r.cnt : 0);
        if (lane < 20 && (lane & 3) == 0) {
            atomic_add_f64(out_row_ptr + (size_t)(lane >> 2) * nb + key, tot);
        } else if (lane == 20) {
            atomic_add_i64(out_row_ptr + (size_t)5 * nb + key, (long long)cnt);
        }
        if (mine) run_clear(r);
        pending = __ballot_sync(0xffffffffu, r.key >= 0);
    }
}

// ------------------------------------------------------------------------------------------
// product kernel
// ------------------------------------------------------------------------------------------
#define XI_THREADS 512
#define XI_CHUNK 16

template <int R, bool FAST>
__global__ void __launch_bounds__(XI_THREADS, 1)
pb2_xi_auto_tiled(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, XiFast F,
                  const int32_t *__restrict__ out_row, double *__restrict__ out,
                  unsigned long long *__restrict__ g_counter)
{
    __shared__ long long s_e0;
    __shared__ int s_ctr;
    const int lane = threadIdx.x & 31;
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const unsigned np_u = (unsigned)P.num_bins_r_par, nt_u = (unsigned)P.num_bins_r_trans;
    const int tmax = F.tmax;
    const double zerr_ang = mul_rn(P.zerr_cut_deg, PB2_PI) / 180.0;  // cf.py:321
    const double close_rp = div_rn(sub_rn(P.r_par_max, P.r_par_min), (double)P.num_bins_r_par);

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_e0 = (long long)atomicAdd(g_counter, (unsigned long long)XI_CHUNK);
            s_ctr = 0;
        }
        __syncthreads();
        const long long e0 = s_e0;
        if (e0 >= pr.n_pairs) break;
        const long long left = pr.n_pairs - e0;
        const int nunits = (int)(left < XI_CHUNK ? left : XI_CHUNK) * tmax;

        for (;;) {
            int u = 0;
            if (lane == 0) u = atomicAdd(&s_ctr, 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= nunits) break;
            const long long e = e0 + u / tmax;
            const int tile = u % tmax;

            const int k = pr.nb_f1[e];
            const int f1 = pr.f1_index[k];
            const int f2 = pr.nb_f2[e];
            const long long a = c1.offset[f1];
            const int n1 = (int)(c1.offset[f1 + 1] - a);
            const int i0 = tile * 32 * R;
            if (i0 >= n1) continue;
            const long long b = c2.offset[f2];
            const int n2 = (int)(c2.offset[f2 + 1] - b);
            if (n2 == 0) continue;
            const double ang = pr.nb_ang[e];
            const double ch = pr.nb_cos[e], sh = pr.nb_sin[e];
            double *__restrict__ orow = out + (size_t)out_row[k] * 6 * nb;

            const bool zerr_on = P.has_zerr_cut && (ang < zerr_ang);
            const bool shp = P.remove_same_half_plate_close_pairs &&
                             pb2_same_half_plate(c1, c2, f1, f2);
            const bool zcut = P.has_z_min_pairs || P.has_z_max_pairs;
            const double zq1 = c1.z_qso[f1], zq2 = c2.z_qso[f2];

            // ---- rows of this tile -> registers
            double rc1[R], dm1[R], z1[R], w1[R], dw1[R];
            bool v1[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = i0 + 32 * r + lane;
                const bool ok = i < n1;
                const long long p = a + (ok ? i : 0);
                rc1[r] = __ldg(c1.r_comov + p);
                dm1[r] = __ldg(c1.dist_m + p);
                z1[r] = __ldg(c1.z + p);
                w1[r] = __ldg(c1.weights + p);
                dw1[r] = __ldg(c1.delta_w + p);
                v1[r] = ok && (w1[r] != 0.);                                 // cf.py:318
                if (zerr_on && v1[r] && pb2_zerr_close(P, z1[r], zq2)) v1[r] = false;  // :321-328
            }

            // ---- column window per row set
            int jl[R], jh[R];
            int JL = 0, JH = n2;
            if (FAST) {
                // |r_par| < r_par_max  (or r_par_min <= r_par < r_par_max when signed) and
                // r_trans < r_trans_max, widened by 1e-9 relative: a superset, never a cut.
                const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
                const double dmax = P.r_par_max * inv_c * (1. + 1e-9) + 1e-9;
                const double dmin = P.x_correlation ? (P.r_par_min * inv_c) : -dmax;
                const double dlow = P.x_correlation ? (dmin - fabs(dmin) * 1e-9 - 1e-9) : dmin;
                const double tmax_sum = P.r_trans_max * inv_s * (1. + 1e-9) + 1e-9;
                JL = n2;
                JH = 0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int ifirst = i0 + 32 * r;
                    if (ifirst >= n1) {
                        jl[r] = n2;
                        jh[r] = 0;
                        continue;
                    }
                    const int ilast = min(ifirst + 31, n1 - 1);
                    const double rc_first = __ldg(c1.r_comov + a + ifirst);
                    const double rc_last = __ldg(c1.r_comov + a + ilast);
                    const double dm_first = __ldg(c1.dist_m + a + ifirst);
                    // need rc2 > rc1 - dmax  and  rc2 < rc1 - dlow  and dm2 < tmax_sum - dm1
                    const int lo = warp_lower_bound(c2.r_comov + b, n2, rc_first - dmax, false, lane);
                    int hi = warp_lower_bound(c2.r_comov + b, n2, rc_last - dlow, true, lane);
                    if (isfinite(tmax_sum)) {
                        const int hi2 = warp_lower_bound(c2.dist_m + b, n2, tmax_sum - dm_first,
                                                         true, lane);
                        hi = min(hi, hi2);
                    }
                    jl[r] = lo;
                    jh[r] = hi;
                    if (hi > lo) {
                        JL = min(JL, lo);
                        JH = max(JH, hi);
                    }
                }
                if (JH <= JL) continue;
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    jl[r] = 0;
                    jh[r] = n2;
                }
            }

            Run live[R], parked[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                run_clear(live[r]);
                run_clear(parked[r]);
            }

            const int nsteps = JH - JL + 31;
            const double *__restrict__ p_rc2 = c2.r_comov + b;
            const double *__restrict__ p_dm2 = c2.dist_m + b;
            const double *__restrict__ p_z2 = c2.z + b;
            const double *__restrict__ p_w2 = c2.weights + b;
            const double *__restrict__ p_dw2 = c2.delta_w + b;

            // software pipeline: columns of step s+1 are loaded while step s is computed
            double n_rc2 = 0., n_dm2 = 0., n_z2 = 0., n_w2 = 0., n_dw2 = 0.;
            {
                const int j = JL - 31 + lane;
                if (j >= 0 && j < n2) {
                    n_rc2 = __ldg(p_rc2 + j);
                    n_dm2 = __ldg(p_dm2 + j);
                    n_z2 = __ldg(p_z2 + j);
                    n_w2 = __ldg(p_w2 + j);
                    n_dw2 = __ldg(p_dw2 + j);
                }
            }
            for (int s = 0; s < nsteps; s++) {
                const int j0 = JL - 31 + s;  // column of lane 0
                const int j = j0 + lane;
                const double rc2 = n_rc2, dm2 = n_dm2, z2 = n_z2, w2 = n_w2, dw2 = n_dw2;
                {
                    const int jn = j + 1;
                    n_w2 = 0.;
                    if (jn >= 0 && jn < n2 && s + 1 < nsteps) {
                        n_rc2 = __ldg(p_rc2 + jn);
                        n_dm2 = __ldg(p_dm2 + jn);
                        n_z2 = __ldg(p_z2 + jn);
                        n_w2 = __ldg(p_w2 + jn);
                        n_dw2 = __ldg(p_dw2 + jn);
                    }
                }
                bool v2 = (j >= 0) && (j < n2) && (w2 != 0.);                  // cf.py:331
                if (zerr_on && v2 && pb2_zerr_close(P, z2, zq1)) v2 = false;    // cf.py:341-348

#pragma unroll
                for (int r = 0; r < R; r++) {
                    // warp-uniform: does any lane of this row set see a column of its window?
                    if (j0 + 31 < jl[r] || j0 >= jh[r]) continue;

                    bool in;
                    int bin;
                    double rp, rt;
                    if (FAST) {
                        rp = mul_rn(sub_rn(rc1[r], rc2), ch);
                        if (!P.x_correlation) rp = fabs(rp);
                        rt = mul_rn(add_rn(dm1[r], dm2), sh);
                        const double x = sub_rn(rp, P.r_par_min);
                        const int bpl = __double2loint(__fma_rd(x, F.kp_lo, PB2_MAGIC));
                        const int bph = __double2loint(__fma_rd(x, F.kp_hi, PB2_MAGIC));
                        const int btl = __double2loint(__fma_rd(rt, F.kt_lo, PB2_MAGIC));
                        const int bth = __double2loint(__fma_rd(rt, F.kt_hi, PB2_MAGIC));
                        const bool both = v1[r] && v2;
                        const bool sure = (bpl == bph) && (btl == bth);
                        in = both && sure && ((unsigned)bpl < np_u) && ((unsigned)btl < nt_u);
                        bin = btl + (int)nt_u * bpl;
                        if (__any_sync(0xffffffffu, both && !sure)) {
                            if (both && !sure) {  // reference expression, true divisions
                                PairGeom g = pb2_pair_exact(P, rc1[r], dm1[r], rc2, dm2, ang, ch,
                                                            sh, false, false);
                                in = g.bin >= 0;
                                bin = g.bin;
                            }
                        }
                    } else {
                        PairGeom g = pb2_pair_exact(P, rc1[r], dm1[r], rc2, dm2, ang, ch, sh,
                                                    false, false);
                        in = v1[r] && v2 && (g.bin >= 0);
                        bin = g.bin;
                        rp = g.r_par;
                        rt = g.r_trans;
                    }
                    const double zz = add_rn(z1[r], z2);
                    if (zcut || shp) {
                        const double zm = div_rn(zz, 2.);
                        if (P.has_z_min_pairs && zm < P.z_min_pairs) in = false;  // cf.py:336
                        if (P.has_z_max_pairs && zm > P.z_max_pairs) in = false;
                        if (shp && fabs(rp) < close_rp) in = false;               // cf.py:378-380
                    }

                    const bool brk = in && (bin != live[r].key);
                    if (__any_sync(0xffffffffu, brk)) {
                        const bool need = brk && live[r].key >= 0 && parked[r].key >= 0;
                        if (__any_sync(0xffffffffu, need)) flush_runs(parked[r], orow, nb, lane);
                        if (brk) {
                            if (live[r].key >= 0) parked[r] = live[r];
                            run_clear(live[r]);
                            live[r].key = bin;
                        }
                    }
                    if (in) {
                        const double w12 = mul_rn(w1[r], w2);
                        live[r].we += w12;
                        live[r].xi = fma(dw1[r], dw2, live[r].xi);
                        live[r].rp = fma(rp, w12, live[r].rp);
                        live[r].rt = fma(rt, w12, live[r].rt);
                        live[r].z = fma(zz, w12, live[r].z);  // halved at flush time
                        live[r].cnt += 1;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                flush_runs(parked[r], orow, nb, lane);
                flush_runs(live[r], orow, nb, lane);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// validation kernel: one warp per forest pair, every pixel pair evaluated with the reference
// expression, six global atomics per binned pair.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pb2_xi_auto_brute(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr,
                  const int32_t *__restrict__ out_row, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const double zerr_ang = mul_rn(P.zerr_cut_deg, PB2_PI) / 180.0;
    for (long long e = warp; e < pr.n_pairs; e += nwarps) {
        const int k = pr.nb_f1[e];
        const int f1 = pr.f1_index[k], f2 = pr.nb_f2[e];
        const long long a = c1.offset[f1], b = c2.offset[f2];
        const long long n1 = c1.offset[f1 + 1] - a, n2 = c2.offset[f2 + 1] - b;
        const double ang = pr.nb_ang[e], ch = pr.nb_cos[e], sh = pr.nb_sin[e];
        const bool zerr_on = P.has_zerr_cut && (ang < zerr_ang);
        const bool shp = P.remove_same_half_plate_close_pairs && pb2_same_half_plate(c1, c2, f1, f2);
        const double zq1 = c1.z_qso[f1], zq2 = c2.z_qso[f2];
        double *orow = out + (size_t)out_row[k] * 6 * nb;
        for (long long idx = lane; idx < n1 * n2; idx += 32) {
            const long long i = idx / n2, j = idx - i * n2;
            const double w1 = c1.weights[a + i], w2 = c2.weights[b + j];
            if (w1 == 0. || w2 == 0.) continue;
            const double z1 = c1.z[a + i], z2 = c2.z[b + j];
            if (zerr_on && pb2_zerr_close(P, z1, zq2)) continue;
            const double z = div_rn(add_rn(z1, z2), 2.);
            if ((P.has_z_min_pairs && z < P.z_min_pairs) || (P.has_z_max_pairs && z > P.z_max_pairs))
                continue;
            if (zerr_on && pb2_zerr_close(P, z2, zq1)) continue;
            PairGeom g = pb2_pair_exact(P, c1.r_comov[a + i], c1.dist_m[a + i], c2.r_comov[b + j],
                                        c2.dist_m[b + j], ang, ch, sh, false, shp);
            if (g.bin < 0) continue;
            const double w12 = mul_rn(w1, w2);
            atomic_add_f64(orow + 0 * (size_t)nb + g.bin, w12);
            atomic_add_f64(orow + 1 * (size_t)nb + g.bin,
                           mul_rn(c1.delta_w[a + i], c2.delta_w[b + j]));
            atomic_add_f64(orow + 2 * (size_t)nb + g.bin, mul_rn(g.r_par, w12));
            atomic_add_f64(orow + 3 * (size_t)nb + g.bin, mul_rn(g.r_trans, w12));
            atomic_add_f64(orow + 4 * (size_t)nb + g.bin, mul_rn(z, w12));
            atomic_add_i64(orow + 5 * (size_t)nb + g.bin, 1);
        }
    }
}

__global__ void pb2_xi_normalise_kernel(long long n_rows, int nb, double *out)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * nb) return;
    const long long row = idx / nb;
    const int bin = (int)(idx - row * nb);
    double *base = out + row * 6 * (long long)nb;
    const double w = base[bin];
    if (w > 0.) {  // cf.py:242-246
        base[1 * (size_t)nb + bin] = div_rn(base[1 * (size_t)nb + bin], w);
        base[2 * (size_t)nb + bin] = div_rn(base[2 * (size_t)nb + bin], w);
        base[3 * (size_t)nb + bin] = div_rn(base[3 * (size_t)nb + bin], w);
        base[4 * (size_t)nb + bin] = div_rn(base[4 * (size_t)nb + bin], w);
    }
}

// ------------------------------------------------------------------------------------------
static unsigned long long *g_counter_dev = nullptr;

static int32_t get_counter(unsigned long long **ptr, cudaStream_t s)
{
    if (!g_counter_dev) PB2_CUDA(cudaMalloc(&g_counter_dev, 64));
    PB2_CUDA(cudaMemsetAsync(g_counter_dev, 0, 64, s));
    *ptr = g_counter_dev;
    return 0;
}

template <int R>
static int32_t launch_tiled(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                            const pb2_pairs *pairs, const int32_t *d_out_row, double *d_out,
                            cudaStream_t s)
{
    XiFast F;
    const double kp = (double)par->num_bins_r_par / (par->r_par_max - par->r_par_min);
    const double kt = (double)par->num_bins_r_trans / par->r_trans_max;
    const double eps = 9.094947017729282e-13;  // 2^-40
    F.kp_lo = kp * (1. - eps);
    F.kp_hi = kp * (1. + eps);
    F.kt_lo = kt * (1. - eps);
    F.kt_hi = kt * (1. + eps);
    F.fast = (!par->rmu_binning && !par->ang_correlation && c1->sorted && c2->sorted &&
              par->num_bins_r_par <= 4096 && par->num_bins_r_trans <= 4096 &&
              par->r_par_max > par->r_par_min && par->r_trans_max > 0.) ? 1 : 0;
    F.tmax = (c1->max_pix + 32 * R - 1) / (32 * R);
    if (F.tmax < 1) F.tmax = 1;
    unsigned long long *ctr = nullptr;
    if (int32_t e = get_counter(&ctr, s)) return e;
    int dev = 0, sms = 0;
    PB2_CUDA(cudaGetDevice(&dev));
    PB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long want = (pairs->n_pairs + XI_CHUNK - 1) / XI_CHUNK;
    int blocks = (int)(want < sms ? want : sms);
    if (blocks < 1) blocks = 1;
    if (F.fast)
        pb2_xi_auto_tiled<R, true><<<blocks, XI_THREADS, 0, s>>>(*c1, *c2, *par, *pairs, F,
                                                                  d_out_row, d_out, ctr);
    else
        pb2_xi_auto_tiled<R, false><<<blocks, XI_THREADS, 0, s>>>(*c1, *c2, *par, *pairs, F,
                                                                   d_out_row, d_out, ctr);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_xi_auto_tiled");
}

extern "C" {

int32_t pb2_xi_auto(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                    const pb2_pairs *pairs, const int32_t *d_out_row, int64_t n_rows,
                    double *d_out, int32_t variant, void *stream)
{
    if (!cat1 || !cat2 || !par || !pairs || !d_out_row || !d_out) {
        pb2_set_error("pb2_xi_auto: null pointer argument");
        return PB2_EINVAL;
    }
    (void)n_rows;
    if (pairs->n_pairs <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    pb2_timing_begin(s);
    int32_t rc = 0;
    if (variant == 1) {
        long long blocks = (pairs->n_pairs + 7) / 8;
        if (blocks > 148 * 16) blocks = 148 * 16;
        pb2_xi_auto_brute<<<(unsigned)blocks, 256, 0, s>>>(*cat1, *cat2, *par, *pairs, d_out_row,
                                                           d_out);
        pb2_count_launch(1);
        rc = pb2_check_launch("pb2_xi_auto_brute");
    } else {
        rc = launch_tiled<2>(cat1, cat2, par, pairs, d_out_row, d_out, s);
    }
    pb2_timing_end(s);
    return rc;
}

int32_t pb2_xi_normalise(int64_t n_rows, int32_t nb, double *d_out, void *stream)
{
    if (n_rows <= 0) return 0;
    const long long total = n_rows * nb;
    pb2_xi_normalise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n_rows, nb, d_out);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_xi_normalise");
}

}  // extern "C"
