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
    double magic;                       // 2^52 + 2^51 (kept out of the immediate field)
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

// ------------------------------------------------------------------------------------------
// product kernel
// ------------------------------------------------------------------------------------------
// A "slot" is a warp-level run: every lane holds private partial sums for ONE warp-uniform bin
// (the key).  Two slots per row set cover the common situations (a bin boundary crossing the
// warp, or the warp straddling two r_trans bins); a third simultaneous bin falls back to direct
// atomics.  Partial sums are kept in factored form (row constants applied at flush time):
//   sw  = sum w2            -> weight  = w1 * sw
//   sdw = sum delta2*w2     -> xi      = (delta1*w1) * sdw
//   srp = sum r_par * w2    -> r_par   = w1 * srp
//   srt = sum r_trans * w2  -> r_trans = w1 * srt
//   szw = sum z2*w2         -> z       = (z1 * weight + w1 * szw) / 2      (cf.py:334, :386)
struct Slot {
    double sw, sdw, srp, srt, szw;
    int cnt;
};

__device__ __forceinline__ void slot_clear(Slot &s)
{
    s.sw = s.sdw = s.srp = s.srt = s.szw = 0.;
    s.cnt = 0;
}

// Reduce one slot over the warp and add it to bin `key` of the output row ([6][nb]).
__device__ __forceinline__ void slot_flush(Slot &s, int key, double w1, double dw1, double z1,
                                           double *__restrict__ orow, int nb, int lane)
{
    const double we = w1 * s.sw;
    const double tot = transpose_reduce5(we, dw1 * s.sdw, w1 * s.srp, w1 * s.srt,
                                         0.5 * (z1 * we + w1 * s.szw), lane);
    const int cnt = __reduce_add_sync(0xffffffffu, s.cnt);
    if (lane < 20 && (lane & 3) == 0) {
        atomic_add_f64(orow + (size_t)(lane >> 2) * nb + key, tot);
    } else if (lane == 20) {
        atomic_add_i64(orow + (size_t)5 * nb + key, (long long)cnt);
    }
    slot_clear(s);
}

#define XI_THREADS 384
#define XI_CHUNK 8

template <int R, bool FAST>
__global__ void __launch_bounds__(XI_THREADS, 1)
pb2_xi_auto_tiled(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, XiFast F,
                  const int32_t *__restrict__ out_row, double *__restrict__ out)
{
    // CTA b owns chunks b, b + G, b + 2G, ... of XI_CHUNK forest pairs; its warps take
    // (pair, row tile) units from a CTA-local counter, so warps of one SM work on the same few
    // forest pairs at a time (L1 reuse of forest 2) without any block-wide barrier.
    __shared__ unsigned s_ctr;
    if (threadIdx.x == 0) s_ctr = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const unsigned np_u = (unsigned)P.num_bins_r_par, nt_u = (unsigned)P.num_bins_r_trans;
    const unsigned tmax = (unsigned)F.tmax;
    const unsigned units_per_chunk = XI_CHUNK * tmax;
    const double zerr_ang = mul_rn(P.zerr_cut_deg, PB2_PI) / 180.0;  // cf.py:321
    const double close_rp = div_rn(sub_rn(P.r_par_max, P.r_par_min), (double)P.num_bins_r_par);
    const bool zcut = P.has_z_min_pairs || P.has_z_max_pairs;
    const double magic = F.magic;

    for (;;) {
        unsigned u = 0;
        if (lane == 0) u = atomicAdd(&s_ctr, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        const long long chunk = (long long)blockIdx.x + (long long)(u / units_per_chunk) * gridDim.x;
        const unsigned local = u % units_per_chunk;
        const long long e = chunk * XI_CHUNK + local / tmax;
        if (chunk * XI_CHUNK >= pr.n_pairs) break;
        if (e >= pr.n_pairs) continue;
        const int tile = (int)(local % tmax);

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
        const bool shp = P.remove_same_half_plate_close_pairs && pb2_same_half_plate(c1, c2, f1, f2);
        const bool extras = zcut || zerr_on || shp;
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
            v1[r] = ok && (w1[r] != 0.);                                           // cf.py:318
            if (zerr_on && v1[r] && pb2_zerr_close(P, z1[r], zq2)) v1[r] = false;  // cf.py:321-328
        }

        // ---- column window per row set (conservative superset; the exact test still runs)
        int jl[R], jh[R];
        int JL = 0, JH = n2;
        if (FAST) {
            const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
            const double dmax = P.r_par_max * inv_c * (1. + 1e-9) + 1e-9;
            const double dmin = P.r_par_min * inv_c;
            const double dlow = P.x_correlation ? (dmin - fabs(dmin) * 1e-9 - 1e-9) : -dmax;
            const double tsum = P.r_trans_max * inv_s * (1. + 1e-9) + 1e-9;
            JL = n2;
            JH = 0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int ifirst = i0 + 32 * r;
                jl[r] = n2;
                jh[r] = 0;
                if (ifirst < n1) {
                    const int ilast = min(ifirst + 31, n1 - 1);
                    const double rc_first = __ldg(c1.r_comov + a + ifirst);
                    const double rc_last = __ldg(c1.r_comov + a + ilast);
                    const double dm_first = __ldg(c1.dist_m + a + ifirst);
                    // rc2 > rc1 - dmax, rc2 < rc1 - dlow, dm2 < tsum - dm1
                    const int lo = warp_lower_bound(c2.r_comov + b, n2, rc_first - dmax, false, lane);
                    int hi = warp_lower_bound(c2.r_comov + b, n2, rc_last - dlow, true, lane);
                    if (isfinite(tsum))
                        hi = min(hi, warp_lower_bound(c2.dist_m + b, n2, tsum - dm_first, true, lane));
                    jl[r] = lo;
                    jh[r] = hi;
                    if (hi > lo) {
                        JL = min(JL, lo);
                        JH = max(JH, hi);
                    }
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

        Slot sa[R], sb[R];
        int ka[R], kb[R];  // warp-uniform keys, -1 = empty
#pragma unroll
        for (int r = 0; r < R; r++) {
            slot_clear(sa[r]);
            slot_clear(sb[r]);
            ka[r] = kb[r] = -1;
        }

        const int nsteps = JH - JL + 31;
        const double *__restrict__ p_rc2 = c2.r_comov + b;
        const double *__restrict__ p_dm2 = c2.dist_m + b;
        const double *__restrict__ p_z2 = c2.z + b;
        const double *__restrict__ p_w2 = c2.weights + b;
        const double *__restrict__ p_dw2 = c2.delta_w + b;
        const double *__restrict__ p_zw2 = c2.z_w + b;

        for (int s = 0; s < nsteps; s++) {
            const int j0 = JL - 31 + s;  // column of lane 0
            const int j = j0 + lane;
            double rc2 = 0., dm2 = 0., w2 = 0., dw2 = 0., zw2 = 0., z2 = 0.;
            const bool jok = (j >= 0) && (j < n2);
            if (jok) {
                rc2 = __ldg(p_rc2 + j);
                dm2 = __ldg(p_dm2 + j);
                w2 = __ldg(p_w2 + j);
                dw2 = __ldg(p_dw2 + j);
                zw2 = __ldg(p_zw2 + j);
                if (extras) z2 = __ldg(p_z2 + j);
            }
            bool v2 = jok && (w2 != 0.);                                  // cf.py:331
            if (zerr_on && v2 && pb2_zerr_close(P, z2, zq1)) v2 = false;  // cf.py:341-348

#pragma unroll
            for (int r = 0; r < R; r++) {
                // warp-uniform: does any lane of this row set see a column of its window?
                if (j0 + 31 < jl[r] || j0 >= jh[r]) continue;

                bool in, unsure = false;
                int bin;
                double rp, rt;
                if (FAST) {
                    rp = mul_rn(sub_rn(rc1[r], rc2), ch);
                    if (!P.x_correlation) rp = fabs(rp);
                    rt = mul_rn(add_rn(dm1[r], dm2), sh);
                    const double x = sub_rn(rp, P.r_par_min);
                    const int bpl = __double2loint(__fma_rd(x, F.kp_lo, magic));
                    const int bph = __double2loint(__fma_rd(x, F.kp_hi, magic));
                    const int btl = __double2loint(__fma_rd(rt, F.kt_lo, magic));
                    const int bth = __double2loint(__fma_rd(rt, F.kt_hi, magic));
                    const bool both = v1[r] && v2;
                    const bool sure = (bpl == bph) && (btl == bth);
                    in = both && sure && ((unsigned)bpl < np_u) && ((unsigned)btl < nt_u);
                    bin = btl + (int)nt_u * bpl;
                    unsure = both && !sure;
                } else {
                    PairGeom g = pb2_pair_exact(P, rc1[r], dm1[r], rc2, dm2, ang, ch, sh, false,
                                                false);
                    in = v1[r] && v2 && (g.bin >= 0);
                    bin = g.bin;
                    rp = g.r_par;
                    rt = g.r_trans;
                }
                if (extras && in) {
                    const double zm = div_rn(add_rn(z1[r], z2), 2.);              // cf.py:334
                    if (P.has_z_min_pairs && zm < P.z_min_pairs) in = false;      // cf.py:336
                    if (P.has_z_max_pairs && zm > P.z_max_pairs) in = false;
                    if (shp && fabs(rp) < close_rp) in = false;                   // cf.py:378-380
                }

                bool is_a = in && (bin == ka[r]);
                bool is_b = in && (bin == kb[r]);
                if (__any_sync(0xffffffffu, unsure || (in && !is_a && !is_b))) {
                    // ---- rare path: resolve borderline bins exactly, then open / recycle slots
                    if (unsure) {
                        PairGeom g = pb2_pair_exact(P, rc1[r], dm1[r], rc2, dm2, ang, ch, sh,
                                                    false, false);
                        in = g.bin >= 0;
                        bin = g.bin;
                        if (extras && in) {
                            const double zm = div_rn(add_rn(z1[r], z2), 2.);
                            if (P.has_z_min_pairs && zm < P.z_min_pairs) in = false;
                            if (P.has_z_max_pairs && zm > P.z_max_pairs) in = false;
                            if (shp && fabs(rp) < close_rp) in = false;
                        }
                    }
                    for (;;) {
                        is_a = in && (bin == ka[r]);
                        is_b = in && (bin == kb[r]);
                        const unsigned other = __ballot_sync(0xffffffffu, in && !is_a && !is_b);
                        if (!other) break;
                        const int key = __shfl_sync(0xffffffffu, bin, __ffs(other) - 1);
                        const unsigned ma = __ballot_sync(0xffffffffu, is_a);
                        const unsigned mb = __ballot_sync(0xffffffffu, is_b);
                        if (ka[r] < 0) {
                            ka[r] = key;
                        } else if (kb[r] < 0) {
                            kb[r] = key;
                        } else if (!ma) {
                            slot_flush(sa[r], ka[r], w1[r], dw1[r], z1[r], orow, nb, lane);
                            ka[r] = key;
                        } else if (!mb) {
                            slot_flush(sb[r], kb[r], w1[r], dw1[r], z1[r], orow, nb, lane);
                            kb[r] = key;
                        } else {
                            // three bins live at once: these lanes add their pair directly
                            if (in && bin == key) {
                                const double w12 = mul_rn(w1[r], w2);
                                atomic_add_f64(orow + 0 * (size_t)nb + bin, w12);
                                atomic_add_f64(orow + 1 * (size_t)nb + bin, mul_rn(dw1[r], dw2));
                                atomic_add_f64(orow + 2 * (size_t)nb + bin, mul_rn(rp, w12));
                                atomic_add_f64(orow + 3 * (size_t)nb + bin, mul_rn(rt, w12));
                                atomic_add_f64(orow + 4 * (size_t)nb + bin,
                                               0.5 * (z1[r] * w12 + w1[r] * zw2));
                                atomic_add_i64(orow + 5 * (size_t)nb + bin, 1);
                                in = false;
                            }
                        }
                    }
                }
                if (__any_sync(0xffffffffu, is_b)) {
                    if (is_b) {
                        sb[r].sw += w2;
                        sb[r].sdw += dw2;
                        sb[r].szw += zw2;
                        sb[r].srp = fma(rp, w2, sb[r].srp);
                        sb[r].srt = fma(rt, w2, sb[r].srt);
                        sb[r].cnt += 1;
                    }
                }
                if (is_a) {
                    sa[r].sw += w2;
                    sa[r].sdw += dw2;
                    sa[r].szw += zw2;
                    sa[r].srp = fma(rp, w2, sa[r].srp);
                    sa[r].srt = fma(rt, w2, sa[r].srt);
                    sa[r].cnt += 1;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (ka[r] >= 0) slot_flush(sa[r], ka[r], w1[r], dw1[r], z1[r], orow, nb, lane);
            if (kb[r] >= 0) slot_flush(sb[r], kb[r], w1[r], dw1[r], z1[r], orow, nb, lane);
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
    F.magic = PB2_MAGIC;
    int dev = 0, sms = 0;
    PB2_CUDA(cudaGetDevice(&dev));
    PB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long want = (pairs->n_pairs + XI_CHUNK - 1) / XI_CHUNK;
    int blocks = (int)(want < sms ? want : sms);
    if (blocks < 1) blocks = 1;
    if (F.fast)
        pb2_xi_auto_tiled<R, true><<<blocks, XI_THREADS, 0, s>>>(*c1, *c2, *par, *pairs, F,
                                                                  d_out_row, d_out);
    else
        pb2_xi_auto_tiled<R, false><<<blocks, XI_THREADS, 0, s>>>(*c1, *c2, *par, *pairs, F,
                                                                   d_out_row, d_out);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_xi_auto_tiled");
}

int32_t pb2_launch_xi_fast(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                           const pb2_pairs *pairs, const int32_t *d_out_row, double *d_out,
                           cudaStream_t s);

bool pb2_xi_diag_eligible(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                          int64_t n_rows);
int32_t pb2_launch_xi_diag(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                           const pb2_pairs *pairs, const int32_t *d_out_row, int64_t n_rows,
                           double *d_out, cudaStream_t s);

// the specialised kernels (pb2_xi_diag.cu, pb2_xi_fast.cu) cover the standard binning without
// per-pair cuts
static bool fast_eligible(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par)
{
    return !par->rmu_binning && !par->ang_correlation && c1->sorted && c2->sorted &&
           !par->has_z_min_pairs && !par->has_z_max_pairs && !par->has_zerr_cut &&
           !par->remove_same_half_plate_close_pairs && par->num_bins_r_par <= 4096 &&
           par->num_bins_r_trans <= 4096 && par->r_par_max > par->r_par_min &&
           par->r_trans_max > 0.;
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
    } else if (variant == 0 && fast_eligible(cat1, cat2, par) &&
               pb2_xi_diag_eligible(cat1, cat2, par, n_rows)) {
        rc = pb2_launch_xi_diag(cat1, cat2, par, pairs, d_out_row, n_rows, d_out, s);
    } else if ((variant == 0 || variant == 3) && fast_eligible(cat1, cat2, par)) {
        rc = pb2_launch_xi_fast(cat1, cat2, par, pairs, d_out_row, d_out, s);
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
