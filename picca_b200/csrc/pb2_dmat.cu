// Distortion matrix: replaces cf.compute_dmat's pair loop + cf.compute_dmat_forest_pairs_fast
// (reference py/picca/cf.py:424-502, 520-887) and the xcf equivalents (py/picca/xcf.py:360-409,
// 427-674).
//
// The reference updates, for every selected pixel pair (i,j) in data bin A and every model bin k
// the forest pair touches,
//     dmat[A,k] += w12 * ( [k == B(i,j)] zf + eta5[k] + eta6[k] dll2_j + eta7[k] dll1_i
//                          + eta8[k] dll1_i dll2_j - eta1[i,k] - eta2[j,k]
//                          - eta3[i,k] dll2_j - eta4[j,k] dll1_i )                 (cf.py:851-887)
// i.e. N_selected x U read-modify-writes into a 50 MB matrix per forest pair.  Summing over the
// pairs of a data bin first turns this into one small dense contraction per forest pair
// (SURVEY.md Appendix B):
//     dmat[A,k] += sum_r X[r,A] * Y[r,k]       r over {pixels of forest 1 (x2), pixels of forest 2
//                                                (x2), four rank-1 terms}
//     X = (-Q1, -Q1d, -Q2, -Q2d, P0, P2, P1, P12),   Y = (eta1, eta3, eta2, eta4, eta5..eta8)
// with Q1[A,i] = w1_i sum_{j in S,A} w2_j etc.  One CTA per kept forest pair:
//   pass 0  exact bins of every in-range pixel pair -> sets of touched model / data bins, compact
//           indices (U, UA), early exit when nothing is in range (cf.py:547-571);
//   sweep 1 one thread per pixel of forest 1 walks its row: run-length sums per (A,B) segment give
//           eta1/eta3 and Q1/Q1d rows (no atomics: the thread owns the row) and, per segment,
//           native fp64 reductions for the diagonal term, weights_dmat, the effective r_par /
//           r_trans / z / weight and the rank-1 factors;
//   sweep 2 one thread per pixel of forest 2 walks its column: eta2/eta4 and Q2/Q2d rows;
//   GEMM    64x64 register-tiled fp64 contraction over r, epilogue = red.global.add.f64 into dmat.
// All bins use the reference expression with true IEEE divisions (no fast path here: the sweeps
// are <5 % of the work).  The sums are re-associated with respect to the reference, which stays
// within 1e-9 (measured 2e-11 in the survey) but is not bit-exact; counts of pairs are exact.
#include <stdlib.h>

#include "pb2_dmat.cuh"

__global__ void dmat_compact_kernel(pb2_pairs pr, DmatWork W)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= pr.n_pairs) return;
    if (pr.nb_keep == nullptr || pr.nb_keep[e]) {
        const unsigned long long slot = atomicAdd(W.count, 1ull);
        W.kept[slot] = e;
    }
}

// ------------------------------------------------------------------------------------------
// auto / delta x delta
// ------------------------------------------------------------------------------------------
// per-forest constants of the projection (cf.py:577-594) and the per-pixel redshift-evolution
// factor (cf.py:680-685), once per call instead of once per forest pair.  One warp per forest.
__global__ void dmat_prologue_kernel(pb2_catalog c, pb2_params P, double alpha, double *__restrict__ fz,
                                     double *__restrict__ dl, double2 *__restrict__ fs)
{
    const int lane = threadIdx.x & 31;
    const long long f = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= c.n_los) return;
    const long long a = c.offset[f];
    const int n = (int)(c.offset[f + 1] - a);
    const double *w = c.weights + a, *ll = c.log_lambda + a, *z = c.z + a;
    double t = 0., u = 0.;
    for (int i = lane; i < n; i += 32) {
        t += w[i];
        u += ll[i] * w[i];
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        t += __shfl_xor_sync(0xffffffffu, t, m);
        u += __shfl_xor_sync(0xffffffffu, u, m);
    }
    const double mll = u / t;
    double q = 0.;
    for (int i = lane; i < n; i += 32) {
        const double d = ll[i] - mll;
        dl[a + i] = d;
        q += w[i] * (d * d);
        fz[a + i] = P.redshift_evolution_in_distortion_matrix
                        ? pow((1. + z[i]) / (1. + P.z_ref), alpha - 1.) : 1.;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) q += __shfl_xor_sync(0xffffffffu, q, m);
    if (lane == 0) fs[f] = make_double2(t, q);
}

__global__ void __launch_bounds__(DM_THREADS, 2)
pb2_dmat_auto_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, DmatWork W,
                     double *__restrict__ weights_dmat, double *__restrict__ dmat,
                     double *__restrict__ r_par_eff, double *__restrict__ r_trans_eff,
                     double *__restrict__ z_eff, double *__restrict__ weight_eff)
{
    __shared__ double red[DM_THREADS / 32];
    __shared__ long long s_e;
    __shared__ int s_cnt[2];   // in-range pairs that are not close, in-range pairs
    __shared__ int s_U, s_UA;
    __shared__ int s_cxlo[DM_MAXCH], s_cxhi[DM_MAXCH], s_cylo[DM_MAXCH], s_cyhi[DM_MAXCH];
    __shared__ int s_act[DM_MAXCH];
    __shared__ int s_nact;
    __shared__ __align__(16) double Xs[DM_KCH][DM_TILE];
    __shared__ __align__(16) double Ys[DM_KCH][DM_TILE];

    const int tid = threadIdx.x;
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const int nbm = P.num_model_bins_r_par * P.num_model_bins_r_trans;
    const double zerr_ang = mul_rn(P.zerr_cut_deg, PB2_PI) / 180.0;
    const bool windows = !P.rmu_binning && c1.sorted && c2.sorted;

    // per-CTA scratch carve-up
    char *base = W.cta_base + (long long)blockIdx.x * W.cta_stride;
    int *kidx = (int *)base;                     // [nbm] compact model index, -1 = untouched
    int *aidx = kidx + nbm;                      // [nb]
    int *klist = aidx + nb;                      // [nbm]
    int *alist = klist + nbm;                    // [nb]
    double *X = (double *)(((uintptr_t)(alist + nb) + 15) & ~(uintptr_t)15);  // [rows_max][cap]
    double *Y = X + (long long)W.rows_max * W.cap;
    // per pixel (forest 1 then forest 2): first / last compact column written in X and in Y
    int *xlo = (int *)(Y + (long long)W.rows_max * W.cap);
    int *xhi = xlo + (c1.max_pix + c2.max_pix);
    int *ylo = xhi + (c1.max_pix + c2.max_pix);
    int *yhi = ylo + (c1.max_pix + c2.max_pix);

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const unsigned long long t = atomicAdd(W.count + 1, 1ull);
            s_e = (t < W.count[0]) ? W.kept[t] : -1;
        }
        __syncthreads();
        const long long e = s_e;
        if (e < 0) break;

        const int k1 = pr.nb_f1[e];
        const int f1 = pr.f1_index[k1], f2 = pr.nb_f2[e];
        const long long a = c1.offset[f1], b = c2.offset[f2];
        const int n1 = (int)(c1.offset[f1 + 1] - a), n2 = (int)(c2.offset[f2 + 1] - b);
        const double ang = pr.nb_ang[e], ch = pr.nb_cos[e], sh = pr.nb_sin[e];
        const bool zerr_on = P.has_zerr_cut && (ang < zerr_ang);
        const bool shp = P.remove_same_half_plate_close_pairs && pb2_same_half_plate(c1, c2, f1, f2);
        const double zq1 = c1.z_qso[f1], zq2 = c2.z_qso[f2];
        const int order1 = c1.order[f1], order2 = c2.order[f2];
        const double *rc1 = c1.r_comov + a, *dm1 = c1.dist_m + a, *z1 = c1.z + a;
        const double *w1 = c1.weights + a, *ll1 = c1.log_lambda + a;
        const double *rc2 = c2.r_comov + b, *dm2 = c2.dist_m + b, *z2 = c2.z + b;
        const double *w2 = c2.weights + b, *ll2 = c2.log_lambda + b;

        // ---------------- pass 0: touched bins (cf.py:547-571 and the bins of pass 1)
        for (int x = tid; x < nbm; x += DM_THREADS) kidx[x] = -1;
        for (int x = tid; x < nb; x += DM_THREADS) aidx[x] = -1;
        if (tid == 0) s_cnt[0] = s_cnt[1] = 0;
        __syncthreads();
        {
            int cnt_nc = 0, cnt_in = 0;
            // (every pair loop of this kernel runs a warp-uniform number of iterations with a
            // __syncwarp() on top: without it the lanes of a warp, whose rows have different
            // windows and flush points, end up executing the loop one or two lanes at a time)
            for (int ib = 0; ib < n1; ib += DM_THREADS) {
                const int i = min(ib + tid, n1 - 1);
                const bool rowok = (ib + tid < n1) && (w1[i] != 0.);
                bool i_sel = true;
                if (zerr_on && pb2_zerr_close(P, z1[i], zq2)) i_sel = false;
                int lo = 0, hi = 0;
                if (rowok)
                    row_window(P, windows, rc1[i], dm1[i], rc2, dm2, n2, ch, sh, P.x_correlation, lo, hi);
                const int len = hi - lo, maxlen = warp_max(len);
                for (int t = 0; t < maxlen; t++) {
                    __syncwarp();
                    if (t >= len) continue;
                    const int j = lo + t;
                    if (w2[j] == 0.) continue;
                    DmatGeom g = dmat_pair(P, W.fast, rc1[i], dm1[i], rc2[j], dm2[j], ch, sh, false, shp);
                    if (!g.in) continue;
                    cnt_in++;
                    if (!g.close) cnt_nc++;
                    kidx[g.B] = 0;
                    bool sel = i_sel && !g.close;
                    if (sel) {
                        const double z = div_rn(add_rn(z1[i], z2[j]), 2.);
                        if ((P.has_z_min_pairs && z < P.z_min_pairs) ||
                            (P.has_z_max_pairs && z > P.z_max_pairs)) sel = false;
                        if (sel && zerr_on && pb2_zerr_close(P, z2[j], zq1)) sel = false;
                    }
                    if (sel) aidx[g.A] = 0;
                }
            }
            if (cnt_nc) atomicAdd(&s_cnt[0], cnt_nc);
            if (cnt_in) atomicAdd(&s_cnt[1], cnt_in);
        }
        __syncthreads();
        if (s_cnt[0] == 0) continue;  // cf.py:570-571

        // compact indices of the touched bins (the set of np.unique, cf.py:846-848), ordered
        // r_trans-major: a pixel row meets every r_par bin but only two or three r_trans bins, so
        // its non-zeros in X and Y then sit in a few contiguous column ranges and the contraction
        // can skip the tiles a K chunk of rows does not touch
        if (tid == 0) {
            int u = 0;
            for (int t = 0; t < P.num_model_bins_r_trans; t++)
                for (int q = 0; q < P.num_model_bins_r_par; q++) {
                    const int x = t + P.num_model_bins_r_trans * q;
                    if (kidx[x] == 0) { kidx[x] = u; klist[u++] = x; }
                }
            s_U = u;
            if (W.stats) {
                atomicAdd(W.stats, (double)s_cnt[0] * (15. * u + 4.) + 40. * (double)s_cnt[1]);
                atomicAdd(W.stats + 1, (double)u);
                atomicAdd(W.stats + 2, (double)s_cnt[1]);
            }
            u = 0;
            for (int t = 0; t < P.num_bins_r_trans; t++)
                for (int q = 0; q < P.num_bins_r_par; q++) {
                    const int x = t + P.num_bins_r_trans * q;
                    if (aidx[x] == 0) { aidx[x] = u; alist[u++] = x; }
                }
            s_UA = u;
        }
        __syncthreads();
        const int U = s_U, UA = s_UA;

        // ---------------- per-forest constants (cf.py:577-594), from the prologue
        const double sw1 = W.fs1[f1].x, swsll1 = W.fs1[f1].y;
        const double sw2 = W.fs2[f2].x, swsll2 = W.fs2[f2].y;
        const double *__restrict__ f1z = W.fz1 + a, *__restrict__ dl1 = W.dl1 + a;
        const double *__restrict__ f2z = W.fz2 + b, *__restrict__ dl2 = W.dl2 + b;

        const int rows = 2 * n1 + 2 * n2 + 4;
        for (int kc = 0; kc < U; kc += W.cap) {
            const int Uc = min(W.cap, U - kc);
            const int Upad = (Uc + DM_TILE - 1) / DM_TILE * DM_TILE;
            for (int ac = 0; ac < max(UA, 1); ac += W.cap) {
                const int UAc = min(W.cap, UA - ac);
                if (UAc <= 0) break;
                const int UApad = (UAc + DM_TILE - 1) / DM_TILE * DM_TILE;
                const bool first = (kc == 0 && ac == 0);
                __syncthreads();
                for (long long x = tid; x < (long long)rows * UApad; x += DM_THREADS) X[x] = 0.;
                for (long long x = tid; x < (long long)rows * Upad; x += DM_THREADS) Y[x] = 0.;
                __syncthreads();
                double *Xp = X + (long long)(2 * n1 + 2 * n2) * UApad;  // P0, P2, P1, P12
                double *Yp = Y + (long long)(2 * n1 + 2 * n2) * Upad;   // eta5, eta6, eta7, eta8

                // (the X / Y updates below are fire-and-forget reductions rather than load-add-store:
                // a row is owned by one thread, but the round trip to the scratch in L2 would stall it)
                // ---------------- sweep 1: rows of forest 1
                for (int ib = 0; ib < n1; ib += DM_THREADS) {
                    const int i = min(ib + tid, n1 - 1);
                    const bool rowok = (ib + tid < n1) && (w1[i] != 0.);
                    bool i_sel = true;
                    if (zerr_on && pb2_zerr_close(P, z1[i], zq2)) i_sel = false;
                    int lo = 0, hi = -1;
                    if (rowok)
                        row_window(P, windows, rc1[i], dm1[i], rc2, dm2, n2, ch, sh, P.x_correlation,
                                   lo, hi);
                    const double wi = w1[i], dli = dl1[i], fzi = f1z[i], zi = z1[i];
                    int rxl = 0x7fffffff, rxh = -1, ryl = 0x7fffffff, ryh = -1;
                    int cA = -1, cB = -1;
                    bool cS = false;
                    double e1 = 0., e3 = 0., q1 = 0., q1d = 0., dg = 0., srp = 0., srt = 0., sz = 0.;
                    double e5 = 0., e6 = 0., e7 = 0., e8 = 0.;
                    const int len = hi - lo + 1, maxlen = warp_max(len);
                    for (int t = 0; t < maxlen; t++) {
                        __syncwarp();
                        if (t >= len) continue;
                        const int j = lo + t;
                        DmatGeom g;
                        g.in = false;
                        bool sel = false;
                        double z = 0.;
                        if (j < hi && w2[j] != 0.) {
                            g = dmat_pair(P, W.fast, rc1[i], dm1[i], rc2[j], dm2[j], ch, sh, false, shp);
                            if (g.in) {
                                z = div_rn(add_rn(zi, z2[j]), 2.);
                                sel = i_sel && !g.close;
                                if (sel && ((P.has_z_min_pairs && z < P.z_min_pairs) ||
                                            (P.has_z_max_pairs && z > P.z_max_pairs))) sel = false;
                                if (sel && zerr_on && pb2_zerr_close(P, z2[j], zq1)) sel = false;
                            }
                        }
                        const bool brk = (j == hi) || (g.in && (g.A != cA || g.B != cB || sel != cS));
                        if (brk && cB >= 0) {  // flush the finished segment
                            const int kb = kidx[cB] - kc;
                            if (kb >= 0 && kb < Uc) {
                                ryl = min(ryl, kb);
                                ryh = max(ryh, kb);
                                atomic_add_f64(&Y[(long long)i * Upad + kb], e1 / sw2);                  // eta1
                                if (order2 == 1) atomic_add_f64(&Y[(long long)(n1 + i) * Upad + kb], e3 / swsll2);
                                atomic_add_f64(Yp + 0 * (long long)Upad + kb, e5 / sw1 / sw2);
                                if (order2 == 1) atomic_add_f64(Yp + 1 * (long long)Upad + kb, e6);
                                if (order1 == 1) atomic_add_f64(Yp + 2 * (long long)Upad + kb, e7);
                                if (order1 == 1 && order2 == 1)
                                    atomic_add_f64(Yp + 3 * (long long)Upad + kb, e8);
                            }
                            if (cS) {
                                const int ka = aidx[cA] - ac;
                                if (ka >= 0 && ka < UAc) {
                                    rxl = min(rxl, ka);
                                    rxh = max(rxh, ka);
                                    atomic_add_f64(&X[(long long)i * UApad + ka], -(wi * q1));
                                    atomic_add_f64(&X[(long long)(n1 + i) * UApad + ka], -(wi * q1d));
                                    atomic_add_f64(Xp + 0 * (long long)UApad + ka, wi * q1);
                                    atomic_add_f64(Xp + 1 * (long long)UApad + ka, wi * q1d);
                                    atomic_add_f64(Xp + 2 * (long long)UApad + ka, wi * dli * q1);
                                    atomic_add_f64(Xp + 3 * (long long)UApad + ka, wi * dli * q1d);
                                }
                                if (first) {
                                    atomic_add_f64(dmat + (long long)cA * nbm + cB, dg);   // cf.py:873
                                    atomic_add_f64(weights_dmat + cA, wi * q1);            // cf.py:718
                                    atomic_add_f64(r_par_eff + cB, srp);                   // cf.py:714
                                    atomic_add_f64(r_trans_eff + cB, srt);
                                    atomic_add_f64(z_eff + cB, sz);
                                    atomic_add_f64(weight_eff + cB, wi * q1);
                                }
                            }
                            e1 = e3 = q1 = q1d = dg = srp = srt = sz = e5 = e6 = e7 = e8 = 0.;
                        }
                        if (j == hi || !g.in) continue;
                        cA = g.A;
                        cB = g.B;
                        cS = sel;
                        const double wj = w2[j], dlj = dl2[j];
                        const double zf = mul_rn(fzi, f2z[j]);
                        const double w12 = mul_rn(wi, wj);
                        e1 += zf * wj;                                   // cf.py:767
                        e3 += zf * wj * dlj;                             // cf.py:782-787
                        e5 += zf * w12;                                  // cf.py:775
                        e6 += zf * wi / sw1 * (wj * dlj / swsll2);       // cf.py:793-802
                        e7 += zf * wj / sw2 * (wi * dli / swsll1);       // cf.py:818-827
                        e8 += zf * wi * dli * wj * dlj / swsll1 / swsll2;  // cf.py:835-843
                        if (sel) {
                            q1 += wj;
                            q1d += wj * dlj;
                            dg += w12 * zf;
                            srp += w12 * g.rp;
                            srt += w12 * g.rt;
                            sz += w12 * z;
                        }
                    }
                    if (ib + tid < n1) {
                        xlo[i] = rxl;
                        xhi[i] = rxh;
                        ylo[i] = ryl;
                        yhi[i] = ryh;
                    }
                }

                // ---------------- sweep 2: columns (pixels of forest 2)
                for (int jb = 0; jb < n2; jb += DM_THREADS) {
                    const int j = min(jb + tid, n2 - 1);
                    const bool colok = (jb + tid < n2) && (w2[j] != 0.);
                    bool j_sel0 = true;
                    if (zerr_on && pb2_zerr_close(P, z2[j], zq1)) j_sel0 = false;
                    const double wj = w2[j], dlj = dl2[j], fzj = f2z[j], zj = z2[j];
                    int cA = -1, cB = -1;
                    bool cS = false;
                    int rxl = 0x7fffffff, rxh = -1, ryl = 0x7fffffff, ryh = -1;
                    double e2 = 0., e4 = 0., q2 = 0., q2d = 0.;
                    int ilo = 0, ihi = -1;
                    if (colok)
                        col_window(P, windows, rc2[j], dm2[j], rc1, dm1, n1, ch, sh, P.x_correlation,
                                   ilo, ihi);
                    const int len = ihi - ilo + 1, maxlen = warp_max(len);
                    for (int t = 0; t < maxlen; t++) {
                        __syncwarp();
                        if (t >= len) continue;
                        const int i = ilo + t;
                        DmatGeom g;
                        g.in = false;
                        bool sel = false;
                        if (i < ihi && w1[i] != 0.) {
                            g = dmat_pair(P, W.fast, rc1[i], dm1[i], rc2[j], dm2[j], ch, sh, false, shp);
                            if (g.in) {
                                sel = j_sel0 && !g.close;
                                if (sel && zerr_on && pb2_zerr_close(P, z1[i], zq2)) sel = false;
                                if (sel) {
                                    const double z = div_rn(add_rn(z1[i], zj), 2.);
                                    if ((P.has_z_min_pairs && z < P.z_min_pairs) ||
                                        (P.has_z_max_pairs && z > P.z_max_pairs)) sel = false;
                                }
                            }
                        }
                        const bool brk = (i == ihi) || (g.in && (g.A != cA || g.B != cB || sel != cS));
                        if (brk && cB >= 0) {
                            const int kb = kidx[cB] - kc;
                            if (kb >= 0 && kb < Uc) {
                                ryl = min(ryl, kb);
                                ryh = max(ryh, kb);
                                atomic_add_f64(&Y[(long long)(2 * n1 + j) * Upad + kb], e2 / sw1);         // eta2
                                if (order1 == 1)
                                    atomic_add_f64(&Y[(long long)(2 * n1 + n2 + j) * Upad + kb], e4 / swsll1);
                            }
                            if (cS) {
                                const int ka = aidx[cA] - ac;
                                if (ka >= 0 && ka < UAc) {
                                    rxl = min(rxl, ka);
                                    rxh = max(rxh, ka);
                                    atomic_add_f64(&X[(long long)(2 * n1 + j) * UApad + ka], -(wj * q2));
                                    atomic_add_f64(&X[(long long)(2 * n1 + n2 + j) * UApad + ka], -(wj * q2d));
                                }
                            }
                            e2 = e4 = q2 = q2d = 0.;
                        }
                        if (i == ihi || !g.in) continue;
                        cA = g.A;
                        cB = g.B;
                        cS = sel;
                        const double zf = mul_rn(f1z[i], fzj);
                        e2 += zf * w1[i];                  // cf.py:771
                        e4 += zf * w1[i] * dl1[i];         // cf.py:808-813
                        if (sel) {
                            q2 += w1[i];
                            q2d += w1[i] * dl1[i];
                        }
                    }
                    if (jb + tid < n2) {
                        xlo[n1 + j] = rxl;
                        xhi[n1 + j] = rxh;
                        ylo[n1 + j] = ryl;
                        yhi[n1 + j] = ryh;
                    }
                }
                __syncthreads();
                __threadfence_block();

                // ---------------- column ranges of every K chunk of rows.  Rows: [0, n1) and
                // [n1, 2 n1) belong to the pixels of forest 1, the next 2 n2 to forest 2, the last
                // four (P0, P2, P1, P12 / eta5..8) are dense.
                const int nch = (rows + DM_KCH - 1) / DM_KCH;
                const bool skipping = nch <= DM_MAXCH;
                if (skipping) {
                    for (int cch = tid; cch < nch; cch += DM_THREADS) {
                        int cxl = 0x7fffffff, cxh = -1, cyl = 0x7fffffff, cyh = -1;
                        for (int r = cch * DM_KCH; r < min(rows, (cch + 1) * DM_KCH); r++) {
                            if (r >= 2 * n1 + 2 * n2) {
                                cxl = cyl = 0;
                                cxh = cyh = 0x7ffffff0;
                                break;
                            }
                            const int px = r < n1 ? r : r < 2 * n1 ? r - n1
                                           : r < 2 * n1 + n2 ? r - n1 : r - n1 - n2;
                            const bool used = px < n1 ? (w1[px] != 0.) : (w2[px - n1] != 0.);
                            if (!used) continue;  // rows of zero-weight pixels were never written
                            cxl = min(cxl, xlo[px]);
                            cxh = max(cxh, xhi[px]);
                            cyl = min(cyl, ylo[px]);
                            cyh = max(cyh, yhi[px]);
                        }
                        s_cxlo[cch] = cxl;
                        s_cxhi[cch] = cxh;
                        s_cylo[cch] = cyl;
                        s_cyhi[cch] = cyh;
                    }
                }
                __syncthreads();

                // ---------------- contraction  C[a,k] = sum_r X[r,a] Y[r,k]
                const int ty = tid >> 4, tx = tid & 15;
                for (int a0 = 0; a0 < UApad; a0 += DM_TILE) {
                    for (int k0 = 0; k0 < Upad; k0 += DM_TILE) {
                        // K chunks whose rows touch both column tiles, in ascending order
                        __syncthreads();
                        if (tid < 32) {
                            int cnt = 0;
                            for (int cb = 0; cb < nch; cb += 32) {
                                const int cch = cb + tid;
                                const bool on = cch < nch &&
                                    (!skipping || (s_cxlo[cch] < a0 + DM_TILE && s_cxhi[cch] >= a0 &&
                                                   s_cylo[cch] < k0 + DM_TILE && s_cyhi[cch] >= k0));
                                const unsigned m = __ballot_sync(0xffffffffu, on);
                                if (on && cnt + __popc(m & ((1u << tid) - 1u)) < DM_MAXCH)
                                    s_act[cnt + __popc(m & ((1u << tid) - 1u))] = cch;
                                cnt += __popc(m);
                            }
                            if (tid == 0) s_nact = cnt;
                        }
                        __syncthreads();
                        const int nact = s_nact;
                        if (nact == 0) continue;
                        double c[4][4];
#pragma unroll
                        for (int p = 0; p < 4; p++)
#pragma unroll
                            for (int q = 0; q < 4; q++) c[p][q] = 0.;
                        // software pipeline: the next K chunk of X and Y is fetched into
                        // registers while the current one is multiplied out of shared memory
                        constexpr int PER = DM_KCH * DM_TILE / DM_THREADS;
                        double px[PER], py[PER];
                        auto fetch = [&](int r0) {
#pragma unroll
                            for (int m = 0; m < PER; m++) {
                                const int x = tid + m * DM_THREADS;
                                const int r = r0 + x / DM_TILE, cc = x % DM_TILE;
                                px[m] = (r < rows) ? X[(long long)r * UApad + a0 + cc] : 0.;
                                py[m] = (r < rows) ? Y[(long long)r * Upad + k0 + cc] : 0.;
                            }
                        };
                        // (without skipping the list may be cut at DM_MAXCH: then walk all chunks)
                        const int nloop = skipping ? nact : nch;
                        fetch((skipping ? s_act[0] : 0) * DM_KCH);
                        for (int n = 0; n < nloop; n++) {
                            __syncthreads();
#pragma unroll
                            for (int m = 0; m < PER; m++) {
                                const int x = tid + m * DM_THREADS;
                                Xs[x / DM_TILE][x % DM_TILE] = px[m];
                                Ys[x / DM_TILE][x % DM_TILE] = py[m];
                            }
                            __syncthreads();
                            if (n + 1 < nloop) fetch((skipping ? s_act[n + 1] : n + 1) * DM_KCH);
#pragma unroll
                            for (int rr = 0; rr < DM_KCH; rr++) {
                                const double4 xa = *reinterpret_cast<const double4 *>(&Xs[rr][ty * 4]);
                                const double4 yk = *reinterpret_cast<const double4 *>(&Ys[rr][tx * 4]);
                                const double xv[4] = {xa.x, xa.y, xa.z, xa.w};
                                const double yv[4] = {yk.x, yk.y, yk.z, yk.w};
#pragma unroll
                                for (int p = 0; p < 4; p++)
#pragma unroll
                                    for (int q = 0; q < 4; q++) c[p][q] = fma(xv[p], yv[q], c[p][q]);
                            }
                        }
#pragma unroll
                        for (int p = 0; p < 4; p++) {
                            const int ai = a0 + ty * 4 + p;
                            if (ai >= UAc) continue;
                            const long long rowp = (long long)alist[ac + ai] * nbm;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                const int ki = k0 + tx * 4 + q;
                                if (ki < Uc && c[p][q] != 0.)
                                    atomic_add_f64(dmat + rowp + klist[kc + ki], c[p][q]);
                            }
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// forest x object: one warp per kept (forest, object) pair
// ------------------------------------------------------------------------------------------

struct XSeg {
    int A, B;
    double e2, e4, q2, q2d;
};

__global__ void __launch_bounds__(128)
pb2_dmat_cross_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, DmatWork W,
                      double *__restrict__ weights_dmat, double *__restrict__ dmat,
                      double *__restrict__ r_par_eff, double *__restrict__ r_trans_eff,
                      double *__restrict__ z_eff, double *__restrict__ weight_eff)
{
    // segment lists live in the per-CTA scratch: [4 warps][seg_cap], seg_cap = max_pix1 + 1
    // (a forest of n pixels has at most n segments)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int seg_cap = W.rows_max;
    XSeg *segs = (XSeg *)(W.cta_base + (long long)blockIdx.x * W.cta_stride) + (long long)wid * seg_cap;
    const int nbm = P.num_model_bins_r_par * P.num_model_bins_r_trans;

    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(W.count + 1, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= W.count[0]) break;
        const long long e = W.kept[t];
        const int k1 = pr.nb_f1[e];
        const int f1 = pr.f1_index[k1], f2 = pr.nb_f2[e];
        const long long a = c1.offset[f1];
        const int n1 = (int)(c1.offset[f1 + 1] - a);
        const long long q = c2.offset[f2];
        const double rcq = c2.r_comov[q], dmq = c2.dist_m[q], zq = c2.z[q], wq = c2.weights[q];
        const double ch = pr.nb_cos[e], sh = pr.nb_sin[e];
        const int order1 = c1.order[f1];
        const double *rc1 = c1.r_comov + a, *dm1 = c1.dist_m + a, *z1 = c1.z + a;
        const double *w1 = c1.weights + a, *ll1 = c1.log_lambda + a;
        if (wq == 0. || n1 == 0) continue;  // xcf.py:516

        // forest constants (xcf.py:475-486), warp reductions
        double s = 0., sl = 0.;
        for (int i = lane; i < n1; i += 32) {
            s += w1[i];
            sl += ll1[i] * w1[i];
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, m);
            sl += __shfl_xor_sync(0xffffffffu, sl, m);
        }
        const double sw1 = s, mll1 = sl / s;
        double sq = 0.;
        for (int i = lane; i < n1; i += 32) {
            const double d = ll1[i] - mll1;
            sq += w1[i] * (d * d);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, m);
        const double swsll1 = sq;
        const double fzq = P.redshift_evolution_in_distortion_matrix
                               ? pow((1. + zq) / (1. + P.z_ref), P.alpha2 - 1.) : 1.;  // xcf.py:544-549

        // lane 0 walks the pixels (bins change every few pixels; the cost is in the outer
        // product below, which all lanes share)
        int nseg = 0;
        if (lane == 0) {
            int cA = -1, cB = -1;
            bool cS = false;
            double e2 = 0., e4 = 0., q2 = 0., q2d = 0., dg = 0., srp = 0., srt = 0., sz = 0.;
            for (int i = 0; i <= n1; i++) {
                DmatGeom g;
                g.in = false;
                bool sel = false;
                double z = 0.;
                if (i < n1 && w1[i] != 0.) {
                    g = dmat_pair(P, W.fast, rc1[i], dm1[i], rcq, dmq, ch, sh, true, false);
                    if (g.in) {
                        z = div_rn(add_rn(z1[i], zq), 2.);
                        sel = !((P.has_z_min_pairs && z < P.z_min_pairs) ||
                                (P.has_z_max_pairs && z > P.z_max_pairs));  // xcf.py:523-526
                    }
                }
                const bool brk = (i == n1) || (g.in && (g.A != cA || g.B != cB || sel != cS));
                if (brk && cB >= 0) {
                    if (nseg < seg_cap) {
                        XSeg sgm;
                        sgm.A = cS ? cA : -1;
                        sgm.B = cB;
                        sgm.e2 = e2 / sw1;
                        sgm.e4 = (order1 == 1) ? e4 / swsll1 : 0.;
                        sgm.q2 = wq * q2;
                        sgm.q2d = wq * q2d;
                        segs[nseg++] = sgm;
                    }
                    if (cS) {
                        atomic_add_f64(dmat + (long long)cA * nbm + cB, dg);  // xcf.py:666
                        atomic_add_f64(weights_dmat + cA, wq * q2);
                        atomic_add_f64(r_par_eff + cB, srp);
                        atomic_add_f64(r_trans_eff + cB, srt);
                        atomic_add_f64(z_eff + cB, sz);
                        atomic_add_f64(weight_eff + cB, wq * q2);
                    }
                    e2 = e4 = q2 = q2d = dg = srp = srt = sz = 0.;
                }
                if (i == n1 || !g.in) continue;
                cA = g.A;
                cB = g.B;
                cS = sel;
                const double f1 = P.redshift_evolution_in_distortion_matrix
                                      ? pow((1. + z1[i]) / (1. + P.z_ref), P.alpha - 1.) : 1.;
                const double zf = mul_rn(f1, fzq);
                const double dli = ll1[i] - mll1;
                e2 += zf * w1[i];            // xcf.py:625
                e4 += zf * w1[i] * dli;      // xcf.py:632-636
                if (sel) {
                    const double w12 = mul_rn(w1[i], wq);
                    q2 += w1[i];
                    q2d += w1[i] * dli;
                    dg += zf * w12;
                    srp += w12 * g.rp;
                    srt += w12 * g.rt;
                    sz += w12 * z;
                }
            }
        }
        nseg = __shfl_sync(0xffffffffu, nseg, 0);
        __syncwarp();
        // dmat[A_s, B_t] -= q2_s * e2_t + q2d_s * e4_t   for every selected segment s, segment t
        const int total = nseg * nseg;
        for (int x = lane; x < total; x += 32) {
            const int sidx = x / nseg, tidx = x - sidx * nseg;
            const XSeg sa = segs[sidx];
            if (sa.A < 0) continue;
            const XSeg tb = segs[tidx];
            const double v = -(sa.q2 * tb.e2 + sa.q2d * tb.e4);
            if (v != 0.) atomic_add_f64(dmat + (long long)sa.A * nbm + tb.B, v);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
static long long auto_cta_bytes(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par)
{
    const long long nb = (long long)par->num_bins_r_par * par->num_bins_r_trans;
    const long long nbm = (long long)par->num_model_bins_r_par * par->num_model_bins_r_trans;
    const long long rows = 2ll * c1->max_pix + 2ll * c2->max_pix + 4;
    long long bytes = (2 * nb + 2 * nbm) * 4 + 64;
    bytes += (2ll * c1->max_pix + 2ll * c2->max_pix) * 8;
    bytes += 2 * rows * DM_CAP * 8;
    bytes += 4ll * (c1->max_pix + c2->max_pix) * 4;
    return (bytes + 255) / 256 * 256;
}

// per-pixel (fz, dl) and per-forest (sum w, sum w dll^2) constants of both catalogues
static long long auto_prologue_bytes(const pb2_catalog *c1, const pb2_catalog *c2)
{
    const long long b = 16 * (c1->n_pix + c2->n_pix) + 16 * (c1->n_los + c2->n_los) + 256;
    return (b + 255) / 256 * 256;
}

static const int DM_AUTO_BLOCKS = 148 * 2;
static const int DM_CROSS_BLOCKS = 148 * 8;

// pb2_dmat_run.cu: the product kernel of the auto / delta x delta distortion matrix
long long pb2_dmat_run_cta_bytes(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par);
int pb2_dmat_run_blocks(void);
int32_t pb2_launch_dmat_run(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                            const pb2_pairs *pairs, const DmatWork &W, int blocks,
                            double *d_weights_dmat, double *d_dmat, double *d_r_par_eff,
                            double *d_r_trans_eff, double *d_z_eff, double *d_weight_eff,
                            cudaStream_t s);

// Two kernels compute the auto / delta x delta distortion matrix (same results to 1e-9; the GPU
// tests run both):
//   run    pb2_dmat_auto_run_kernel (pb2_dmat_run.cu): run lists + local register tiles; every
//          pixel pair evaluated twice, no dense scratch, and the reference's truncated np.unique
//           reproduced.  Default.
//   dense  pb2_dmat_auto_kernel (this file): per-CTA dense X / Y scratch in global memory + 64x64
//          register-tiled contraction with tile skipping (the round-1 kernel).  Kept for r-mu
//          binning (a run's sums are still exact there, but it is the validated path) and as the
//          cross-check of the run kernel.
// PB2_DMAT_KERNEL=run|dense forces one of them (A/B measurements, cross-checks).
static bool use_run_kernel(const pb2_params *par)
{
    if (par->rmu_binning) return false;
    const char *force = getenv("PB2_DMAT_KERNEL");
    if (force && force[0] == 'd') return false;
    const long long nb = (long long)par->num_bins_r_par * par->num_bins_r_trans;
    const long long nbm = (long long)par->num_model_bins_r_par * par->num_model_bins_r_trans;
    return nb < (1 << 24) && nbm < (1 << 24);   // run keys pack both bins in 24 bits each
}

extern "C" {

int64_t pb2_dmat_scratch_bytes(const pb2_catalog *cat1, const pb2_catalog *cat2,
                               const pb2_params *par, int32_t cross)
{
    if (!cat1 || !cat2 || !par) return 0;
    // kept-pair list (worst case: every pair kept) is sized by the caller's pair count at launch;
    // here: counters + per-CTA areas.  The list itself is appended after them.
    long long bytes = 256;
    if (cross)
        bytes += (long long)DM_CROSS_BLOCKS * (4ll * (cat1->max_pix + 1) * (long long)sizeof(XSeg));
    else if (use_run_kernel(par))
        bytes += (long long)pb2_dmat_run_blocks() * pb2_dmat_run_cta_bytes(cat1, cat2, par) +
                 auto_prologue_bytes(cat1, cat2);
    else
        bytes += (long long)DM_AUTO_BLOCKS * auto_cta_bytes(cat1, cat2, par) +
                 auto_prologue_bytes(cat1, cat2);
    return bytes;
}

static int32_t dmat_launch(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                           const pb2_pairs *pairs, double *d_weights_dmat, double *d_dmat,
                           double *d_r_par_eff, double *d_r_trans_eff, double *d_z_eff,
                           double *d_weight_eff, void *d_scratch, int64_t scratch_bytes,
                           void *stream, bool cross)
{
    if (!cat1 || !cat2 || !par || !pairs || !d_dmat || !d_scratch) {
        pb2_set_error("pb2_dmat: null pointer argument");
        return PB2_EINVAL;
    }
    if (par->ang_correlation) {
        pb2_set_error("pb2_dmat: ang_correlation has no distortion matrix in the reference");
        return PB2_ECONFIG;
    }
    if (!cat1->log_lambda || (!cross && !cat2->log_lambda)) {
        pb2_set_error("pb2_dmat: catalogue has no log_lambda");
        return PB2_EINVAL;
    }
    if (pairs->n_pairs <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const long long fixed = pb2_dmat_scratch_bytes(cat1, cat2, par, cross ? 1 : 0);
    const long long need = fixed + pairs->n_pairs * 8;
    if (scratch_bytes < need) {
        pb2_set_error("pb2_dmat: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes, need);
        return PB2_EINVAL;
    }
    DmatWork W;
    char *p = (char *)d_scratch;
    W.count = (unsigned long long *)p;
    W.stats = cross ? nullptr : (double *)(p + 64);  // inside the 256-byte header, zeroed below
    W.cta_base = p + 256;
    const bool run_kernel = !cross && use_run_kernel(par);
    const int auto_blocks = run_kernel ? pb2_dmat_run_blocks() : DM_AUTO_BLOCKS;
    W.cta_stride = cross ? (4ll * (cat1->max_pix + 1) * (long long)sizeof(XSeg))
                   : run_kernel ? pb2_dmat_run_cta_bytes(cat1, cat2, par)
                                : auto_cta_bytes(cat1, cat2, par);
    W.kept = (long long *)(p + fixed);
    W.rows_max = cross ? (cat1->max_pix + 1) : (2 * cat1->max_pix + 2 * cat2->max_pix + 4);
    W.cap = DM_CAP;
    {
        const double eps = 9.094947017729282e-13;  // 2^-40
        const double span = par->r_par_max - par->r_par_min;
        const double kp = (double)par->num_bins_r_par / span, kt = (double)par->num_bins_r_trans / par->r_trans_max;
        const double mp = (double)par->num_model_bins_r_par / span;
        const double mt = (double)par->num_model_bins_r_trans / par->r_trans_max;
        W.fast.kp_lo = kp * (1. - eps); W.fast.kp_hi = kp * (1. + eps);
        W.fast.kt_lo = kt * (1. - eps); W.fast.kt_hi = kt * (1. + eps);
        W.fast.mp_lo = mp * (1. - eps); W.fast.mp_hi = mp * (1. + eps);
        W.fast.mt_lo = mt * (1. - eps); W.fast.mt_hi = mt * (1. + eps);
        W.fast.same = (par->num_model_bins_r_par == par->num_bins_r_par &&
                       par->num_model_bins_r_trans == par->num_bins_r_trans) ? 1 : 0;
    }
    PB2_CUDA(cudaMemsetAsync(W.count, 0, 256, s));
    pb2_timing_begin(s);
    if (!cross) {
        // prologue arrays sit between the per-CTA areas and the kept-pair list
        char *q = p + 256 + (long long)auto_blocks * W.cta_stride;
        W.fs1 = (double2 *)q;
        W.fs2 = W.fs1 + cat1->n_los;
        W.fz1 = (double *)(W.fs2 + cat2->n_los);
        W.dl1 = W.fz1 + cat1->n_pix;
        W.fz2 = W.dl1 + cat1->n_pix;
        W.dl2 = W.fz2 + cat2->n_pix;
        if (cat1->n_los > 0)
            dmat_prologue_kernel<<<(unsigned)((cat1->n_los * 32 + 255) / 256), 256, 0, s>>>(
                *cat1, *par, par->alpha, W.fz1, W.dl1, W.fs1);
        if (cat2->n_los > 0)
            dmat_prologue_kernel<<<(unsigned)((cat2->n_los * 32 + 255) / 256), 256, 0, s>>>(
                *cat2, *par, par->alpha2, W.fz2, W.dl2, W.fs2);
        pb2_count_launch(2);
    }
    dmat_compact_kernel<<<(unsigned)((pairs->n_pairs + 255) / 256), 256, 0, s>>>(*pairs, W);
    if (cross)
        pb2_dmat_cross_kernel<<<DM_CROSS_BLOCKS, 128, 0, s>>>(*cat1, *cat2, *par, *pairs, W,
                                                             d_weights_dmat, d_dmat, d_r_par_eff,
                                                             d_r_trans_eff, d_z_eff, d_weight_eff);
    else if (run_kernel) {
        int32_t rc0 = pb2_launch_dmat_run(cat1, cat2, par, pairs, W, auto_blocks, d_weights_dmat,
                                          d_dmat, d_r_par_eff, d_r_trans_eff, d_z_eff, d_weight_eff, s);
        if (rc0) return rc0;
    } else
        pb2_dmat_auto_kernel<<<DM_AUTO_BLOCKS, DM_THREADS, 0, s>>>(*cat1, *cat2, *par, *pairs, W,
                                                                  d_weights_dmat, d_dmat,
                                                                  d_r_par_eff, d_r_trans_eff,
                                                                  d_z_eff, d_weight_eff);
    pb2_count_launch(2);
    int32_t rc = pb2_check_launch(cross ? "pb2_dmat_cross_kernel"
                                  : run_kernel ? "pb2_dmat_auto_run_kernel" : "pb2_dmat_auto_kernel");
    pb2_timing_end(s);
    return rc;
}

int32_t pb2_dmat_auto(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                      const pb2_pairs *pairs, double *d_weights_dmat, double *d_dmat,
                      double *d_r_par_eff, double *d_r_trans_eff, double *d_z_eff,
                      double *d_weight_eff, void *d_scratch, int64_t scratch_bytes, void *stream)
{
    return dmat_launch(cat1, cat2, par, pairs, d_weights_dmat, d_dmat, d_r_par_eff, d_r_trans_eff,
                       d_z_eff, d_weight_eff, d_scratch, scratch_bytes, stream, false);
}

int32_t pb2_dmat_cross(const pb2_catalog *cat1, const pb2_catalog *objs, const pb2_params *par,
                       const pb2_pairs *pairs, double *d_weights_dmat, double *d_dmat,
                       double *d_r_par_eff, double *d_r_trans_eff, double *d_z_eff,
                       double *d_weight_eff, void *d_scratch, int64_t scratch_bytes, void *stream)
{
    return dmat_launch(cat1, objs, par, pairs, d_weights_dmat, d_dmat, d_r_par_eff, d_r_trans_eff,
                       d_z_eff, d_weight_eff, d_scratch, scratch_bytes, stream, true);
}

/* measurement: the statistics the last pb2_dmat_auto call on this scratch buffer accumulated --
 * out[0] as-written FP64 ops of the reference algorithm, out[1] sum over used forest pairs of the
 * unique model bins U, out[2] in-range pixel pairs.  Synchronises the stream. */
int32_t pb2_dmat_stats(const void *d_scratch, double *out3, void *stream)
{
    if (!d_scratch || !out3) {
        pb2_set_error("pb2_dmat_stats: null pointer argument");
        return PB2_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    PB2_CUDA(cudaMemcpyAsync(out3, (const char *)d_scratch + 64, 3 * sizeof(double),
                             cudaMemcpyDeviceToHost, s));
    PB2_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // extern "C"
