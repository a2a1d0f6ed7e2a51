// Distortion matrix, product kernel for the standard binning (everything but rmu_binning):
// cf.compute_dmat_forest_pairs_fast, reference py/picca/cf.py:520-887.
//
// Same algebra as pb2_dmat.cu: per forest pair
//     dmat[A,k] += sum_{(i,j) in S,A} w12 zf [k = B(i,j)]
//                  - sum_i ( Q1[A,i] eta1[i,k] + Q1d[A,i] eta3[i,k] )
//                  - sum_j ( Q2[A,j] eta2[j,k] + Q2d[A,j] eta4[j,k] )
//                  + P0[A] eta5[k] + P2[A] eta6[k] + P1[A] eta7[k] + P12[A] eta8[k]
// but nothing of it goes through global memory any more:
//  * a pixel row i of forest 1 meets the pixels j of forest 2 in RUNS of equal (data bin, model
//    bin, selected) -- r_par moves ~0.6 Mpc/h per pixel -- and every sum the reference forms over
//    a run factorises into (row constants) x (sums over the run's columns of w, w dll, fz w,
//    fz w dll, w r_comov, w dist_m, w z).  Those column sums are differences of per-forest PREFIX
//    sums (built once per forest pair, 11 short arrays in an L2-resident slab).  One warp per row,
//    lane = column: the lanes evaluate the exact bins of 32 consecutive columns (sandwich proof or
//    the reference's IEEE expression, pb2_dmat.cuh), a ballot marks where the key changes, and
//    only the last lane of a run touches memory: a handful of shared-memory adds.
//  * rows are processed 16 at a time (one per warp): their Q1 / Q1d / eta1 / eta3 rows live in
//    shared memory as two [32][<=128] blocks, and a rank-32 update takes them into the
//    (data bins) x (model bins) tile of the forest pair held in REGISTERS (8 x 4 per thread,
//    512 threads, compact bin indices); the same for the columns of forest 2 (Q2, Q2d, eta2,
//    eta4).  P0..P12 / eta5..eta8 are column sums of those blocks; the four rank-1 terms and the
//    scatter into dmat (one native red.global.add.f64 per touched cell) happen once per forest pair.
//  * weights_dmat, the effective r_par / r_trans / z / weight and the diagonal term are summed per
//    forest pair in shared memory and flushed once.
// Pixel pairs are evaluated three times (pass 0: touched bins and the early exit of cf.py:570-571;
// row sweep; column sweep), 25 instructions each; the contraction executes
// 2 (n1 + n2) x UA x U DFMAs on compact indices (U, UA ~ 85 at config 4).
// Roofline: FP64 issue.  Sums are re-associated w.r.t. the reference (1e-9; prefix differences add
// ~1e-13), pair counts and the sets of touched bins are exact, including the reference's
// truncated np.unique when same-half-plate close pairs exist (see pass 0).
#include "pb2_dmat.cuh"

#define DR_THREADS 512
#define DR_WARPS 16
#define DR_ROWS 16   // pixel rows per block = one per warp
#define DR_W 128     // compact columns (data bins / model bins) per pass
#define DR_NPF2 7    // prefix arrays of the swept-over forest in the row sweep
#define DR_NPF1 4    // ... in the column sweep

struct DrShared {
    double Xb[2][DR_ROWS][DR_W];   // Q1, Q1d (row sweep) / Q2, Q2d (column sweep), permuted columns
    double Yb[2][DR_ROWS][DR_W];   // eta1, eta3 / eta2, eta4
    double vP[4][DR_W];            // P0, P2, P1, P12   by compact data bin
    double vE[4][DR_W];            // eta5 .. eta8       by compact model bin
    double vO[5][DR_W];            // weight_eff, r_par_eff, r_trans_eff, z_eff, diagonal term
    double Ob[5][DR_ROWS][DR_W];   // ... their per-row parts during the row sweep
    double rowc[DR_ROWS][4];       // dll_i, w_i / sw1, w_i dll_i / swsll1, (unused)
    long long e;
    int cnt[2];
    int U, UA;
};

// position of compact column k inside a block row: thread (ty, tx) owns the data bins ty + 16 p and
// the model bins tx + 32 q, stored contiguously per thread so that its operands are 128-bit loads
__device__ __forceinline__ int dr_xpos(int ka) { return ((ka & 15) << 3) | (ka >> 4); }
__device__ __forceinline__ int dr_ypos(int kb) { return ((kb & 31) << 2) | (kb >> 5); }

// inclusive prefix sums of f(x) over x = 0..n-1 into out[1..n] (out[0] = 0), one warp
template <typename F>
__device__ __forceinline__ void dr_warp_prefix(double *__restrict__ out, int n, int lane, F f)
{
    double carry = 0.;
    if (lane == 0) out[0] = 0.;
    for (int b = 0; b < n; b += 32) {
        const int x = b + lane;
        double v = x < n ? f(x) : 0.;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += t;
        }
        v += carry;
        if (x < n) out[x + 1] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

struct DrPair {
    // everything a sweep needs about the forest pair
    const double *rc1, *dm1, *z1, *w1, *f1z, *dl1;
    const double *rc2, *dm2, *z2, *w2, *f2z, *dl2;
    int n1, n2, order1, order2;
    double ch, sh, zq1, zq2, sw1, swsll1, sw2, swsll2;
    bool zerr_on, shp, windows;
};

// selection of an in-range pixel pair beyond the geometry (cf.py:629-658, :669-671): z-pair cut,
// zerr cut on either side, same-half-plate close pairs
__device__ __forceinline__ bool dr_selected(const pb2_params &P, const DrPair &D, const DmatGeom &g,
                                            double zi, double zj, bool i_sel)
{
    if (!i_sel || g.close) return false;
    if (P.has_z_min_pairs || P.has_z_max_pairs) {
        const double z = div_rn(add_rn(zi, zj), 2.);
        if ((P.has_z_min_pairs && z < P.z_min_pairs) || (P.has_z_max_pairs && z > P.z_max_pairs))
            return false;
    }
    if (D.zerr_on && pb2_zerr_close(P, zj, D.zq1)) return false;
    return true;
}

// ------------------------------------------------------------------------------------------
// one row of the row sweep (ROW = true: pixel i of forest 1 against the columns of forest 2) or of
// the column sweep (ROW = false: pixel j of forest 2 against the pixels of forest 1), whole warp
// ------------------------------------------------------------------------------------------
template <bool ROW>
__device__ __forceinline__ void dr_sweep_row(
    const pb2_params &P, const DmatFast &F, const DrPair &D, DrShared &S, int r, int px,
    const double *__restrict__ pf, int pf_stride, const int *__restrict__ kidx,
    const int *__restrict__ aidx, int kc, int Uc, int ac, int UAc, bool first, bool big,
    long long nbm, double *__restrict__ dmat, double *__restrict__ r_par_eff,
    double *__restrict__ r_trans_eff, double *__restrict__ z_eff, double *__restrict__ weight_eff)
{
    const int lane = threadIdx.x & 31;
    // the fixed pixel and the swept forest
    const double rc_f = ROW ? D.rc1[px] : D.rc2[px], dm_f = ROW ? D.dm1[px] : D.dm2[px];
    const double z_f = ROW ? D.z1[px] : D.z2[px], w_f = ROW ? D.w1[px] : D.w2[px];
    const double fz_f = ROW ? D.f1z[px] : D.f2z[px];
    const double *__restrict__ rcs = ROW ? D.rc2 : D.rc1, *__restrict__ dms = ROW ? D.dm2 : D.dm1;
    const double *__restrict__ zs = ROW ? D.z2 : D.z1, *__restrict__ ws = ROW ? D.w2 : D.w1;
    const int ns = ROW ? D.n2 : D.n1;
    // row sweep: zerr cut of the fixed pixel of forest 1 against quasar 2 (cf.py:629-636)
    bool f_sel = true;
    if (ROW && D.zerr_on && pb2_zerr_close(P, z_f, D.zq2)) f_sel = false;
    int lo, hi;
    if (ROW) row_window(P, D.windows, rc_f, dm_f, rcs, dms, ns, D.ch, D.sh, P.x_correlation, lo, hi);
    else col_window(P, D.windows, rc_f, dm_f, rcs, dms, ns, D.ch, D.sh, P.x_correlation, lo, hi);
    // normalisations of the eta rows (cf.py:767-813): the swept forest's sums
    const double inv_e = ROW ? 1. / D.sw2 : 1. / D.sw1;
    const double inv_e3 = ROW ? 1. / D.swsll2 : 1. / D.swsll1;
    const bool has_e3 = ROW ? (D.order2 == 1) : (D.order1 == 1);
    const double rc0 = rcs[0], dm0 = dms[0];
    for (int sb = lo; sb < hi; sb += 32) {
        const int s = sb + lane;
        long long key = -1;
        bool sel = false;
        int A = 0, B = 0;
        if (s < hi && ws[s] != 0.) {
            const DmatGeom g = ROW ? dmat_pair(P, F, rc_f, dm_f, rcs[s], dms[s], D.ch, D.sh, false, D.shp)
                                   : dmat_pair(P, F, rcs[s], dms[s], rc_f, dm_f, D.ch, D.sh, false, D.shp);
            if (g.in) {
                // pixel of forest 1 = the fixed one (row sweep) or the swept one (column sweep)
                const bool i_sel = ROW ? f_sel
                                       : !(D.zerr_on && pb2_zerr_close(P, zs[s], D.zq2));
                sel = dr_selected(P, D, g, ROW ? z_f : zs[s], ROW ? zs[s] : z_f, i_sel);
                A = g.A;
                B = g.B;
                // the sign of r_par before abs() decides the sign of a run's sum of r_par
                const long long sgn = (!P.x_correlation && (ROW ? rc_f < rcs[s] : rcs[s] < rc_f)) ? 1 : 0;
                key = (long long)A | ((long long)B << 24) | ((long long)(sel ? 1 : 0) << 48) |
                      (sgn << 49);
            }
        }
        const long long prev = __shfl_up_sync(0xffffffffu, key, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != key || key < 0);
        // the last lane of every run writes the run's sums into the row's cells
        const bool tail = key >= 0 && (lane == 31 || ((heads >> (lane + 1)) & 1u));
        int kb = -1, ka = -1;
        if (tail) {
            kb = kidx[B] - kc;       // (< 0: a model bin outside the truncated np.unique, Q8)
            if (kb >= Uc) kb = -1;
            if (sel) {
                ka = aidx[A] - ac;
                if (ka < 0 || ka >= UAc) ka = -1;
            }
        }
        // The row's cells belong to this warp: plain read-add-write, no atomics (shared-memory
        // fp64 atomics are compare-and-swap loops whose latency would bound the sweep).  Two runs
        // of ONE step can still meet in a cell (the bins fold back where r_par changes sign; runs
        // that differ only in `selected`): then the tails take turns.
        const unsigned same_b = __match_any_sync(0xffffffffu, kb >= 0 ? kb : -1 - lane);
        const unsigned same_a = __match_any_sync(0xffffffffu, ka >= 0 ? ka : -1 - lane);
        const bool clash = __any_sync(0xffffffffu, (same_b & (same_b - 1)) || (same_a & (same_a - 1)));
        unsigned turns = clash ? __ballot_sync(0xffffffffu, tail) : 1u;
        while (turns) {
            const bool mine = tail && (!clash || lane == __ffs(turns) - 1);
            if (mine) {
                const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
                const int s0 = sb + start, s1 = s + 1;   // the run covers swept pixels [s0, s1)
                const double S0 = pf[0 * pf_stride + s1] - pf[0 * pf_stride + s0];   // sum w
                const double F0 = pf[2 * pf_stride + s1] - pf[2 * pf_stride + s0];   // sum fz w
                if (kb >= 0) {
                    S.Yb[0][r][dr_ypos(kb)] += fz_f * F0 * inv_e;                    // eta1 / eta2
                    if (has_e3) {
                        const double F1 = pf[3 * pf_stride + s1] - pf[3 * pf_stride + s0];  // sum fz w dll
                        S.Yb[1][r][dr_ypos(kb)] += fz_f * F1 * inv_e3;               // eta3 / eta4
                    }
                }
                if (ka >= 0) {
                    const double S1 = pf[1 * pf_stride + s1] - pf[1 * pf_stride + s0];   // sum w dll
                    S.Xb[0][r][dr_xpos(ka)] += w_f * S0;                             // Q1 / Q2
                    S.Xb[1][r][dr_xpos(ka)] += w_f * S1;                             // Q1d / Q2d
                }
                if (ROW && first && sel) {
                    // cf.py:714-718, :873 summed over the run
                    const double R = pf[4 * pf_stride + s1] - pf[4 * pf_stride + s0];   // sum w (rc - rc0)
                    const double Dm = pf[5 * pf_stride + s1] - pf[5 * pf_stride + s0];  // sum w (dm - dm0)
                    const double Z = pf[6 * pf_stride + s1] - pf[6 * pf_stride + s0];   // sum w z
                    double rp = D.ch * ((rc_f - rc0) * S0 - R);
                    if (!P.x_correlation) rp = fabs(rp);
                    const double rt = D.sh * ((dm_f + dm0) * S0 + Dm);
                    const double zz = 0.5 * (z_f * S0 + Z);
                    const double dg = w_f * fz_f * F0;
                    if (!big && kb >= 0) {
                        S.Ob[0][r][kb] += w_f * S0;
                        S.Ob[1][r][kb] += w_f * rp;
                        S.Ob[2][r][kb] += w_f * rt;
                        S.Ob[3][r][kb] += w_f * zz;
                        if (F.same) S.Ob[4][r][kb] += dg;
                        else atomic_add_f64(dmat + (long long)A * nbm + B, dg);
                    } else {
                        atomic_add_f64(weight_eff + B, w_f * S0);
                        atomic_add_f64(r_par_eff + B, w_f * rp);
                        atomic_add_f64(r_trans_eff + B, w_f * rt);
                        atomic_add_f64(z_eff + B, w_f * zz);
                        atomic_add_f64(dmat + (long long)A * nbm + B, dg);
                    }
                }
            }
            turns = clash ? (turns & (turns - 1)) : 0u;
            __syncwarp();
        }
    }
}

// rank-(2 DR_ROWS) update of the register tile: c[p][q] -= sum_r Xb[.][r][ty + 16 p] Yb[.][r][tx + 32 q]
__device__ __forceinline__ void dr_rank_update(const DrShared &S, double (&c)[8][4], int np_, int nq_)
{
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    if (np_ == 0 || nq_ == 0) return;
#pragma unroll 1
    for (int r = 0; r < DR_ROWS; r++) {
#pragma unroll
        for (int kind = 0; kind < 2; kind++) {
            const double4 xa = *reinterpret_cast<const double4 *>(&S.Xb[kind][r][ty * 8]);
            const double4 xb = *reinterpret_cast<const double4 *>(&S.Xb[kind][r][ty * 8 + 4]);
            const double4 yv = *reinterpret_cast<const double4 *>(&S.Yb[kind][r][tx * 4]);
            const double x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const double y[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int p = 0; p < 8; p++) {
                if (p < np_) {
#pragma unroll
                    for (int q = 0; q < 4; q++) c[p][q] = fma(-x[p], y[q], c[p][q]);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(DR_THREADS, 1)
pb2_dmat_auto_run_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, DmatWork W,
                         double *__restrict__ weights_dmat, double *__restrict__ dmat,
                         double *__restrict__ r_par_eff, double *__restrict__ r_trans_eff,
                         double *__restrict__ z_eff, double *__restrict__ weight_eff)
{
    extern __shared__ __align__(16) unsigned char dr_smem[];
    DrShared &S = *reinterpret_cast<DrShared *>(dr_smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const int nbm = P.num_model_bins_r_par * P.num_model_bins_r_trans;
    const double zerr_ang = mul_rn(P.zerr_cut_deg, PB2_PI) / 180.0;

    // per-CTA scratch (global, L2-resident): compact-index tables, per-row counts, prefix sums
    char *base = W.cta_base + (long long)blockIdx.x * W.cta_stride;
    int *kidx = (int *)base;                     // [nbm] compact model index, -1 = untouched
    int *aidx = kidx + nbm;                      // [nb]
    int *klist = aidx + nb;                      // [nbm]
    int *alist = klist + nbm;                    // [nb]
    int *rowcnt = alist + nb;                    // [max_pix1 + 1] in-range pairs per row (Q8)
    const int pstride = max(c1.max_pix, c2.max_pix) + 1;
    double *pf2 = (double *)(((uintptr_t)(rowcnt + c1.max_pix + 2) + 15) & ~(uintptr_t)15);
    double *pf1 = pf2 + (long long)DR_NPF2 * pstride;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const unsigned long long t = atomicAdd(W.count + 1, 1ull);
            S.e = (t < W.count[0]) ? W.kept[t] : -1;
            S.cnt[0] = S.cnt[1] = 0;
        }
        __syncthreads();
        const long long e = S.e;
        if (e < 0) break;

        DrPair D;
        const int f1 = pr.f1_index[pr.nb_f1[e]], f2 = pr.nb_f2[e];
        const long long a = c1.offset[f1], b = c2.offset[f2];
        D.n1 = (int)(c1.offset[f1 + 1] - a);
        D.n2 = (int)(c2.offset[f2 + 1] - b);
        if (D.n1 <= 0 || D.n2 <= 0) continue;
        D.ch = pr.nb_cos[e];
        D.sh = pr.nb_sin[e];
        D.zerr_on = P.has_zerr_cut && (pr.nb_ang[e] < zerr_ang);
        D.shp = P.remove_same_half_plate_close_pairs && pb2_same_half_plate(c1, c2, f1, f2);
        D.windows = c1.sorted && c2.sorted;
        D.zq1 = c1.z_qso[f1];
        D.zq2 = c2.z_qso[f2];
        D.order1 = c1.order[f1];
        D.order2 = c2.order[f2];
        D.rc1 = c1.r_comov + a; D.dm1 = c1.dist_m + a; D.z1 = c1.z + a; D.w1 = c1.weights + a;
        D.rc2 = c2.r_comov + b; D.dm2 = c2.dist_m + b; D.z2 = c2.z + b; D.w2 = c2.weights + b;
        D.f1z = W.fz1 + a; D.dl1 = W.dl1 + a; D.f2z = W.fz2 + b; D.dl2 = W.dl2 + b;
        D.sw1 = W.fs1[f1].x; D.swsll1 = W.fs1[f1].y;
        D.sw2 = W.fs2[f2].x; D.swsll2 = W.fs2[f2].y;
        const int n1 = D.n1, n2 = D.n2;

        // ---------------- pass 0: touched bins, pair counts (cf.py:547-571 + the bins of pass 1)
        for (int x = tid; x < nbm; x += DR_THREADS) kidx[x] = -1;
        for (int x = tid; x < nb; x += DR_THREADS) aidx[x] = -1;
        __syncthreads();
        {
            int cnt_nc = 0, cnt_in = 0;
            for (int i = warp; i < n1; i += DR_WARPS) {
                int row_in = 0;
                if (D.w1[i] != 0.) {
                    bool i_sel = true;
                    if (D.zerr_on && pb2_zerr_close(P, D.z1[i], D.zq2)) i_sel = false;
                    int lo, hi;
                    row_window(P, D.windows, D.rc1[i], D.dm1[i], D.rc2, D.dm2, n2, D.ch, D.sh,
                               P.x_correlation, lo, hi);
                    for (int jb = lo; jb < hi; jb += 32) {
                        const int j = jb + lane;
                        bool in = false, nc = false;
                        if (j < hi && D.w2[j] != 0.) {
                            const DmatGeom g = dmat_pair(P, W.fast, D.rc1[i], D.dm1[i], D.rc2[j],
                                                         D.dm2[j], D.ch, D.sh, false, D.shp);
                            if (g.in) {
                                in = true;
                                nc = !g.close;
                                kidx[g.B] = 0;
                                if (dr_selected(P, D, g, D.z1[i], D.z2[j], i_sel)) aidx[g.A] = 0;
                            }
                        }
                        const int n_in = __popc(__ballot_sync(0xffffffffu, in));
                        cnt_in += n_in;
                        row_in += n_in;
                        cnt_nc += __popc(__ballot_sync(0xffffffffu, nc));
                    }
                }
                if (lane == 0) rowcnt[i] = row_in;
            }
            if (lane == 0) {
                if (cnt_nc) atomicAdd(&S.cnt[0], cnt_nc);
                if (cnt_in) atomicAdd(&S.cnt[1], cnt_in);
            }
        }
        __syncthreads();
        if (S.cnt[0] == 0) continue;  // cf.py:570-571
        if (S.cnt[1] != S.cnt[0]) {
            // same-half-plate close pairs are in range but not counted by the
            // reference's pass 0 (cf.py:565-568), so its `all_model_bins` -- sized by that count,
            // filled by EVERY in-range pair in (i, j) order (cf.py:702-703) -- only keeps the model
            // bins of the first `count` in-range pairs inside the array np.unique sees
            // (cf.py:846-848).  Re-mark the model bins with that rank limit.
            const int limit = S.cnt[0];
            __syncthreads();
            for (int x = tid; x < nbm; x += DR_THREADS) kidx[x] = -1;
            if (warp == 0) {   // exclusive scan of the per-row counts, in place
                int carry = 0;
                for (int ib = 0; ib < n1; ib += 32) {
                    const int i = ib + lane;
                    const int v = i < n1 ? rowcnt[i] : 0;
                    int incl = v;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += t;
                    }
                    if (i < n1) rowcnt[i] = carry + incl - v;
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            __syncthreads();
            for (int i = warp; i < n1; i += DR_WARPS) {
                if (D.w1[i] == 0.) continue;
                int rank = rowcnt[i];
                if (rank >= limit) continue;
                int lo, hi;
                row_window(P, D.windows, D.rc1[i], D.dm1[i], D.rc2, D.dm2, n2, D.ch, D.sh,
                           P.x_correlation, lo, hi);
                for (int jb = lo; jb < hi && rank < limit; jb += 32) {
                    const int j = jb + lane;
                    bool in = false;
                    DmatGeom g;
                    g.B = 0;
                    if (j < hi && D.w2[j] != 0.) {
                        g = dmat_pair(P, W.fast, D.rc1[i], D.dm1[i], D.rc2[j], D.dm2[j], D.ch, D.sh,
                                      false, D.shp);
                        in = g.in;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (in && rank + __popc(m & ((1u << lane) - 1u)) < limit) kidx[g.B] = 0;
                    rank += __popc(m);
                }
            }
            __syncthreads();
        }

        // compact indices of the touched bins (the set of np.unique, cf.py:846-848), r_trans-major
        if (warp < 2) {
            int *idx = warp == 0 ? kidx : aidx;
            int *list = warp == 0 ? klist : alist;
            const int nt_ = warp == 0 ? P.num_model_bins_r_trans : P.num_bins_r_trans;
            const int np2 = warp == 0 ? P.num_model_bins_r_par : P.num_bins_r_par;
            const int tot = nt_ * np2;
            int u = 0;
            for (int xb = 0; xb < tot; xb += 32) {
                const int o = xb + lane;           // position in r_trans-major order
                int x = 0;
                bool on = false;
                if (o < tot) {
                    const int t = o / np2, q = o - t * np2;
                    x = t + nt_ * q;
                    on = idx[x] == 0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, on);
                if (on) {
                    const int k = u + __popc(m & ((1u << lane) - 1u));
                    idx[x] = k;
                    list[k] = x;
                }
                u += __popc(m);
            }
            if (lane == 0) {
                if (warp == 0) S.U = u; else S.UA = u;
            }
        }
        // prefix sums of both forests (11 arrays, one warp each)
        if (warp >= 2 && warp < 2 + DR_NPF2 + DR_NPF1) {
            const int k = warp - 2;
            const double *w2 = D.w2, *dl2 = D.dl2, *f2z = D.f2z, *w1 = D.w1, *dl1 = D.dl1, *f1z = D.f1z;
            const double *rc2 = D.rc2, *dm2 = D.dm2, *z2 = D.z2;
            const double rc0 = rc2[0], dm0 = dm2[0];
            switch (k) {
            case 0: dr_warp_prefix(pf2 + 0 * pstride, n2, lane, [&](int x) { return w2[x]; }); break;
            case 1: dr_warp_prefix(pf2 + 1 * pstride, n2, lane, [&](int x) { return w2[x] * dl2[x]; }); break;
            case 2: dr_warp_prefix(pf2 + 2 * pstride, n2, lane, [&](int x) { return f2z[x] * w2[x]; }); break;
            case 3: dr_warp_prefix(pf2 + 3 * pstride, n2, lane, [&](int x) { return f2z[x] * w2[x] * dl2[x]; }); break;
            case 4: dr_warp_prefix(pf2 + 4 * pstride, n2, lane, [&](int x) { return w2[x] * (rc2[x] - rc0); }); break;
            case 5: dr_warp_prefix(pf2 + 5 * pstride, n2, lane, [&](int x) { return w2[x] * (dm2[x] - dm0); }); break;
            case 6: dr_warp_prefix(pf2 + 6 * pstride, n2, lane, [&](int x) { return w2[x] * z2[x]; }); break;
            case 7: dr_warp_prefix(pf1 + 0 * pstride, n1, lane, [&](int x) { return w1[x]; }); break;
            case 8: dr_warp_prefix(pf1 + 1 * pstride, n1, lane, [&](int x) { return w1[x] * dl1[x]; }); break;
            case 9: dr_warp_prefix(pf1 + 2 * pstride, n1, lane, [&](int x) { return f1z[x] * w1[x]; }); break;
            default: dr_warp_prefix(pf1 + 3 * pstride, n1, lane, [&](int x) { return f1z[x] * w1[x] * dl1[x]; }); break;
            }
        }
        __syncthreads();
        const int U = S.U, UA = S.UA;
        if (tid == 0 && W.stats) {
            atomicAdd(W.stats, (double)S.cnt[0] * (15. * U + 4.) + 40. * (double)S.cnt[1]);
            atomicAdd(W.stats + 1, (double)U);
            atomicAdd(W.stats + 2, (double)S.cnt[1]);
        }
        const bool big = U > DR_W || UA > DR_W;

        for (int kc = 0; kc < U; kc += DR_W) {
            const int Uc = min(DR_W, U - kc);
            for (int ac = 0; ac < max(UA, 1); ac += DR_W) {
                const int UAc = min(DR_W, UA - ac);   // <= 0: no selected pair (effective sums only)
                const bool first = (kc == 0 && ac == 0);
                const int np_ = UAc > 0 ? min(8, (UAc - (tid >> 5) + 15) / 16) : 0;   // ka = ty + 16 p
                const int nq_ = min(4, (Uc - (tid & 31) + 31) / 32);                   // kb = tx + 32 q
                double c[8][4];
#pragma unroll
                for (int p = 0; p < 8; p++)
#pragma unroll
                    for (int q = 0; q < 4; q++) c[p][q] = 0.;
                __syncthreads();
                for (int x = tid; x < 4 * DR_W; x += DR_THREADS) {
                    (&S.vP[0][0])[x] = 0.;
                    (&S.vE[0][0])[x] = 0.;
                }
                for (int x = tid; x < 5 * DR_W; x += DR_THREADS) (&S.vO[0][0])[x] = 0.;

                // ---------------- row sweep: pixels of forest 1, 16 at a time
                for (int ib = 0; ib < n1; ib += DR_ROWS) {
                    __syncthreads();
                    for (int x = tid; x < 2 * DR_ROWS * DR_W; x += DR_THREADS) {
                        (&S.Xb[0][0][0])[x] = 0.;
                        (&S.Yb[0][0][0])[x] = 0.;
                    }
                    if (first && !big)
                        for (int x = tid; x < 5 * DR_ROWS * DR_W; x += DR_THREADS)
                            (&S.Ob[0][0][0])[x] = 0.;
                    __syncthreads();
                    const int i = ib + warp;
                    const bool rowok = i < n1 && D.w1[i] != 0.;
                    if (lane == 0) {
                        const double wi = rowok ? D.w1[i] : 0., dli = rowok ? D.dl1[i] : 0.;
                        S.rowc[warp][0] = dli;
                        S.rowc[warp][1] = wi / D.sw1;
                        S.rowc[warp][2] = wi * dli / D.swsll1;
                    }
                    if (rowok)
                        dr_sweep_row<true>(P, W.fast, D, S, warp, i, pf2, pstride, kidx, aidx, kc, Uc,
                                           ac, UAc, first, big, nbm, dmat, r_par_eff, r_trans_eff,
                                           z_eff, weight_eff);
                    __syncthreads();
                    // column sums of the block: P0, P2, P1, P12 and eta5 .. eta8 (cf.py:775-843)
                    if (tid < DR_W) {
                        const int ka = tid;
                        if (ka < UAc) {
                            double p0 = 0., p2 = 0., p1 = 0., p12 = 0.;
                            const int xp = dr_xpos(ka);
#pragma unroll 4
                            for (int r = 0; r < DR_ROWS; r++) {
                                const double q1 = S.Xb[0][r][xp], q1d = S.Xb[1][r][xp];
                                const double dli = S.rowc[r][0];
                                p0 += q1;
                                p2 += q1d;
                                p1 += dli * q1;
                                p12 += dli * q1d;
                            }
                            S.vP[0][ka] += p0;
                            S.vP[1][ka] += p2;
                            S.vP[2][ka] += p1;
                            S.vP[3][ka] += p12;
                        }
                    } else if (tid < 2 * DR_W) {
                        const int kb = tid - DR_W;
                        if (kb < Uc) {
                            double e5 = 0., e6 = 0., e7 = 0., e8 = 0.;
                            const int yp = dr_ypos(kb);
#pragma unroll 4
                            for (int r = 0; r < DR_ROWS; r++) {
                                const double e1 = S.Yb[0][r][yp], e3 = S.Yb[1][r][yp];
                                const double a5 = S.rowc[r][1], a7 = S.rowc[r][2];
                                e5 += a5 * e1;
                                e6 += a5 * e3;
                                e7 += a7 * e1;
                                e8 += a7 * e3;
                            }
                            S.vE[0][kb] += e5;
                            if (D.order2 == 1) S.vE[1][kb] += e6;
                            if (D.order1 == 1) S.vE[2][kb] += e7;
                            if (D.order1 == 1 && D.order2 == 1) S.vE[3][kb] += e8;
                        }
                    } else if (tid < 3 * DR_W && first && !big) {
                        const int kb = tid - 2 * DR_W;
                        if (kb < Uc) {
#pragma unroll
                            for (int o = 0; o < 5; o++) {
                                double t = 0.;
#pragma unroll 4
                                for (int r = 0; r < DR_ROWS; r++) t += S.Ob[o][r][kb];
                                S.vO[o][kb] += t;
                            }
                        }
                    }
                    dr_rank_update(S, c, np_, nq_);
                }

                // ---------------- column sweep: pixels of forest 2
                for (int jb = 0; jb < n2; jb += DR_ROWS) {
                    __syncthreads();
                    for (int x = tid; x < 2 * DR_ROWS * DR_W; x += DR_THREADS) {
                        (&S.Xb[0][0][0])[x] = 0.;
                        (&S.Yb[0][0][0])[x] = 0.;
                    }
                    __syncthreads();
                    const int j = jb + warp;
                    if (j < n2 && D.w2[j] != 0.)
                        dr_sweep_row<false>(P, W.fast, D, S, warp, j, pf1, pstride, kidx, aidx, kc, Uc,
                                            ac, UAc, false, big, nbm, dmat, r_par_eff, r_trans_eff,
                                            z_eff, weight_eff);
                    __syncthreads();
                    dr_rank_update(S, c, np_, nq_);
                }
                __syncthreads();

                // ---------------- rank-1 terms, scatter of the tile, per-forest-pair sums
                {
                    const int ty = tid >> 5, tx = tid & 31;
#pragma unroll
                    for (int p = 0; p < 8; p++) {
                        const int ka = ty + 16 * p;
                        if (p >= np_ || ka >= UAc) continue;
                        const double p0 = S.vP[0][ka], p2 = S.vP[1][ka], p1 = S.vP[2][ka], p12 = S.vP[3][ka];
                        const long long rowp = (long long)alist[ac + ka] * nbm;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int kb = tx + 32 * q;
                            if (q >= nq_ || kb >= Uc) continue;
                            const double v = c[p][q] + p0 * S.vE[0][kb] + p2 * S.vE[1][kb] +
                                             p1 * S.vE[2][kb] + p12 * S.vE[3][kb];
                            if (v != 0.) atomic_add_f64(dmat + rowp + klist[kc + kb], v);
                        }
                    }
                    if (first) {
                        // weights_dmat[A] = sum over the selected pairs of the bin = P0[A] (cf.py:718)
                        for (int ka = tid; ka < UAc; ka += DR_THREADS)
                            if (S.vP[0][ka] != 0.) atomic_add_f64(weights_dmat + alist[ac + ka], S.vP[0][ka]);
                        if (!big) {
                            for (int kb = tid; kb < Uc; kb += DR_THREADS) {
                                const int Bm = klist[kc + kb];
                                if (S.vO[0][kb] != 0.) {
                                    atomic_add_f64(weight_eff + Bm, S.vO[0][kb]);
                                    atomic_add_f64(r_par_eff + Bm, S.vO[1][kb]);
                                    atomic_add_f64(r_trans_eff + Bm, S.vO[2][kb]);
                                    atomic_add_f64(z_eff + Bm, S.vO[3][kb]);
                                    if (W.fast.same)
                                        atomic_add_f64(dmat + (long long)Bm * nbm + Bm, S.vO[4][kb]);
                                }
                            }
                        }
                    } else if (kc == 0) {
                        // further data-bin passes of a big forest pair: their weights_dmat
                        for (int ka = tid; ka < UAc; ka += DR_THREADS)
                            if (S.vP[0][ka] != 0.) atomic_add_f64(weights_dmat + alist[ac + ka], S.vP[0][ka]);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
long long pb2_dmat_run_cta_bytes(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par)
{
    const long long nb = (long long)par->num_bins_r_par * par->num_bins_r_trans;
    const long long nbm = (long long)par->num_model_bins_r_par * par->num_model_bins_r_trans;
    const long long pstride = (c1->max_pix > c2->max_pix ? c1->max_pix : c2->max_pix) + 1;
    long long bytes = (2 * nb + 2 * nbm) * 4 + (c1->max_pix + 2) * 4 + 64;
    bytes += (DR_NPF2 + DR_NPF1) * pstride * 8;
    return (bytes + 255) / 256 * 256;
}

int pb2_dmat_run_blocks(void)
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

int32_t pb2_launch_dmat_run(const pb2_catalog *cat1, const pb2_catalog *cat2, const pb2_params *par,
                            const pb2_pairs *pairs, const DmatWork &W, int blocks,
                            double *d_weights_dmat, double *d_dmat, double *d_r_par_eff,
                            double *d_r_trans_eff, double *d_z_eff, double *d_weight_eff,
                            cudaStream_t s)
{
    const size_t smem = sizeof(DrShared);
    PB2_CUDA(cudaFuncSetAttribute(pb2_dmat_auto_run_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pb2_dmat_auto_run_kernel<<<blocks, DR_THREADS, smem, s>>>(*cat1, *cat2, *par, *pairs, W,
                                                              d_weights_dmat, d_dmat, d_r_par_eff,
                                                              d_r_trans_eff, d_z_eff, d_weight_eff);
    return 0;
}
