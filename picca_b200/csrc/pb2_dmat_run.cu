// Distortion matrix, run-list kernel for the standard binning (everything but rmu_binning):
// cf.compute_dmat_forest_pairs_fast, reference py/picca/cf.py:520-887.
//
// Same algebra as pb2_dmat.cu: per forest pair
//     dmat[A,k] += sum_{(i,j) in S,A} w12 zf [k = B(i,j)]
//                  - sum_i ( Q1[A,i] eta1[i,k] + Q1d[A,i] eta3[i,k] )
//                  - sum_j ( Q2[A,j] eta2[j,k] + Q2d[A,j] eta4[j,k] )
//                  + P0[A] eta5[k] + P2[A] eta6[k] + P1[A] eta7[k] + P12[A] eta8[k]
// What the measurements of round 2 say about a forest pair of config 4 (two forests of ~500
// pixels, 1e5 in-range pixel pairs): it touches U ~ 95 model bins (up to 260), but ONE pixel row
// only ~34 of them, in runs of ~6 columns; the dense X / Y scratch of pb2_dmat.cu is therefore
// mostly zeros, its contraction mostly multiplications by zero, and its sweeps spend their time
// in ~3e5 global reductions per forest pair.  Here:
//  S  each pixel pair is evaluated twice (once per sweep; pb2_dmat.cu: three times): a thread
//     walks the row of one pixel of forest 1 (then: one pixel of forest 2), accumulates the sums
//     of a RUN of equal (data bin, model bin, selected) in registers and appends one record per
//     run to the row's list in an L2-resident slab -- plain stores, no atomics, no zero fill.
//     The pair counts, the touched bins (the set of np.unique, cf.py:846-848, including its
//     truncation when same-half-plate close pairs exist) come out of the same sweep.
//  V  the per-bin vectors (P0..P12, eta5..eta8, weights_dmat, the effective r_par / r_trans / z /
//     weight, the diagonal term) are summed from the run records into shared memory.
//  C  contraction: 16 rows at a time.  The bins those rows touch (~45) get LOCAL indices, their
//     runs are expanded into two [32][64] blocks in shared memory, a rank-32 update takes them
//     into a 64 x 64 register tile, and the tile is added into the forest pair's (data bins) x
//     (model bins) matrix in shared memory (128 x 128 compact bins; larger forest pairs replay the
//     run lists per 128 x 128 window -- the sweeps are not repeated).
//  F  rank-1 terms and ONE native red.global.add.f64 per touched cell of dmat.
// Roofline: FP64 issue / shared-memory bandwidth.  Sums are re-associated w.r.t. the reference
// (1e-9 tolerance); pair counts and the sets of touched bins are exact.
// (An earlier form of this kernel -- lane = column, prefix sums, 128 x 128 register tile -- is kept
// under profiles/experiments/r02_dmat_prefix_run_kernel.cu.txt: 2.4x slower than pb2_dmat.cu.)
#include "pb2_dmat.cuh"

#define RL_THREADS 512
#define RL_WARPS 16
#define RL_ROWS 16    // rows per contraction group = one per warp
#define RL_LOC 128    // local bins of a group (data and model): up to the whole window
#define RL_W 128      // window of compact bins held in shared memory
#define RL_VEC 256    // compact bins whose per-pair vectors live in shared memory
#define RL_SEL (1 << 30)

struct RlRun1 {       // a run of the row sweep (pixel i of forest 1 against forest 2)
    int A, Bs, n, pad;   // data bin, model bin | RL_SEL when selected, in-range pixel pairs
    double e1, e3;       // sum zf w2, sum zf w2 dll2        (eta1, eta3 before normalisation)
    double q1, q1d;      // w1 sum w2, w1 sum w2 dll2        (selected runs; Q1, Q1d)
    double rp, rt, zz, dg;  // sums of w12 r_par, w12 r_trans, w12 z, w12 zf (selected runs)
};
struct RlRun2 {       // a run of the column sweep (pixel j of forest 2 against forest 1)
    int A, Bs;
    double e2, e4, q2, q2d;
};

struct RlShared {
    double C[RL_W][RL_W];              // (compact data bin, compact model bin) window
    double Xl[2][RL_ROWS][RL_LOC];     // Q1, Q1d / Q2, Q2d of the group, local permuted columns
    double Yl[2][RL_ROWS][RL_LOC];     // eta1, eta3 / eta2, eta4
    double vP[4][RL_VEC];              // P0, P2, P1, P12       by compact data bin
    double vE[4][RL_VEC];              // eta5 .. eta8          by compact model bin
    double vO[5][RL_VEC];              // weight_eff, r_par_eff, r_trans_eff, z_eff, diagonal
    short locA[RL_W], locB[RL_W];      // window-relative compact bin -> local index (-1: absent)
    short lstA[RL_LOC], lstB[RL_LOC];  // local index -> window-relative compact bin
    unsigned char flgA[RL_W], flgB[RL_W];
    long long e;
    int cnt[2];
    int U, UA, nA, nB;
};

// thread (ty, tx) of the contraction owns the local data bins ty + 16 p (p < 8) and the local
// model bins tx + 32 q (q < 4); a block row stores them contiguously per thread.  A group of 16
// rows touches ~43 bins (p90: 71, max ~105 at config 4), so p < 3, q < 2 is the common case.
__device__ __forceinline__ int rl_xpos(int la) { return ((la & 15) << 3) | (la >> 4); }
__device__ __forceinline__ int rl_ypos(int lb) { return ((lb & 31) << 2) | (lb >> 5); }

struct RlPair {
    const double *rc1, *dm1, *z1, *w1, *f1z, *dl1;
    const double *rc2, *dm2, *z2, *w2, *f2z, *dl2;
    int n1, n2, order1, order2;
    double ch, sh, zq1, zq2, sw1, swsll1, sw2, swsll2;
    bool zerr_on, shp, windows;
};

// ------------------------------------------------------------------------------------------
// S: one sweep.  ROW: thread = pixel of forest 1 walking forest 2 (records RlRun1, counts, bin
// marks); !ROW: thread = pixel of forest 2 walking forest 1 (records RlRun2).
// ------------------------------------------------------------------------------------------
template <bool ROW>
__device__ __forceinline__ void rl_sweep(const pb2_params &P, const DmatFast &F, const RlPair &D,
                                         void *__restrict__ runs, int cap, int *__restrict__ nruns,
                                         int *__restrict__ rowcnt, int *__restrict__ kidx,
                                         int *__restrict__ aidx, int &cnt_nc, int &cnt_in)
{
    const int tid = threadIdx.x;
    const int nf = ROW ? D.n1 : D.n2, ns = ROW ? D.n2 : D.n1;
    const double *__restrict__ rcf = ROW ? D.rc1 : D.rc2, *__restrict__ dmf = ROW ? D.dm1 : D.dm2;
    const double *__restrict__ zf_ = ROW ? D.z1 : D.z2, *__restrict__ wf = ROW ? D.w1 : D.w2;
    const double *__restrict__ fzf = ROW ? D.f1z : D.f2z;
    const double *__restrict__ rcs = ROW ? D.rc2 : D.rc1, *__restrict__ dms = ROW ? D.dm2 : D.dm1;
    const double *__restrict__ zs = ROW ? D.z2 : D.z1, *__restrict__ ws = ROW ? D.w2 : D.w1;
    const double *__restrict__ fzs = ROW ? D.f2z : D.f1z, *__restrict__ dls = ROW ? D.dl2 : D.dl1;
    for (int fb = 0; fb < nf; fb += RL_THREADS) {
        const int f = min(fb + tid, nf - 1);
        const bool live = (fb + tid < nf) && (wf[f] != 0.);
        const double rc_f = rcf[f], dm_f = dmf[f], z_f = zf_[f], w_f = wf[f], fz_f = fzf[f];
        // zerr cut of the fixed pixel against the OTHER quasar (cf.py:629-636, :650-658)
        bool f_sel = true;
        if (D.zerr_on && pb2_zerr_close(P, z_f, ROW ? D.zq2 : D.zq1)) f_sel = false;
        int lo = 0, hi = -1;
        if (live) {
            if (ROW) row_window(P, D.windows, rc_f, dm_f, rcs, dms, ns, D.ch, D.sh, P.x_correlation, lo, hi);
            else col_window(P, D.windows, rc_f, dm_f, rcs, dms, ns, D.ch, D.sh, P.x_correlation, lo, hi);
        }
        int cA = -1, cB = -1, cn = 0, nrun = 0, row_in = 0;
        bool cS = false;
        double ea = 0., eb = 0., qa = 0., qb = 0., srp = 0., srt = 0., sz = 0., dg = 0.;
        // (a warp-uniform number of iterations with a __syncwarp() on top: the rows of a warp have
        // different windows and flush points and would otherwise run the loop a few lanes at a time)
        const int len = hi - lo + 1, maxlen = warp_max(len);
        // the swept pixel of the NEXT step is loaded one step ahead (the loop is latency-bound at
        // 16 warps per SM: without it every step waits for its own loads)
        const int last = ns - 1;
        double n_rc = rcs[min(lo, last)], n_dm = dms[min(lo, last)], n_w = ws[min(lo, last)];
        for (int t = 0; t < maxlen; t++) {
            __syncwarp();
            if (t >= len) continue;
            const int s = lo + t;
            const double s_rc = n_rc, s_dm = n_dm, s_w = n_w;
            {
                const int sn = min(s + 1, last);
                n_rc = rcs[sn];
                n_dm = dms[sn];
                n_w = ws[sn];
            }
            DmatGeom g;
            g.in = false;
            bool sel = false;
            double z = 0.;
            if (s < hi && s_w != 0.) {
                g = ROW ? dmat_pair(P, F, rc_f, dm_f, s_rc, s_dm, D.ch, D.sh, false, D.shp)
                        : dmat_pair(P, F, s_rc, s_dm, rc_f, dm_f, D.ch, D.sh, false, D.shp);
                if (g.in) {
                    z = mul_rn(add_rn(z_f, zs[s]), 0.5);   // == (z1 + z2) / 2 exactly (cf.py:690)
                    sel = f_sel && !g.close;
                    if (sel && ((P.has_z_min_pairs && z < P.z_min_pairs) ||
                                (P.has_z_max_pairs && z > P.z_max_pairs))) sel = false;
                    if (sel && D.zerr_on && pb2_zerr_close(P, zs[s], ROW ? D.zq1 : D.zq2)) sel = false;
                    if (ROW) {
                        row_in++;
                        if (!g.close) cnt_nc++;
                    }
                }
            }
            const bool brk = (s == hi) || (g.in && (g.A != cA || g.B != cB || sel != cS));
            if (brk && cB >= 0) {   // the finished run
                if (ROW) {
                    RlRun1 r;
                    r.A = cA; r.Bs = cB | (cS ? RL_SEL : 0); r.n = cn; r.pad = 0;
                    r.e1 = ea; r.e3 = eb; r.q1 = w_f * qa; r.q1d = w_f * qb;
                    r.rp = srp; r.rt = srt; r.zz = sz; r.dg = dg;
                    reinterpret_cast<RlRun1 *>(runs)[(long long)f * cap + nrun] = r;
                    kidx[cB] = 0;   // every in-range pair marks its model bin (cf.py:702)
                } else {
                    RlRun2 r;
                    r.A = cA; r.Bs = cB | (cS ? RL_SEL : 0);
                    r.e2 = ea; r.e4 = eb; r.q2 = w_f * qa; r.q2d = w_f * qb;
                    reinterpret_cast<RlRun2 *>(runs)[(long long)f * cap + nrun] = r;
                }
                if (ROW && cS) aidx[cA] = 0;
                nrun++;
                ea = eb = qa = qb = srp = srt = sz = dg = 0.;
                cn = 0;
                cB = -1;
            }
            if (s == hi || !g.in) continue;
            cA = g.A;
            cB = g.B;
            cS = sel;
            cn++;
            const double wj = s_w, dlj = dls[s];
            const double zf = mul_rn(fz_f, fzs[s]);
            ea += zf * wj;            // cf.py:767 / :771
            eb += zf * wj * dlj;      // cf.py:782-787 / :808-813
            if (sel) {
                qa += wj;
                qb += wj * dlj;
                if (ROW) {
                    const double w12 = mul_rn(w_f, wj);
                    dg += w12 * zf;       // cf.py:873
                    srp += w12 * g.rp;    // cf.py:714-717
                    srt += w12 * g.rt;
                    sz += w12 * z;
                }
            }
        }
        if (fb + tid < nf) {
            nruns[f] = nrun;
            if (ROW) rowcnt[f] = row_in;
        }
        if (ROW) cnt_in += row_in;
    }
}

// rank-(2 RL_ROWS) update of the local tile: thread (ty, tx) accumulates the local data bins
// ty + 16 p (p < np_w) x the local model bins tx + 32 q (q < NQ) over the rows of the group and
// subtracts the result from the window.  NQ is a template parameter and np_w warp-uniform: real
// branches, no predicated multiplications by zero.
template <int NQ>
__device__ __forceinline__ void rl_rank_update(RlShared &S, int np_w, int nA, int nB)
{
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    double c[8][NQ];
#pragma unroll
    for (int p = 0; p < 8; p++)
#pragma unroll
        for (int q = 0; q < NQ; q++) c[p][q] = 0.;
#pragma unroll 2
    for (int r = 0; r < RL_ROWS; r++) {
#pragma unroll
        for (int kind = 0; kind < 2; kind++) {
            double y[NQ];
            if (NQ >= 3) {
                const double4 ya = *reinterpret_cast<const double4 *>(&S.Yl[kind][r][tx * 4]);
                y[0] = ya.x; y[1] = ya.y; y[2] = ya.z;
                if (NQ == 4) y[NQ - 1] = ya.w;
            } else {
                const double2 ya = *reinterpret_cast<const double2 *>(&S.Yl[kind][r][tx * 4]);
                y[0] = ya.x;
                if (NQ == 2) y[NQ - 1] = ya.y;
            }
            const double *xr = &S.Xl[kind][r][ty * 8];
#pragma unroll
            for (int p = 0; p < 8; p++) {
                if (p >= np_w) break;
                const double x = xr[p];
#pragma unroll
                for (int q = 0; q < NQ; q++) c[p][q] = fma(x, y[q], c[p][q]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 8; p++) {
        if (p >= np_w) break;
        const int la = ty + 16 * p;
        if (la >= nA) continue;
        const int ka = S.lstA[la];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int lb = tx + 32 * q;
            if (lb < nB && c[p][q] != 0.) S.C[ka][S.lstB[lb]] -= c[p][q];
        }
    }
}

// double-precision add into shared memory (compare-and-swap; used where many threads issue
// independent adds, so the latency of the loop is hidden)
__device__ __forceinline__ void rl_sadd(double *p, double v) { atomicAdd(p, v); }

__global__ void __launch_bounds__(RL_THREADS, 1)
pb2_dmat_auto_run_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, DmatWork W,
                         int cap1, int cap2, double *__restrict__ weights_dmat,
                         double *__restrict__ dmat, double *__restrict__ r_par_eff,
                         double *__restrict__ r_trans_eff, double *__restrict__ z_eff,
                         double *__restrict__ weight_eff)
{
    extern __shared__ __align__(16) unsigned char rl_smem[];
    RlShared &S = *reinterpret_cast<RlShared *>(rl_smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const int nbm = P.num_model_bins_r_par * P.num_model_bins_r_trans;
    const double zerr_ang = mul_rn(P.zerr_cut_deg, PB2_PI) / 180.0;

    // per-CTA scratch (global, L2-resident)
    char *base = W.cta_base + (long long)blockIdx.x * W.cta_stride;
    int *kidx = (int *)base;                     // [nbm] compact model index, -1 = untouched
    int *aidx = kidx + nbm;                      // [nb]
    int *klist = aidx + nb;                      // [nbm]
    int *alist = klist + nbm;                    // [nb]
    int *nrun1 = alist + nb;                     // [max_pix1] runs per row
    int *nrun2 = nrun1 + c1.max_pix + 1;         // [max_pix2] runs per column
    int *rowcnt = nrun2 + c2.max_pix + 1;        // [max_pix1] in-range pairs per row
    RlRun1 *R1 = (RlRun1 *)(((uintptr_t)(rowcnt + c1.max_pix + 1) + 15) & ~(uintptr_t)15);
    RlRun2 *R2 = (RlRun2 *)(R1 + (long long)c1.max_pix * cap1);
    double *gvec = (double *)(R2 + (long long)c2.max_pix * cap2);   // vectors of pairs with > RL_VEC bins

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

        RlPair D;
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

        // ---------------- S: row sweep (counts, touched bins, run lists of forest 1)
        for (int x = tid; x < nbm; x += RL_THREADS) kidx[x] = -1;
        for (int x = tid; x < nb; x += RL_THREADS) aidx[x] = -1;
        __syncthreads();
        {
            int cnt_nc = 0, cnt_in = 0;
            rl_sweep<true>(P, W.fast, D, R1, cap1, nrun1, rowcnt, kidx, aidx, cnt_nc, cnt_in);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                cnt_nc += __shfl_xor_sync(0xffffffffu, cnt_nc, m);
                cnt_in += __shfl_xor_sync(0xffffffffu, cnt_in, m);
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
            // (cf.py:846-848).  Re-mark the model bins from the run lists with that rank limit.
            const int limit = S.cnt[0];
            for (int x = tid; x < nbm; x += RL_THREADS) kidx[x] = -1;
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
            for (int i = tid; i < n1; i += RL_THREADS) {
                int rank = rowcnt[i];
                const RlRun1 *rr = R1 + (long long)i * cap1;
                for (int k = 0; k < nrun1[i] && rank < limit; k++) {
                    kidx[rr[k].Bs & ~RL_SEL] = 0;   // the run's first pair has rank < limit
                    rank += rr[k].n;
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
        // ---------------- S: column sweep (run lists of forest 2); warps 0 / 1 join after the
        // compaction, which only touches kidx / aidx (not read by this sweep)
        {
            int d0 = 0, d1 = 0;
            rl_sweep<false>(P, W.fast, D, R2, cap2, nrun2, nullptr, nullptr, nullptr, d0, d1);
        }
        __syncthreads();
        const int U = S.U, UA = S.UA;
        if (tid == 0 && W.stats) {
            atomicAdd(W.stats, (double)S.cnt[0] * (15. * U + 4.) + 40. * (double)S.cnt[1]);
            atomicAdd(W.stats + 1, (double)U);
            atomicAdd(W.stats + 2, (double)S.cnt[1]);
        }

        // ---------------- V: per-bin vectors from the row runs
        const bool vec_smem = U <= RL_VEC && UA <= RL_VEC;
        const int vstride = vec_smem ? RL_VEC : max(nb, nbm);
        double *vP = vec_smem ? &S.vP[0][0] : gvec;
        double *vE = vP + 4 * (long long)vstride;
        double *vO = vE + 4 * (long long)vstride;
        for (int x = tid; x < 13 * vstride; x += RL_THREADS) vP[x] = 0.;
        __syncthreads();
        for (int i = warp; i < n1; i += RL_WARPS) {
            const int nr = nrun1[i];
            if (nr == 0) continue;
            const double dli = D.dl1[i], wi = D.w1[i];
            const double a5 = wi / D.sw1, a7 = wi * dli / D.swsll1;
            const double fe1 = 1. / D.sw2, fe3 = 1. / D.swsll2;   // eta1 / eta3 normalisations
            const RlRun1 *rr = R1 + (long long)i * cap1;
            RlRun1 *rr_w = R1 + (long long)i * cap1;
            for (int k = lane; k < nr; k += 32) {
                const RlRun1 r = rr[k];
                const int B = r.Bs & ~RL_SEL;
                const int kb = kidx[B];
                // from here on the record carries COMPACT indices: data bin (-1 when the run is
                // not selected) and model bin (-1 when outside the truncated np.unique, Q8)
                rr_w[k].A = (r.Bs & RL_SEL) ? aidx[r.A] : -1;
                rr_w[k].Bs = kb;
                if (kb >= 0) {   // (< 0: outside the truncated np.unique, Q8)
                    const double e1 = r.e1 * fe1, e3 = D.order2 == 1 ? r.e3 * fe3 : 0.;
                    rl_sadd(vE + 0 * vstride + kb, a5 * e1);                          // cf.py:775
                    if (D.order2 == 1) rl_sadd(vE + 1 * vstride + kb, a5 * e3);       // :793-802
                    if (D.order1 == 1) rl_sadd(vE + 2 * vstride + kb, a7 * e1);       // :818-827
                    if (D.order1 == 1 && D.order2 == 1) rl_sadd(vE + 3 * vstride + kb, a7 * e3);
                }
                if (r.Bs & RL_SEL) {
                    const int ka = aidx[r.A];
                    rl_sadd(vP + 0 * vstride + ka, r.q1);
                    rl_sadd(vP + 1 * vstride + ka, r.q1d);
                    rl_sadd(vP + 2 * vstride + ka, dli * r.q1);
                    rl_sadd(vP + 3 * vstride + ka, dli * r.q1d);
                    if (kb >= 0) {
                        rl_sadd(vO + 0 * vstride + kb, r.q1);   // weight_eff: sum of w12 (cf.py:717)
                        rl_sadd(vO + 1 * vstride + kb, r.rp);
                        rl_sadd(vO + 2 * vstride + kb, r.rt);
                        rl_sadd(vO + 3 * vstride + kb, r.zz);
                        if (W.fast.same) rl_sadd(vO + 4 * vstride + kb, r.dg);
                        else atomic_add_f64(dmat + (long long)r.A * nbm + B, r.dg);
                    } else {
                        atomic_add_f64(weight_eff + B, r.q1);
                        atomic_add_f64(r_par_eff + B, r.rp);
                        atomic_add_f64(r_trans_eff + B, r.rt);
                        atomic_add_f64(z_eff + B, r.zz);
                        atomic_add_f64(dmat + (long long)r.A * nbm + B, r.dg);
                    }
                }
            }
        }
        for (int j = warp; j < n2; j += RL_WARPS) {   // the same rewrite for the column runs
            const int nr = nrun2[j];
            RlRun2 *rr2 = R2 + (long long)j * cap2;
            for (int k = lane; k < nr; k += 32) {
                const int A = rr2[k].A, Bs = rr2[k].Bs;
                rr2[k].A = (Bs & RL_SEL) ? aidx[A] : -1;
                rr2[k].Bs = kidx[Bs & ~RL_SEL];
            }
        }
        __syncthreads();
        // flush of the per-pair sums that are complete now
        for (int ka = tid; ka < UA; ka += RL_THREADS)
            if (vP[ka] != 0.) atomic_add_f64(weights_dmat + alist[ka], vP[ka]);   // cf.py:718
        for (int kb = tid; kb < U; kb += RL_THREADS) {
            const double we = vO[kb];
            if (we != 0.) {
                const int Bm = klist[kb];
                atomic_add_f64(weight_eff + Bm, we);
                atomic_add_f64(r_par_eff + Bm, vO[1 * vstride + kb]);
                atomic_add_f64(r_trans_eff + Bm, vO[2 * vstride + kb]);
                atomic_add_f64(z_eff + Bm, vO[3 * vstride + kb]);
                if (W.fast.same) atomic_add_f64(dmat + (long long)Bm * nbm + Bm, vO[4 * vstride + kb]);
            }
        }

        // ---------------- C: contraction, one window of RL_W x RL_W compact bins at a time
        const int ty = tid >> 5, tx = tid & 31;
        for (int kc = 0; kc < U; kc += RL_W) {
            const int Uc = min(RL_W, U - kc);
            for (int ac = 0; ac < UA; ac += RL_W) {
                const int UAc = min(RL_W, UA - ac);
                __syncthreads();
                for (int x = tid; x < RL_W * RL_W; x += RL_THREADS) (&S.C[0][0])[x] = 0.;
                // rows of forest 1 (side 0), then columns of forest 2 (side 1)
                for (int side = 0; side < 2; side++) {
                    const int nrows = side == 0 ? n1 : n2;
                    const int *nrun = side == 0 ? nrun1 : nrun2;
                    for (int g0 = 0; g0 < nrows; g0 += RL_ROWS) {
                        __syncthreads();
                        if (tid < RL_W) {
                            S.flgA[tid] = 0;
                            S.flgB[tid] = 0;
                        }
                        for (int x = tid; x < 2 * RL_ROWS * RL_LOC; x += RL_THREADS) {
                            (&S.Xl[0][0][0])[x] = 0.;
                            (&S.Yl[0][0][0])[x] = 0.;
                        }
                        __syncthreads();
                        // -- the bins of the window this group touches (warp = row)
                        const int i = g0 + warp;
                        const int nr = i < nrows ? nrun[i] : 0;
                        const char *rbase = side == 0 ? (const char *)(R1 + (long long)i * cap1)
                                                      : (const char *)(R2 + (long long)i * cap2);
                        const int rsz = side == 0 ? (int)sizeof(RlRun1) : (int)sizeof(RlRun2);
                        // (the first 64 runs of the row stay in registers for the expansion below)
                        int2 cab[2];
                        double cv[2][4];
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int k = lane + 32 * h;
                            cab[h] = make_int2(-1, -1);
                            if (k < nr) {
                                const char *rec = rbase + (long long)k * rsz;
                                cab[h] = *reinterpret_cast<const int2 *>(rec);
                                const double *v = reinterpret_cast<const double *>(rec + (side == 0 ? 16 : 8));
                                cv[h][0] = v[0]; cv[h][1] = v[1]; cv[h][2] = v[2]; cv[h][3] = v[3];
                                const int ka = cab[h].x - ac, kb = cab[h].y - kc;
                                if (cab[h].y >= 0 && kb >= 0 && kb < Uc) S.flgB[kb] = 1;
                                if (cab[h].x >= 0 && ka >= 0 && ka < UAc) S.flgA[ka] = 1;
                            }
                        }
                        for (int k = lane + 64; k < nr; k += 32) {
                            const int2 ab = *reinterpret_cast<const int2 *>(rbase + (long long)k * rsz);
                            const int ka = ab.x - ac, kb = ab.y - kc;
                            if (ab.y >= 0 && kb >= 0 && kb < Uc) S.flgB[kb] = 1;
                            if (ab.x >= 0 && ka >= 0 && ka < UAc) S.flgA[ka] = 1;
                        }
                        __syncthreads();
                        if (warp < 2) {   // local indices: warp 0 the data bins, warp 1 the model bins
                            const unsigned char *flg = warp == 0 ? S.flgA : S.flgB;
                            short *loc = warp == 0 ? S.locA : S.locB;
                            short *lst = warp == 0 ? S.lstA : S.lstB;
                            int u = 0;
                            for (int xb = 0; xb < RL_W; xb += 32) {
                                const bool on = flg[xb + lane] != 0;
                                const unsigned m = __ballot_sync(0xffffffffu, on);
                                const int k = u + __popc(m & ((1u << lane) - 1u));
                                loc[xb + lane] = on ? (short)k : (short)-1;
                                if (on) lst[k] = (short)(xb + lane);
                                u += __popc(m);
                            }
                            if (lane == 0) {
                                if (warp == 0) S.nA = u; else S.nB = u;
                            }
                        }
                        __syncthreads();
                        const int nA = S.nA, nB = S.nB;
                        if (nA == 0 || nB == 0) continue;
                        // normalisations of the eta rows (cf.py:767-813)
                        const double fa = side == 0 ? 1. / D.sw2 : 1. / D.sw1;
                        const double fb3 = side == 0 ? (D.order2 == 1 ? 1. / D.swsll2 : 0.)
                                                     : (D.order1 == 1 ? 1. / D.swsll1 : 0.);
                        // -- expand the runs of the group
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int2 ab = cab[h];
                            const int ka = ab.x - ac, kb = ab.y - kc;
                            if (ab.y >= 0 && kb >= 0 && kb < Uc) {
                                const int p = rl_ypos(S.locB[kb]);
                                rl_sadd(&S.Yl[0][warp][p], cv[h][0] * fa);
                                if (fb3 != 0.) rl_sadd(&S.Yl[1][warp][p], cv[h][1] * fb3);
                            }
                            if (ab.x >= 0 && ka >= 0 && ka < UAc) {
                                const int p = rl_xpos(S.locA[ka]);
                                rl_sadd(&S.Xl[0][warp][p], cv[h][2]);
                                rl_sadd(&S.Xl[1][warp][p], cv[h][3]);
                            }
                        }
                        for (int k = lane + 64; k < nr; k += 32) {
                            const char *rec = rbase + (long long)k * rsz;
                            const int2 ab = *reinterpret_cast<const int2 *>(rec);
                            const double *v = reinterpret_cast<const double *>(rec + (side == 0 ? 16 : 8));
                            const int ka = ab.x - ac, kb = ab.y - kc;
                            if (ab.y >= 0 && kb >= 0 && kb < Uc) {
                                const int p = rl_ypos(S.locB[kb]);
                                rl_sadd(&S.Yl[0][warp][p], v[0] * fa);
                                if (fb3 != 0.) rl_sadd(&S.Yl[1][warp][p], v[1] * fb3);
                            }
                            if (ab.x >= 0 && ka >= 0 && ka < UAc) {
                                const int p = rl_xpos(S.locA[ka]);
                                rl_sadd(&S.Xl[0][warp][p], v[2]);
                                rl_sadd(&S.Xl[1][warp][p], v[3]);
                            }
                        }
                        __syncthreads();
                        // -- rank-32 update of the local tile in registers, added into the window
                        {
                            const int np_w = (nA + 15) >> 4, nq_w = (nB + 31) >> 5;
                            if (ty < nA) {
                                if (nq_w == 1) rl_rank_update<1>(S, np_w, nA, nB);
                                else if (nq_w == 2) rl_rank_update<2>(S, np_w, nA, nB);
                                else if (nq_w == 3) rl_rank_update<3>(S, np_w, nA, nB);
                                else rl_rank_update<4>(S, np_w, nA, nB);
                            }
                        }
                    }
                }
                __syncthreads();
                // ---------------- F: rank-1 terms and the scatter of the window
                for (int x = tid; x < UAc * Uc; x += RL_THREADS) {
                    const int ka = x / Uc, kb = x - ka * Uc;
                    const int ga = ac + ka, gb = kc + kb;
                    const double v = S.C[ka][kb] + vP[0 * vstride + ga] * vE[0 * vstride + gb] +
                                     vP[1 * vstride + ga] * vE[1 * vstride + gb] +
                                     vP[2 * vstride + ga] * vE[2 * vstride + gb] +
                                     vP[3 * vstride + ga] * vE[3 * vstride + gb];
                    if (v != 0.) atomic_add_f64(dmat + (long long)alist[ga] * nbm + klist[gb], v);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Runs of one row: at most one per pixel of the other forest; and, when the forests are sorted,
// r_par is monotone along the row (V-shaped after abs()) and r_trans increasing, so the key
// (data bin, model bin, selected) changes at most 2 np + nt + 2 npm + ntm times plus a few flips
// of `selected` (z-pair cut, zerr cut, half-plate cut: each an interval of the row).
static int rl_cap(int n_other, const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par)
{
    long long cap = (long long)n_other + 2;
    if (c1->sorted && c2->sorted) {
        const long long geo = 2ll * (par->num_bins_r_par + par->num_model_bins_r_par) +
                              par->num_bins_r_trans + par->num_model_bins_r_trans + 16;
        if (geo < cap) cap = geo;
    }
    return (int)cap;
}

long long pb2_dmat_run_cta_bytes(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par)
{
    const long long nb = (long long)par->num_bins_r_par * par->num_bins_r_trans;
    const long long nbm = (long long)par->num_model_bins_r_par * par->num_model_bins_r_trans;
    long long bytes = (2 * nb + 2 * nbm) * 4 + (2ll * c1->max_pix + c2->max_pix + 3) * 4 + 64;
    bytes += (long long)c1->max_pix * rl_cap(c2->max_pix, c1, c2, par) * (long long)sizeof(RlRun1);
    bytes += (long long)c2->max_pix * rl_cap(c1->max_pix, c1, c2, par) * (long long)sizeof(RlRun2);
    bytes += 13 * (nb > nbm ? nb : nbm) * 8;
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
    const size_t smem = sizeof(RlShared);
    PB2_CUDA(cudaFuncSetAttribute(pb2_dmat_auto_run_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pb2_dmat_auto_run_kernel<<<blocks, RL_THREADS, smem, s>>>(
        *cat1, *cat2, *par, *pairs, W, rl_cap(cat2->max_pix, cat1, cat2, par),
        rl_cap(cat1->max_pix, cat1, cat2, par),
        d_weights_dmat, d_dmat, d_r_par_eff, d_r_trans_eff, d_z_eff, d_weight_eff);
    return 0;
}
