// Auto / delta x delta pixel-pair histogram, standard (r_par, r_trans) binning, no per-pair cuts:
// "diagonal lanes" kernel.  Replaces cf.compute_xi's pair loop + cf.compute_xi_forest_pairs_fast
// (reference py/picca/cf.py:161-240, 250-387).
//
// For one forest pair the pixel pairs (i, j) are visited along diagonals d = j - i.  On a diagonal
// r_par = (rc1[i] - rc2[j]) cos(ang/2) is nearly constant and r_trans = (dm1[i] + dm2[j]) sin(ang/2)
// grows slowly, so a diagonal stays in ONE (r_par, r_trans) bin for ~100 consecutive pairs.  Each
// lane owns two adjacent diagonals and keeps, per diagonal, the current bin ("run"), its partial
// sums and three thresholds bounding the run in registers:
//   per pair   d = rc1 - rc2, t = dm1 + dm2, three compares against the thresholds and seven
//              accumulate instructions -- 12 FP64 instructions, no division, no bin arithmetic;
//   run change (about once per 100 pairs per diagonal) the lane flushes the finished run with
//              native red.global.add.f64 and finds the new bin with the sandwich test of
//              pb2_xi.cu (reference expression with true divisions when it cannot prove the bin);
//              the new thresholds are the bin's edges mapped to d and t, shrunk by a 1e-12 guard
//              band -- a pair inside the guard band becomes a one-value run, so every bin
//              assignment is either proven or computed by the reference expression: bit-exact.
// Work unit = (forest pair, block of 64 diagonals), one warp.  The warp walks the rows; the row's
// five values are uniform loads; each lane fetches ONE new column element per step (the other
// slides in a register) from an interleaved-by-2 copy of forest 2, which makes the load coalesced.
#include "pb2_common.cuh"

#ifndef DG_THREADS
#define DG_THREADS 512
#endif
#define DG_WARPS (DG_THREADS / 32)
#ifndef DG_CHUNK
#define DG_CHUNK 8
#endif
#define DG_BLOCK 64
#define DG_EPS 1e-12
#define DG_DEAD_RC (-1e300)

struct DiagConst {
    double dbin_p;   // (r_par_max - r_par_min) / np
    double dbin_t;   // r_trans_max / nt
    double rp_scale; // max(|r_par_min|, |r_par_max|)
    double kp_lo, kp_hi, kt_lo, kt_hi, magic;  // sandwich constants (see pb2_xi.cu)
    int gmax;        // diagonal blocks per forest pair (longest forests)
};

__device__ __forceinline__ double dg_next_up(double v)
{
    if (v == 0.) return 4.9406564584124654e-324;
    const long long b = __double_as_longlong(v);
    return __longlong_as_double(v > 0. ? b + 1 : b - 1);
}

__device__ __forceinline__ int dg_exact_bin(const pb2_params &P, double rc1, double dm1, double rc2,
                                         double dm2, double ang, double ch, double sh)
{
    return pb2_pair_exact(P, rc1, dm1, rc2, dm2, ang, ch, sh, false, false).bin;
}

// first index in non-decreasing a[0..n) with a[idx] > v (strict) or a[idx] >= v
__device__ __forceinline__ int dg_bound(const double *__restrict__ a, int n, double v, bool strict)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double x = __ldg(a + mid);
        if (strict ? (x <= v) : (x < v)) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// ---- per-diagonal state on named scalars (r = 0, 1)
// The five sums and the count are RUNNING totals over the whole diagonal (never cleared, also
// advanced while the diagonal is in a dead run); a run's contribution is total - snapshot, the
// snapshot being taken when the run starts.  This keeps the per-pair path free of predication
// and the run-change path free of writes to the accumulators.
#define DG_DECL(r)                                                                       \
    double lo_##r = 1e299, hi_##r = inf, thi_##r = inf; /* dead run: column outside */   \
    double sw_##r = 0., sxi_##r = 0., srp_##r = 0., srt_##r = 0., sz_##r = 0.;           \
    int bp_##r = -1, bt_##r = 0, cnt_##r = 0; /* bp < 0: dead run */                     \
    snap[r][0][lane] = snap[r][1][lane] = snap[r][2][lane] = snap[r][3][lane] =          \
        snap[r][4][lane] = 0.;                                                           \
    qcnt[r][lane] = 0;

#define DG_COLS(c)  double c##_rc, c##_dm, c##_w, c##_dw, c##_z;

// one column element of forest 2 (interleaved layout), `pos` = slot, `jj` = its pixel index
#define DG_LOAD(c, pos, jj)                                                              \
    {                                                                                    \
        c##_rc = DG_DEAD_RC;                                                             \
        c##_dm = c##_w = c##_dw = c##_z = 0.;                                            \
        if ((jj) >= 0 && (jj) < n2) {                                                    \
            const double2 a2 = __ldg(p_rcdm2 + (pos));                                   \
            const double2 b2 = __ldg(p_wdw2 + (pos));                                    \
            c##_z = __ldg(p_z2 + (pos));                                                 \
            c##_rc = a2.x;                                                               \
            c##_dm = a2.y;                                                               \
            c##_w = b2.x;                                                                \
            c##_dw = b2.y;                                                               \
        }                                                                                \
    }

#define DG_TEST(r, c)                                                                    \
    const double d_##r = sub_rn(rc1, c##_rc);                                            \
    const double v_##r = XCORR ? d_##r : fabs(d_##r);                                    \
    const double t_##r = add_rn(dm1, c##_dm);                                            \
    const bool p_##r = (v_##r >= lo_##r) && (v_##r < hi_##r) && (t_##r < thi_##r);

// add the finished run (totals - snapshot) to its bin and take the snapshot for the next run;
// snapshots live in lane-private shared-memory slots (only this rare path touches them)
#define DG_RED(r)                                                                        \
    {                                                                                    \
        const int dc = cnt_##r - qcnt[r][lane];                                          \
        if (bp_##r >= 0 && dc > 0) {                                                     \
            double *const dst = orow + (bt_##r + nt_i * bp_##r);                         \
            atomic_add_f64(dst + 0 * (size_t)nb, sw_##r - snap[r][0][lane]);             \
            atomic_add_f64(dst + 1 * (size_t)nb, sxi_##r - snap[r][1][lane]);            \
            atomic_add_f64(dst + 2 * (size_t)nb, (srp_##r - snap[r][2][lane]) * edge[5]); \
            atomic_add_f64(dst + 3 * (size_t)nb, (srt_##r - snap[r][3][lane]) * edge[6]); \
            atomic_add_f64(dst + 4 * (size_t)nb, 0.5 * (sz_##r - snap[r][4][lane]));     \
            atomic_add_i64(dst + 5 * (size_t)nb, (long long)dc);                         \
        }                                                                                \
        snap[r][0][lane] = sw_##r;                                                       \
        snap[r][1][lane] = sxi_##r;                                                      \
        snap[r][2][lane] = srp_##r;                                                      \
        snap[r][3][lane] = srt_##r;                                                      \
        snap[r][4][lane] = sz_##r;                                                       \
        qcnt[r][lane] = cnt_##r;                                                         \
    }

// The pair left its run.  Common case, handled first: a live run steps into the ADJACENT r_par
// bin (or the next r_trans bin) and the pair lies outside the guard band of the new bin, which
// proves the new bin without evaluating it.  Otherwise: sandwich test / reference expression.
#define DG_LEAVE(r, c)                                                                   \
    if (!p_##r) {                                                                        \
        double nlo = lo_##r, nhi = hi_##r, nthi = thi_##r;                               \
        int nbp = bp_##r, nbt = bt_##r;                                                  \
        bool done = false;                                                               \
        const double e0 = edge[0], ep = edge[1], et = edge[2];                           \
        const double abs_p = edge[3], abs_t = edge[4];                                   \
        if (bp_##r >= 0) {                                                               \
            const bool up = v_##r >= hi_##r, dn = v_##r < lo_##r, tup = t_##r >= thi_##r; \
            if (!tup && (up != dn)) { /* adjacent r_par bin */                           \
                const int cb = bp_##r + (up ? 1 : -1);                                   \
                const double e_lo = fma((double)cb, ep, e0), e_hi = e_lo + ep;           \
                double a_lo = e_lo + fma(fabs(e_lo), DG_EPS, abs_p);                     \
                const double a_hi = e_hi - fma(fabs(e_hi), DG_EPS, abs_p);               \
                if (!XCORR && e_lo <= 0.) a_lo = -1.;                                    \
                if (cb >= 0 && cb < np_i && v_##r >= a_lo && v_##r < a_hi) {             \
                    nbp = cb;                                                            \
                    nlo = a_lo;                                                          \
                    nhi = a_hi;                                                          \
                    done = true;                                                         \
                }                                                                        \
            } else if (tup && !up && !dn) { /* next r_trans bin */                       \
                const double t_lo = (double)(bt_##r + 1) * et, t_hi = t_lo + et;         \
                const double b_lo = t_lo + fma(t_lo, DG_EPS, abs_t);                     \
                const double b_hi = t_hi - fma(t_hi, DG_EPS, abs_t);                     \
                if (bt_##r + 1 < nt_i && t_##r >= b_lo && t_##r < b_hi) {                \
                    nbt = bt_##r + 1;                                                    \
                    nthi = b_hi;                                                         \
                    done = true;                                                         \
                }                                                                        \
            }                                                                            \
        }                                                                                \
        if (!done) {                                                                     \
            nlo = -inf;                                                                  \
            nhi = inf;                                                                   \
            nthi = inf;                                                                  \
            nbp = -1;                                                                    \
            nbt = 0;                                                                     \
            if (c##_rc == DG_DEAD_RC) {                                                  \
                nlo = 1e299;                                                             \
            } else {                                                                     \
                const double ch = edge[5], sh = edge[6];                                 \
                const double rp = mul_rn(v_##r, ch); /* |fl(d ch)| == fl(|d| ch) */      \
                const double rt = mul_rn(t_##r, sh);                                     \
                const double x = sub_rn(rp, P.r_par_min);                                \
                const int bpl = __double2loint(__fma_rd(x, C.kp_lo, C.magic));           \
                const int bph = __double2loint(__fma_rd(x, C.kp_hi, C.magic));           \
                const int btl = __double2loint(__fma_rd(rt, C.kt_lo, C.magic));          \
                const int bth = __double2loint(__fma_rd(rt, C.kt_hi, C.magic));          \
                if (bpl != bph || btl != bth || !(fabs(x) < 1e15) || !(rt < 1e15)) {     \
                    const int bin = dg_exact_bin(P, rc1, dm1, c##_rc, c##_dm, ang, ch, sh); \
                    if (bin >= 0) {                                                      \
                        nbp = bin / nt_i;                                                \
                        nbt = bin - nbp * nt_i;                                          \
                    }                                                                    \
                    nlo = v_##r;                                                         \
                    nhi = dg_next_up(v_##r);                                             \
                    nthi = dg_next_up(t_##r);                                            \
                } else if (btl < nt_i) { /* else r_trans >= max: dead for good */        \
                    const int bc = max(-1, min(bpl, np_i)); /* -1, np: rejected sides */ \
                    const double e_lo = fma((double)bc, ep, e0), e_hi = e_lo + ep;       \
                    if (bc >= 0) nlo = e_lo + fma(fabs(e_lo), DG_EPS, abs_p);            \
                    if (bc < np_i) nhi = e_hi - fma(fabs(e_hi), DG_EPS, abs_p);          \
                    if (!XCORR && e_lo <= 0.) nlo = -1.; /* |r_par| >= 0 always */       \
                    if (bc >= 0 && bc < np_i) {                                          \
                        const double t_hi = (double)(btl + 1) * et;                      \
                        nthi = isfinite(t_hi) ? t_hi - fma(t_hi, DG_EPS, abs_t) : inf;   \
                        if (!(t_##r < nthi)) nthi = dg_next_up(t_##r);                   \
                        nbp = bc;                                                        \
                        nbt = btl;                                                       \
                    }                                                                    \
                    if (!(v_##r >= nlo && v_##r < nhi)) { /* guard band: one-value run */ \
                        nlo = v_##r;                                                     \
                        nhi = dg_next_up(v_##r);                                         \
                    }                                                                    \
                }                                                                        \
            }                                                                            \
        }                                                                                \
        if (nbp != bp_##r || nbt != bt_##r) {                                            \
            DG_RED(r)                                                                    \
            bp_##r = nbp;                                                                \
            bt_##r = nbt;                                                                \
        }                                                                                \
        lo_##r = nlo;                                                                    \
        hi_##r = nhi;                                                                    \
        thi_##r = nthi;                                                                  \
    }

#define DG_ACC(r, c)                                                                     \
    {                                                                                    \
        const double w12 = mul_rn(w1, c##_w);                                            \
        sw_##r += w12;                                                                   \
        sxi_##r = fma(dw1, c##_dw, sxi_##r);                                             \
        srp_##r = fma(v_##r, w12, srp_##r);                                              \
        srt_##r = fma(t_##r, w12, srt_##r);                                              \
        sz_##r = fma(add_rn(z1, c##_z), w12, sz_##r);                                    \
        if (row_ok && __double2hiint(c##_w) != 0) cnt_##r += 1;                          \
    }

// row ii against the lane's two diagonals, which meet columns ca (d = D0 + 2 lane) and cb (+1)
#define DG_STEP(ii, ca, cb)                                                              \
    {                                                                                    \
        const double2 ra = __ldg(p_rcdm1 + (ii));                                        \
        const double2 rb = __ldg(p_wdw1 + (ii));                                         \
        const double z1 = __ldg(p_z1 + (ii));                                            \
        const double rc1 = ra.x, dm1 = ra.y, w1 = rb.x, dw1 = rb.y;                      \
        const bool row_ok = (w1 != 0.);                                                  \
        DG_TEST(0, ca)                                                                   \
        DG_TEST(1, cb)                                                                   \
        if (__any_sync(0xffffffffu, !(p_0 && p_1))) {                                    \
            DG_LEAVE(0, ca)                                                              \
            DG_LEAVE(1, cb)                                                              \
        }                                                                                \
        DG_ACC(0, ca)                                                                    \
        DG_ACC(1, cb)                                                                    \
    }

template <bool XCORR>
__global__ void __launch_bounds__(DG_THREADS, 1)
pb2_xi_auto_diag(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, DiagConst C,
                 const int32_t *__restrict__ out_row, double *__restrict__ out)
{
    __shared__ unsigned s_ctr;
    __shared__ double s_snap[DG_WARPS][2][5][32];
    __shared__ int s_qcnt[DG_WARPS][2][32];
    __shared__ double s_edge[DG_WARPS][8];
    if (threadIdx.x == 0) s_ctr = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double (*snap)[5][32] = s_snap[threadIdx.x >> 5];
    int (*qcnt)[32] = s_qcnt[threadIdx.x >> 5];
    double *edge = s_edge[threadIdx.x >> 5];
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const int np_i = P.num_bins_r_par, nt_i = P.num_bins_r_trans;
    const unsigned gmax = (unsigned)C.gmax;
    const unsigned units_per_chunk = DG_CHUNK * gmax;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);

    for (;;) {
        unsigned u = 0;
        if (lane == 0) u = atomicAdd(&s_ctr, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        const long long chunk = (long long)blockIdx.x + (long long)(u / units_per_chunk) * gridDim.x;
        if (chunk * DG_CHUNK >= pr.n_pairs) break;
        const unsigned local = u % units_per_chunk;
        const long long e = chunk * DG_CHUNK + local / gmax;
        if (e >= pr.n_pairs) continue;
        const int g = (int)(local % gmax);

        const int k = pr.nb_f1[e];
        const int f1 = pr.f1_index[k];
        const int f2 = pr.nb_f2[e];
        const long long a = c1.offset[f1];
        const int n1 = (int)(c1.offset[f1 + 1] - a);
        const long long b = c2.offset[f2];
        const int n2 = (int)(c2.offset[f2 + 1] - b);
        if (n1 == 0 || n2 == 0) continue;
        const double ch = pr.nb_cos[e], sh = pr.nb_sin[e], ang = pr.nb_ang[e];
        const double *__restrict__ p_rc1 = c1.r_comov + a;
        const double *__restrict__ p_dm1 = c1.dist_m + a;
        const double *__restrict__ p_z1 = c1.z + a;

        // ---- diagonal range of the forest pair and row range of this block (supersets).
        // lane = segment of L consecutive rows; columns of row i in range: [jlo(i), jhi(i))
        const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
        const double dmax = P.r_par_max * inv_c * (1. + 1e-9) + 1e-9;
        const double dmin = P.r_par_min * inv_c;
        const double dlow = XCORR ? (dmin - fabs(dmin) * 1e-9 - 1e-9) : -dmax;
        const double tsum = P.r_trans_max * inv_s * (1. + 1e-9) + 1e-9;
        const int L = (n1 + 31) >> 5;
        const int s0 = lane * L, s1 = min(n1, s0 + L) - 1;
        int dlo = 0x7fffffff, dhi = -0x7fffffff;
        if (s0 < n1) {
            const int jlo = dg_bound(c2.r_comov + b, n2, __ldg(p_rc1 + s0) - dmax, true);
            int jhi = dg_bound(c2.r_comov + b, n2, __ldg(p_rc1 + s1) - dlow, false);
            if (isfinite(tsum))
                jhi = min(jhi, dg_bound(c2.dist_m + b, n2, tsum - __ldg(p_dm1 + s0), false));
            if (jhi > jlo) {
                dlo = jlo - s1;
                dhi = jhi - 1 - s0;
            }
        }
        int Dmin = dlo, Dmax = dhi;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            Dmin = min(Dmin, __shfl_xor_sync(0xffffffffu, Dmin, m));
            Dmax = max(Dmax, __shfl_xor_sync(0xffffffffu, Dmax, m));
        }
        if (Dmin > Dmax) continue;
        const int D0 = Dmin + DG_BLOCK * g;
        if (D0 > Dmax) continue;
        const int D1 = min(D0 + DG_BLOCK - 1, Dmax);
        const unsigned segs = __ballot_sync(0xffffffffu, dlo <= D1 && dhi >= D0);
        if (!segs) continue;
        int ibeg = (__ffs(segs) - 1) * L;
        int iend = min(n1, (32 - __clz(segs)) * L);
        ibeg = max(ibeg, max(0, -D1));
        iend = min(iend, n2 - D0);
        if (ibeg >= iend) continue;

        const long long pb = c2.perm_offset[f2];
        const double2 *__restrict__ p_rcdm2 = reinterpret_cast<const double2 *>(c2.rcdm_p) + pb;
        const double2 *__restrict__ p_wdw2 = reinterpret_cast<const double2 *>(c2.wdw_p) + pb;
        const double *__restrict__ p_z2 = c2.z_p + pb;
        const double2 *__restrict__ p_rcdm1 = reinterpret_cast<const double2 *>(c1.rcdm) + a;
        const double2 *__restrict__ p_wdw1 = reinterpret_cast<const double2 *>(c1.wdw) + a;
        const int S2 = (n2 + 1) >> 1;
        double *__restrict__ orow = out + (size_t)out_row[k] * 6 * nb;
        __syncwarp();
        if (lane == 0) {  // bin edges in units of d and t, guard band, cos / sin of ang/2
            edge[0] = P.r_par_min * inv_c;
            edge[1] = C.dbin_p * inv_c;
            edge[2] = C.dbin_t * inv_s;
            edge[3] = DG_EPS * C.rp_scale * inv_c;
            edge[4] = DG_EPS * P.r_trans_max * inv_s;
            edge[5] = ch;
            edge[6] = sh;
        }
        DG_DECL(0)
        DG_DECL(1)
        __syncwarp();
        DG_COLS(c0)
        DG_COLS(c1)
        // c0 / c1 = columns i + D0 + 2 lane + {0, 1}; each step the window slides by one column.
        // i advances by 2 per iteration, so c0 always reads one parity half of the interleaved
        // copy and c1 the other: two slot counters that advance by one per iteration.
        const int jb = ibeg + D0;
        int pos0 = (jb & 1) * S2 + (jb >> 1) + lane;              // slot of column jb + 2 lane
        int pos1 = ((jb + 1) & 1) * S2 + ((jb + 1) >> 1) + lane;  // slot of column jb + 1 + 2 lane
        int jj0 = jb + 2 * lane;
        DG_LOAD(c0, pos0, jj0)
        DG_LOAD(c1, pos1, jj0 + 1)
        for (int i = ibeg; i < iend; i += 2) {
            DG_STEP(i, c0, c1)
            pos0 += 1;
            DG_LOAD(c0, pos0, jj0 + 2)
            if (i + 1 < iend) {
                DG_STEP(i + 1, c1, c0)
                pos1 += 1;
                DG_LOAD(c1, pos1, jj0 + 3)
            }
            jj0 += 2;
        }
        DG_RED(0)
        DG_RED(1)
    }
}

int32_t pb2_launch_xi_diag(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                           const pb2_pairs *pairs, const int32_t *d_out_row, double *d_out,
                           cudaStream_t s)
{
    DiagConst C;
    C.dbin_p = (par->r_par_max - par->r_par_min) / par->num_bins_r_par;
    C.dbin_t = par->r_trans_max / par->num_bins_r_trans;
    const double a0 = par->r_par_min < 0 ? -par->r_par_min : par->r_par_min;
    const double a1 = par->r_par_max < 0 ? -par->r_par_max : par->r_par_max;
    C.rp_scale = a0 > a1 ? a0 : a1;
    C.gmax = (c1->max_pix + c2->max_pix + DG_BLOCK - 1) / DG_BLOCK;
    if (C.gmax < 1) C.gmax = 1;
    const double kp = (double)par->num_bins_r_par / (par->r_par_max - par->r_par_min);
    const double kt = (double)par->num_bins_r_trans / par->r_trans_max;
    const double eps = 9.094947017729282e-13;  // 2^-40
    C.kp_lo = kp * (1. - eps);
    C.kp_hi = kp * (1. + eps);
    C.kt_lo = kt * (1. - eps);
    C.kt_hi = kt * (1. + eps);
    C.magic = 6755399441055744.0;  // 2^52 + 2^51
    int dev = 0, sms = 0;
    PB2_CUDA(cudaGetDevice(&dev));
    PB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long want = (pairs->n_pairs + DG_CHUNK - 1) / DG_CHUNK;
    int blocks = (int)(want < sms ? want : sms);
    if (blocks < 1) blocks = 1;
    if (par->x_correlation)
        pb2_xi_auto_diag<true><<<blocks, DG_THREADS, 0, s>>>(*c1, *c2, *par, *pairs, C, d_out_row, d_out);
    else
        pb2_xi_auto_diag<false><<<blocks, DG_THREADS, 0, s>>>(*c1, *c2, *par, *pairs, C, d_out_row, d_out);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_xi_auto_diag");
}
