// Auto / delta x delta pixel-pair histogram, standard (r_par, r_trans) binning, no per-pair cuts:
// "diagonal lanes" kernel -- the product path of picca_cf.py's default mode.  Replaces
// cf.compute_xi's pair loop + cf.compute_xi_forest_pairs_fast (reference py/picca/cf.py:161-240,
// 250-387).
//
// For one forest pair the pixel pairs (i, j) are visited along diagonals j - i = const.  On a
// diagonal r_par = (rc1[i] - rc2[j]) cos(ang/2) is nearly constant and r_trans = (dm1[i] + dm2[j])
// sin(ang/2) grows slowly, so a diagonal stays in ONE (r_par, r_trans) bin for ~80 consecutive
// pairs (a "run").  Each lane owns DG_C adjacent diagonals and keeps, per diagonal, the bin of the
// current run and five running totals in registers.
//   per pair    d = rc1 - rc2, t = dm1 + dm2, then the bin straight from four round-down FMAs
//               against 2^52 + 2^51 (floor(x K (1 -+ 2^-40)) in the low word): when the two r_par
//               values agree and the two r_trans values agree, the reference's
//               floor((r - min) / (max - min) * n) (cf.py:372-376) is sandwiched and the bin is
//               proven, and "0 <= value < n" is the reference's range test (cf.py:364).  Four
//               integer compares of the low words with the run's bin, seven FP64 instructions to
//               accumulate: 13 FP64 instructions, no division, no predication, no bounds check
//               and no pair counter -- zero-weight pixels (skipped by the reference,
//               cf.py:318,331) are compacted away in the packed copies, so num_pairs of a run is
//               its length, and columns outside forest 2 read dummies (distance 1e300, weight 0)
//               that add zeros and land in no bin.
//   run change  (about once per 40-80 pairs per diagonal) the lane adds its five sums and the run
//               length to the finished run's bin with six native red.global.add.f64 / .u64 into
//               the L2-resident per-HEALPix histogram (~30 instructions).  The sums restart through
//               six selects in the MAIN path (high word := 0 when the bin changed, which leaves
//               at most a 1e-314 denormal behind): the accumulators are then written only by the
//               accumulate instructions and ptxas keeps them in place -- clearing them inside the
//               branch costs ten register moves per pair, snapshots in shared memory saturate the
//               L1 data pipe.  When the two FMAs of a dimension disagree (pair within 2^-40 of a
//               bin edge, ~1e-10 of the pairs) the bin comes from the reference expression with
//               IEEE divisions.
// Work unit = (forest pair, block of 32 * DG_C diagonals), one warp.  The warp walks the rows: the
// row's values are uniform 128-bit loads; each lane loads ONE new column element per row (the
// other DG_C - 1 slide through registers, statically renamed by unrolling DG_C rows) from a copy
// of forest 2 interleaved by DG_C, which makes that load coalesced.
#include "pb2_common.cuh"

#define DG_C PB2_DIAG_LANES
#ifndef DG_THREADS
#define DG_THREADS 512
#endif
#ifndef DG_CHUNK
#define DG_CHUNK 8
#endif
#define DG_WARPS (DG_THREADS / 32)
#define DG_BLOCK (32 * DG_C)
#define DG_DEAD_START 0x3fffffff
#define DG_MAGIC 6755399441055744.0  // 2^52 + 2^51
#define DG_MAGIC_HI 0x43380000       // its high word: unchanged by adding 0 <= bin < 2^31
#define DG_NO_BIN ((int)0x80000000)

struct DiagConst {
    double kp_lo, kp_hi, kt_lo, kt_hi;  // n / range * (1 -+ 2^-40)
    int gmax;                           // diagonal blocks per forest pair (longest forests)
};

// first index in non-decreasing a[0], a[2], a[4] ... (n values, stride 2 doubles) with
// a[idx] > v (strict) or a[idx] >= v
__device__ __forceinline__ int dg_bound(const double *__restrict__ a, int n, double v, bool strict)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double x = __ldg(a + 2 * mid);
        if (strict ? (x <= v) : (x < v)) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// reference bin of a pair too close to a bin edge for the sandwich; (DG_NO_BIN, .) when rejected
__device__ __noinline__ int2 dg_exact_bin(const pb2_params &P, double rc1, double dm1, double rc2,
                                          double dm2, double ang, double ch, double sh)
{
    const int bin = pb2_pair_exact(P, rc1, dm1, rc2, dm2, ang, ch, sh, false, false).bin;
    if (bin < 0) return make_int2(DG_NO_BIN, 0);
    const int bp = bin / P.num_bins_r_trans;
    return make_int2(bp, bin - bp * P.num_bins_r_trans);
}

__device__ __forceinline__ void dg_emit(double *__restrict__ dst, int nb, int cnt, double a0,
                                        double a1, double a2, double a3, double a4)
{
    atomic_add_f64(dst, a0);
    atomic_add_f64(dst + (size_t)nb, a1);
    atomic_add_f64(dst + 2 * (size_t)nb, a2);
    atomic_add_f64(dst + 3 * (size_t)nb, a3);
    atomic_add_f64(dst + 4 * (size_t)nb, a4);
    atomic_add_i64(dst + 5 * (size_t)nb, (long long)cnt);
}

// ABS: auto-correlation (r_par = |r_par|, cf.py:361-362).  FOLD: r_par_min == 0, so the r_par bin
// is floor(|d| * (cos * K)) and cos folds into the constant; otherwise x = fl(fl(d cos) - min) is
// formed exactly as the reference does (cf.py:356,372).
template <bool ABS, bool FOLD>
__global__ void __launch_bounds__(DG_THREADS, 1)
pb2_xi_auto_diag(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, DiagConst C,
                 const int32_t *__restrict__ out_row, double *__restrict__ out)
{
    __shared__ unsigned s_ctr;
    if (threadIdx.x == 0) s_ctr = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const int np_i = P.num_bins_r_par, nt_i = P.num_bins_r_trans;
    const unsigned gmax = (unsigned)C.gmax;
    const unsigned units_per_chunk = DG_CHUNK * gmax;

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

        const int k1 = pr.nb_f1[e];
        const int f1 = pr.f1_index[k1];
        const int f2 = pr.nb_f2[e];
        const int n1 = c1.dg_count[f1], n2 = c2.dg_count[f2];
        if (n1 == 0 || n2 == 0) continue;
        const double ch = pr.nb_cos[e], sh = pr.nb_sin[e];
        const double *__restrict__ p_rc1 = c1.dg_rcdm + 2 * c1.dg_offset[f1];  // (rc, dm) pairs
        const double *__restrict__ p_rc2 = c2.dg_rcdm + 2 * c2.dg_offset[f2];

        // ---- diagonal range of the forest pair and row range of this block (supersets).
        // lane = segment of L consecutive rows; columns of row i in range: [jlo(i), jhi(i))
        const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
        const double dmax = P.r_par_max * inv_c * (1. + 1e-9) + 1e-9;
        const double dmin = P.r_par_min * inv_c;
        const double dlow = ABS ? -dmax : (dmin - fabs(dmin) * 1e-9 - 1e-9);
        const double tsum = P.r_trans_max * inv_s * (1. + 1e-9) + 1e-9;
        const int L = (n1 + 31) >> 5;
        const int s0 = lane * L, s1 = min(n1, s0 + L) - 1;
        int dlo = 0x7fffffff, dhi = -0x7fffffff;
        if (s0 < n1) {
            const int jlo = dg_bound(p_rc2, n2, __ldg(p_rc1 + 2 * s0) - dmax, true);
            int jhi = dg_bound(p_rc2, n2, __ldg(p_rc1 + 2 * s1) - dlow, false);
            if (isfinite(tsum))
                jhi = min(jhi, dg_bound(p_rc2 + 1, n2, tsum - __ldg(p_rc1 + 2 * s0 + 1), false));
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
        // blocks start at a multiple of DG_C (the interleaved copy is addressed by phase)
        const int Dbase = Dmin - (((Dmin % DG_C) + DG_C) % DG_C);
        const int D0 = Dbase + DG_BLOCK * g;
        if (D0 > Dmax) continue;
        const int D1 = D0 + DG_BLOCK - 1;
        const unsigned segs = __ballot_sync(0xffffffffu, dlo <= D1 && dhi >= D0);
        if (!segs) continue;
        int ibeg = (__ffs(segs) - 1) * L;
        int iend = min(n1, (32 - __clz(segs)) * L);
        // lanes read columns i + D0 .. i + D0 + DG_BLOCK - 1: rows whose whole span lies outside
        // forest 2 are skipped; everything else stays inside the padding of the packed copies
        ibeg = max(ibeg, max(0, -D0 - (DG_BLOCK - 1)));
        iend = min(iend, n2 - D0);
        if (ibeg >= iend) continue;
        ibeg -= ibeg % DG_C;

        double *__restrict__ const orow = out + (size_t)out_row[k1] * 6 * nb;
        const double ang = pr.nb_ang[e];
        // bin constants of this forest pair
        const double kpl = FOLD ? mul_rn(ch, C.kp_lo) : C.kp_lo;
        const double kph = FOLD ? mul_rn(ch, C.kp_hi) : C.kp_hi;
        const double ktl = mul_rn(sh, C.kt_lo), kth = mul_rn(sh, C.kt_hi);

        // per diagonal: sums of the current run, the step it started at, the low words of its
        // bin; `lv` bit k = the run of diagonal k is in range.  No run is open at the start.
        double a0[DG_C], a1[DG_C], a2[DG_C], a3[DG_C], a4[DG_C];
        int curp[DG_C], curt[DG_C], start[DG_C];
        unsigned lv = 0;
#pragma unroll
        for (int k = 0; k < DG_C; k++) {
            a0[k] = a1[k] = a2[k] = a3[k] = a4[k] = 0.;
            curp[k] = DG_NO_BIN;
            curt[k] = 0;
            start[k] = 0;
        }

        // rows: natural-order packed copy of forest 1 (uniform loads).  32-bit element indices
        // (checked by pb2_xi_diag_eligible) keep the address registers few.
        const unsigned ra = (unsigned)c1.dg_offset[f1];
        const double2 *__restrict__ p_r1 = reinterpret_cast<const double2 *>(c1.dg_rcdm);
        const double2 *__restrict__ p_w1 = reinterpret_cast<const double2 *>(c1.dg_wdw);
        const double *__restrict__ p_z1 = c1.dg_z;
        // columns: copy of forest 2 interleaved by DG_C; padded column jp = j + PB2_DIAG_PAD sits
        // in plane jp % DG_C at il_offset[f2] + jp / DG_C.  This lane's first column is
        // ibeg + D0 + DG_C * lane (a multiple of DG_C after padding).
        const unsigned plane = (unsigned)c2.il_total;
        unsigned cpos = (unsigned)c2.il_offset[f2] + (unsigned)((ibeg + D0 + PB2_DIAG_PAD) / DG_C + lane);
        const double2 *__restrict__ q_r2 = reinterpret_cast<const double2 *>(c2.il_rcdm);
        const double2 *__restrict__ q_w2 = reinterpret_cast<const double2 *>(c2.il_wdw);
        const double *__restrict__ q_z2 = c2.il_z;

        double2 cr[DG_C], cw[DG_C];
        double cz[DG_C];
#pragma unroll
        for (int k = 0; k < DG_C - 1; k++) {
            cr[k] = __ldg(q_r2 + (k * plane + cpos));
            cw[k] = __ldg(q_w2 + (k * plane + cpos));
            cz[k] = __ldg(q_z2 + (k * plane + cpos));
        }

        for (int s = ibeg; s < iend; s += DG_C) {
#pragma unroll
            for (int uu = 0; uu < DG_C; uu++) {
                const unsigned rat = ra + (unsigned)(s + uu);
                const double2 r1 = __ldg(p_r1 + rat);  // (rc1, dm1), uniform
                const double2 w1 = __ldg(p_w1 + rat);  // (w1, delta1 w1), uniform
                const double z1 = __ldg(p_z1 + rat);
                {
                    // the new column of this row: ibeg.. + DG_C * lane + uu + DG_C - 1
                    const int pl = (uu + DG_C - 1) % DG_C;
                    const unsigned at = pl * plane + cpos + (uu + DG_C - 1) / DG_C;
                    cr[pl] = __ldg(q_r2 + at);
                    cw[pl] = __ldg(q_w2 + at);
                    cz[pl] = __ldg(q_z2 + at);
                }
                // phase 1: geometry and bins of the DG_C pairs of this row (independent chains)
                const int sidx = s + uu;
                double v[DG_C], t[DG_C];
                bool chg[DG_C];
                bool any = false;
#pragma unroll
                for (int k = 0; k < DG_C; k++) {
                    const int sl = (uu + k) % DG_C;
                    const double d = sub_rn(r1.x, cr[sl].x);
                    v[k] = ABS ? fabs(d) : d;
                    t[k] = add_rn(r1.y, cr[sl].y);
                    const double x = FOLD ? v[k] : sub_rn(mul_rn(v[k], ch), P.r_par_min);
                    const int bpl = __double2loint(__fma_rd(x, kpl, DG_MAGIC));
                    const int bph = __double2loint(__fma_rd(x, kph, DG_MAGIC));
                    const int btl = __double2loint(__fma_rd(t[k], ktl, DG_MAGIC));
                    const int bth = __double2loint(__fma_rd(t[k], kth, DG_MAGIC));
                    chg[k] = (((bpl ^ curp[k]) | (bph ^ curp[k])) |
                              ((btl ^ curt[k]) | (bth ^ curt[k]))) != 0;
                    any = any || chg[k];
                }
                // phase 2: run changes (one branch per row in the common case)
                if (any) {
#pragma unroll
                    for (int k = 0; k < DG_C; k++) {
                        if (chg[k]) {
                            const int sl = (uu + k) % DG_C;
                            // ---- diagonal k left its run: add the run to its bin
                            if (lv & (1u << k))
                                dg_emit(orow + (curp[k] * nt_i + curt[k]), nb, sidx - start[k], a0[k],
                                        a1[k], a2[k] * ch, a3[k] * sh, a4[k] * 0.5);
                            // ---- the new run: proven bin, or the reference expression.  Dummy
                            // pixels (distance 1e300) leave the high word of the r_trans FMA off
                            // 2^52 + 2^51
                            // (recomputed behind an opaque copy: keeping phase 1's values alive
                            // for this rare path costs four register moves per pair)
                            double x = FOLD ? v[k] : sub_rn(mul_rn(v[k], ch), P.r_par_min);
                            double tt = t[k];
                            asm volatile("" : "+d"(x), "+d"(tt));
                            const double utl = __fma_rd(tt, ktl, DG_MAGIC);
                            const int bpl = __double2loint(__fma_rd(x, kpl, DG_MAGIC));
                            const int bph = __double2loint(__fma_rd(x, kph, DG_MAGIC));
                            const int btl = __double2loint(utl);
                            const int bth = __double2loint(__fma_rd(tt, kth, DG_MAGIC));
                            const bool fmt = __double2hiint(utl) == DG_MAGIC_HI;
                            int nbp = bpl, nbt = btl;
                            bool live = fmt && (unsigned)bpl < (unsigned)np_i &&
                                        (unsigned)btl < (unsigned)nt_i;
                            if (fmt && (bpl != bph || btl != bth)) {
                                const int2 b = dg_exact_bin(P, r1.x, r1.y, cr[sl].x, cr[sl].y, ang,
                                                            ch, sh);
                                nbp = b.x;
                                nbt = b.y;
                                live = b.x != DG_NO_BIN;
                            }
                            curp[k] = nbp;
                            curt[k] = nbt;
                            start[k] = sidx;
                            lv = live ? (lv | (1u << k)) : (lv & ~(1u << k));
                        }
                    }
                }
                // phase 3: restart the sums of changed runs (see the header) and accumulate
#pragma unroll
                for (int k = 0; k < DG_C; k++) {
                    const int sl = (uu + k) % DG_C;
                    a0[k] = __hiloint2double(chg[k] ? 0 : __double2hiint(a0[k]), __double2loint(a0[k]));
                    a1[k] = __hiloint2double(chg[k] ? 0 : __double2hiint(a1[k]), __double2loint(a1[k]));
                    a2[k] = __hiloint2double(chg[k] ? 0 : __double2hiint(a2[k]), __double2loint(a2[k]));
                    a3[k] = __hiloint2double(chg[k] ? 0 : __double2hiint(a3[k]), __double2loint(a3[k]));
                    a4[k] = __hiloint2double(chg[k] ? 0 : __double2hiint(a4[k]), __double2loint(a4[k]));
                    const double w12 = mul_rn(w1.x, cw[sl].x);
                    a0[k] += w12;
                    a1[k] = fma(w1.y, cw[sl].y, a1[k]);
                    a2[k] = fma(v[k], w12, a2[k]);
                    a3[k] = fma(t[k], w12, a3[k]);
                    a4[k] = fma(add_rn(z1, cz[sl]), w12, a4[k]);
                }
            }
            cpos += 1;
        }
        // ---- last runs of the unit's diagonals (rows were walked up to a multiple of DG_C)
        const int send = ibeg + ((iend - ibeg + DG_C - 1) / DG_C) * DG_C;
#pragma unroll
        for (int k = 0; k < DG_C; k++) {
            if (lv & (1u << k))
                dg_emit(orow + (curp[k] * nt_i + curt[k]), nb, send - start[k], a0[k], a1[k],
                        a2[k] * ch, a3[k] * sh, a4[k] * 0.5);
        }
    }
}

template <bool ABS, bool FOLD>
static int32_t dg_launch(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                         const pb2_pairs *pairs, const DiagConst &C, const int32_t *d_out_row,
                         double *d_out, int blocks, cudaStream_t s)
{
    pb2_xi_auto_diag<ABS, FOLD><<<blocks, DG_THREADS, 0, s>>>(*c1, *c2, *par, *pairs, C,
                                                                  d_out_row, d_out);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_xi_auto_diag");
}

// can this launch use the diagonal-lane kernel?  (the caller has checked the binning mode)
bool pb2_xi_diag_eligible(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                          int64_t n_rows)
{
    if (!c1->dg_rcdm || !c2->il_rcdm || c1->dg_lanes != DG_C || c2->dg_lanes != DG_C) return false;
    if (!c1->dg_ok || !c2->dg_ok) return false;
    const double nb6 = 6. * par->num_bins_r_par * par->num_bins_r_trans;
    if ((double)n_rows * nb6 >= 4294967296.) return false;  // 32-bit histogram offsets
    // low-word bins: |x K| must stay far below 2^31 for every real pair
    const double kp = (double)par->num_bins_r_par / (par->r_par_max - par->r_par_min);
    const double kt = (double)par->num_bins_r_trans / par->r_trans_max;
    const double reach = c1->dg_reach + c2->dg_reach;
    const double rmin = par->r_par_min < 0 ? -par->r_par_min : par->r_par_min;
    if (!((reach + rmin) * kp < 1e9) || !(reach * kt < 1e9)) return false;
    return true;
}

int32_t pb2_launch_xi_diag(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                           const pb2_pairs *pairs, const int32_t *d_out_row, double *d_out,
                           cudaStream_t s)
{
    DiagConst C;
    C.gmax = (c1->dg_max_pix + c2->dg_max_pix + DG_C + DG_BLOCK - 1) / DG_BLOCK + 1;
    const double kp = (double)par->num_bins_r_par / (par->r_par_max - par->r_par_min);
    const double kt = (double)par->num_bins_r_trans / par->r_trans_max;
    const double eps = 9.094947017729282e-13;  // 2^-40
    C.kp_lo = kp * (1. - eps);
    C.kp_hi = kp * (1. + eps);
    C.kt_lo = kt * (1. - eps);
    C.kt_hi = kt * (1. + eps);
    int dev = 0, sms = 0;
    PB2_CUDA(cudaGetDevice(&dev));
    PB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long want = (pairs->n_pairs + DG_CHUNK - 1) / DG_CHUNK;
    int blocks = (int)(want < sms ? want : sms);
    if (blocks < 1) blocks = 1;
    if (par->x_correlation)
        return dg_launch<false, false>(c1, c2, par, pairs, C, d_out_row, d_out, blocks, s);
    if (par->r_par_min != 0.)
        return dg_launch<true, false>(c1, c2, par, pairs, C, d_out_row, d_out, blocks, s);
    return dg_launch<true, true>(c1, c2, par, pairs, C, d_out_row, d_out, blocks, s);
}
