// Auto / delta x delta pixel-pair histogram, standard (r_par, r_trans) binning, no per-pair cuts:
// "diagonal lanes" kernel -- the product path of picca_cf.py's default mode.  Replaces
// cf.compute_xi's pair loop + cf.compute_xi_forest_pairs_fast (reference py/picca/cf.py:161-240,
// 250-387).
//
// For one forest pair the pixel pairs (i, j) are visited along diagonals j - i = const.  On a
// diagonal r_par = (rc1[i] - rc2[j]) cos(ang/2) is nearly constant and r_trans = (dm1[i] + dm2[j])
// sin(ang/2) grows slowly, so a diagonal stays in ONE (r_par, r_trans) bin for a "run" of ~45
// consecutive pairs.  Each lane owns DG_C adjacent diagonals and keeps, per diagonal, five sums
// of the current run and its bin window in registers.
//   per pair    d = rc1 - rc2, t = dm1 + dm2, then ONE round-down FMA per dimension against
//               2^52 + 2^51 leaves floor(65536 x K) in the low word of the result: the bin in the
//               upper half, a 16-bit fraction in the lower (K = n / range, with cos or sin of the
//               half angle folded in).  The reference's floor((r - min) / (max - min) * n)
//               (cf.py:372-376) carries a rounding error below 1e-11 bins, so a fraction in
//               [1, 65534] -- the exact product at least 2^-16 bins from an edge -- proves the
//               bin, and "bin < n" is then the reference's range test (cf.py:364).  The pair stays
//               in its run while (unsigned)(low word - (65536 bin + 1)) < 65534 in both
//               dimensions: two integer subtractions and one compare.  With w1 w2 and six
//               accumulate instructions that is 11 FP64 instructions per pair, no division, no
//               predication, no bounds check and no pair counter: zero-weight pixels (skipped by
//               the reference, cf.py:318,331) are compacted away in the packed copies, so
//               num_pairs of a run is its length, and columns outside forest 2 read dummies
//               (distance 1e300, weight 0) that add zeros and land in no bin.
//   run change  the lane adds its five sums and the run length to the finished run's bin with six
//               native red.global.add.f64 / .u64 (~45 instructions, divergent; the LATENCY of this
//               path -- the rest of the warp waits at the reconvergence point -- bounds the kernel,
//               see dg_emit).  The sums restart
//               through five selects in the MAIN path (high word := 0 when the bin changed, which
//               leaves at most a 1e-314 denormal behind): the accumulators are then written only
//               by the accumulate instructions and ptxas keeps them in place -- clearing them
//               inside the branch costs ten register moves per pair, snapshots in shared memory
//               saturate the L1 data pipe.  A pair whose fraction is 65535, or 0 above bin 0, lies
//               within 2^-16 of a bin edge (exactly on it when the spectra share a wavelength
//               grid): its bin comes from the reference expression with IEEE divisions
//               (dg_exact_bin) and it forms a run of its own.  Bin 0 accepts a zero fraction: its
//               lower edge is x = 0 itself (d = 0 for equal wavelengths).
// Work unit = forest pair, one warp; inside it, blocks of 32 * DG_C diagonals.  The warp walks a
// block's rows in chunks of DG_R: one lane issues TMA bulk copies (cp.async.bulk, completion on an
// mbarrier) of the chunk's row records and of the column records it needs from a copy of forest 2
// interleaved by DG_C, into a two-stage per-warp buffer in shared memory, one chunk ahead of the
// arithmetic and across block boundaries.  Pixels are 48-byte records (r_comov, dist_m, weight,
// delta*weight, z/2, 0).  Per row the warp reads the row's record with broadcast loads and each
// lane ONE new column record (the other DG_C - 1 slide through registers, statically renamed by
// unrolling DG_C rows); all addresses are a running shared-memory pointer plus immediates.
// The histogram is accumulated in a [row][bin][8]-slot scratch (the six sums of a bin share one
// 64-byte line: one address per run change) and folded into the caller's [row][6][bin] layout by
// pb2_xi_diag_fold.
#include "pb2_common.cuh"

#define DG_C PB2_DIAG_LANES
#define DG_R PB2_DIAG_CHUNK_ROWS
#ifndef DG_THREADS
#define DG_THREADS 512
#endif
#ifndef DG_CHUNK
#define DG_CHUNK 4
#endif
#define DG_WARPS (DG_THREADS / 32)
#define DG_BLOCK (32 * DG_C)
#define DG_MAGIC 6755399441055744.0  // 2^52 + 2^51
#define DG_MAGIC_HI 0x43380000       // its high word: unchanged by adding 0 <= bin < 2^31
#define DG_NO_BIN ((int)0x80000000)
#define DG_GMAX 96                                  // diagonal blocks per forest pair, at most
#define DG_REC 48                                   // bytes per pixel record
#define DG_ROW_BYTES (DG_R * DG_REC)                // row records of a chunk
#define DG_PLANE_REC (DG_R / DG_C + 33)             // column records of a chunk, per plane
#define DG_PLANE_BYTES (DG_PLANE_REC * DG_REC)
#define DG_STAGE_BYTES (DG_ROW_BYTES + DG_C * DG_PLANE_BYTES)

static_assert(DG_R % DG_C == 0, "chunk rows must be a multiple of the diagonals per lane");
static_assert(PB2_DIAG_PAD % DG_C == 0, "padding must be a multiple of the diagonals per lane");

struct DiagConst {
    double kp16, kt16;  // 65536 n / range
};

// ---- TMA bulk copy + mbarrier (PTX ISA: cp.async.bulk, mbarrier)
__device__ __forceinline__ unsigned dg_saddr(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void dg_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void dg_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void dg_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void dg_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DG_DONE;\n"
        "bra DG_WAIT;\n"
        "DG_DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// plain (schedulable) shared-memory loads at a byte offset of the dynamic buffer
__device__ __forceinline__ double2 dg_lds128(const unsigned char *p)
{
    return *reinterpret_cast<const double2 *>(p);
}
__device__ __forceinline__ double dg_lds64(const unsigned char *p)
{
    return *reinterpret_cast<const double *>(p);
}

// first index in non-decreasing a[0], a[6], a[12] ... (n records of 6 doubles) with
// a[idx] > v (strict) or a[idx] >= v
__device__ __forceinline__ int dg_bound(const double *__restrict__ a, int n, double v, bool strict)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double x = __ldg(a + 6 * mid);
        if (strict ? (x <= v) : (x < v)) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// reference bin (flat; -1 when rejected) of a pair too close to a bin edge for the proof
__device__ __noinline__ int dg_exact_bin(const pb2_params &P, double rc1, double dm1, double rc2,
                                         double dm2, double ang, double ch, double sh)
{
    return pb2_pair_exact(P, rc1, dm1, rc2, dm2, ang, ch, sh, false, false).bin;
}

// one finished run -> the six sums of its bin (one 64-byte line of the scratch histogram), through
// red.global on a 64-bit address + immediates: the row's scratch address lives in ONE register pair
// for the whole forest pair.  (With a C++ pointer ptxas re-adds the kernel parameter at every run
// change -- a constant-bank load and two adds on the latency-critical path of the warp's divergent
// lanes; without them the kernel is 7 % faster.)
__device__ __forceinline__ void dg_emit(unsigned long long dst, int cnt, double a0, double a1,
                                        double a2, double a3, double a4)
{
    asm volatile(
        "red.global.add.f64 [%0], %1;\n\t"
        "red.global.add.f64 [%0+8], %2;\n\t"
        "red.global.add.f64 [%0+16], %3;\n\t"
        "red.global.add.f64 [%0+24], %4;\n\t"
        "red.global.add.f64 [%0+32], %5;\n\t"
        "red.global.add.u64 [%0+40], %6;" ::"l"(dst), "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4),
        "l"((unsigned long long)(unsigned)cnt)
        : "memory");
}
#define DG_BIN_ADDR(srow, bin) ((srow) + ((unsigned long long)(unsigned)(bin) << 6))

// ABS: auto-correlation (r_par = |r_par|, cf.py:361-362).  FOLD: r_par_min == 0, so the r_par bin
// is floor(|d| * (cos * K)) and cos folds into the constant; otherwise x = fl(fl(d cos) - min) is
// formed exactly as the reference does (cf.py:356,372).
template <bool ABS, bool FOLD>
__global__ void __launch_bounds__(DG_THREADS, 1)
pb2_xi_auto_diag(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, DiagConst C,
                 const int32_t *__restrict__ out_row, double *__restrict__ scr,
                 unsigned long long *__restrict__ work_ctr)
{
    extern __shared__ __align__(128) unsigned char dg_smem[];
    __shared__ __align__(8) unsigned long long s_bar[DG_WARPS][2];
    __shared__ int2 s_tb[DG_WARPS][DG_GMAX];
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    if (lane == 0) {
        dg_mbar_init(dg_saddr(&s_bar[wid][0]), 1);
        dg_mbar_init(dg_saddr(&s_bar[wid][1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned bar0 = dg_saddr(&s_bar[wid][0]);                 // stage t: bar0 + 8 t
    const unsigned buf0 = dg_saddr(dg_smem) + wid * (2 * DG_STAGE_BYTES);  // stage t: + t STAGE
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const int np_i = P.num_bins_r_par, nt_i = P.num_bins_r_trans;
    const double *__restrict__ rec1 = c1.dg_rec;
    const double *__restrict__ rec2 = c2.il_rec;
    int2 *const tb = s_tb[wid];  // per block of the current forest pair: (first row, walked rows)
    unsigned stage = 0;  // stage of the next chunk to consume (warp-uniform)
    unsigned phase = 0;  // bit t: parity to wait for on stage t

    for (;;) {
        // ---- claim the next forest pair of the list from ONE device-wide counter: at any time
        // the warps of the whole grid work inside a window of a few thousand consecutive forest
        // pairs (a handful of forests 1 and their shared neighbours), which stays in L2.  A static
        // split lets the CTAs drift apart by a few per cent of the list -- hundreds of MB of
        // records -- and re-reads the catalogue from DRAM ~50 times (ncu: 454 GB per launch).
        unsigned long long u = 0;
        if (lane == 0) u = atomicAdd(work_ctr, 1ull);
        const long long e = (long long)__shfl_sync(0xffffffffu, u, 0);
        if (e >= pr.n_pairs) break;

        const int k1 = pr.nb_f1[e];
        const int f1 = pr.f1_index[k1];
        const int f2 = pr.nb_f2[e];
        const int n1 = c1.dg_count[f1], n2 = c2.dg_count[f2];
        if (n1 == 0 || n2 == 0) continue;
        const double ch = pr.nb_cos[e], sh = pr.nb_sin[e];
        const long long ra = c1.dg_offset[f1];
        const double *__restrict__ p_rc1 = rec1 + 6 * ra;  // records (rc, dm, w, dw, z/2, 0)
        const double *__restrict__ p_rc2 = c2.dg_rec + 6 * c2.dg_offset[f2];

        // ---- diagonal range of the forest pair and row range of every block (supersets).
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
            const int jlo = dg_bound(p_rc2, n2, __ldg(p_rc1 + 6 * s0) - dmax, true);
            int jhi = dg_bound(p_rc2, n2, __ldg(p_rc1 + 6 * s1) - dlow, false);
            if (isfinite(tsum))
                jhi = min(jhi, dg_bound(p_rc2 + 1, n2, tsum - __ldg(p_rc1 + 6 * s0 + 1), false));
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
        const int G = min((Dmax - Dbase) / DG_BLOCK + 1, DG_GMAX);
        __syncwarp();
        for (int g = 0; g < G; g++) {
            const int D0 = Dbase + DG_BLOCK * g, D1 = D0 + DG_BLOCK - 1;
            const unsigned segs = __ballot_sync(0xffffffffu, dlo <= D1 && dhi >= D0);
            int ibeg = 0, nrows = 0;
            if (segs) {
                ibeg = (__ffs(segs) - 1) * L;
                int iend = min(n1, (32 - __clz(segs)) * L);
                // lanes read columns i + D0 .. i + D0 + DG_BLOCK - 1: rows whose whole span lies
                // outside forest 2 are skipped; everything else stays inside the padding of the
                // packed copies
                ibeg = max(ibeg, max(0, -D0 - (DG_BLOCK - 1)));
                iend = min(iend, n2 - D0);
                ibeg -= ibeg % DG_C;
                if (iend > ibeg) nrows = ((iend - ibeg + DG_C - 1) / DG_C) * DG_C;  // walked rows
            }
            if (lane == 0) tb[g] = make_int2(ibeg, nrows);
        }
        __syncwarp();

        // ---- chunk c of block g: rows ibeg + c DG_R ..., and per plane the column records from
        // this lane-0 position on.  Padded column jp = j + PB2_DIAG_PAD sits in plane jp % DG_C at
        // il_offset[f2] + jp / DG_C; lane l's first column is ibeg + D0 + DG_C l.  The two-stage
        // pipeline runs across the blocks of the forest pair: (pg, pc) = next chunk to request.
        const double *__restrict__ g_row = rec1 + 6 * ra;
        const double *__restrict__ g_col = rec2 + 6 * (c2.il_offset[f2] + (Dbase + PB2_DIAG_PAD) / DG_C);
        const long long plane6 = 6 * c2.il_total;
        int pg = 0, pc = 0;
        auto request = [&](unsigned t) {  // all lanes advance (pg, pc); lane 0 issues the copies
            while (pg < G && pc * DG_R >= tb[pg].y) {
                pg++;
                pc = 0;
            }
            if (pg >= G) return;
            if (lane == 0) {
                const int ib = tb[pg].x;
                const unsigned bar = bar0 + 8 * t, dst = buf0 + t * DG_STAGE_BYTES;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                dg_mbar_expect_tx(bar, DG_STAGE_BYTES);
                dg_bulk_g2s(dst, g_row + (size_t)(ib + pc * DG_R) * 6, DG_ROW_BYTES, bar);
                const size_t col = (size_t)((ib + pg * DG_BLOCK) / DG_C + pc * (DG_R / DG_C)) * 6;
#pragma unroll
                for (int p = 0; p < DG_C; p++)
                    dg_bulk_g2s(dst + DG_ROW_BYTES + p * DG_PLANE_BYTES, g_col + p * plane6 + col,
                                DG_PLANE_BYTES, bar);
            }
            pc++;
        };
        request(stage);
        request(stage ^ 1u);

        // the full scratch address of the row in one register pair (see dg_emit)
        unsigned long long srow = (unsigned long long)__cvta_generic_to_global(
            scr + (size_t)out_row[k1] * nb * 8);
        asm volatile("" : "+l"(srow));
        const double ang = pr.nb_ang[e];
        // bin constants of this forest pair
        const unsigned np16 = (unsigned)np_i << 16, nt16 = (unsigned)nt_i << 16;
        const unsigned nt_r = (unsigned)nt_i;
        const double kpf = FOLD ? mul_rn(ch, C.kp16) : C.kp16;
        const double ktf = mul_rn(sh, C.kt16);

        for (int g = 0; g < G; g++) {
        const int ibeg = tb[g].x, nrows = tb[g].y;
        if (nrows == 0) continue;
        const int nchunk = (nrows + DG_R - 1) / DG_R;
        // per diagonal: sums of the current run, the step it started at, its flat bin (-1: out
        // of range) and, per dimension, 65536 * bin + 1: a pair stays in the run while both
        // (unsigned)(low word - that) < 65534.  No run is open at the start.
        double a0[DG_C], a1[DG_C], a2[DG_C], a3[DG_C], a4[DG_C];
        int bp1[DG_C], bt1[DG_C], cb[DG_C], start[DG_C];
#pragma unroll
        for (int k = 0; k < DG_C; k++) {
            a0[k] = a1[k] = a2[k] = a3[k] = a4[k] = 0.;
            bp1[k] = bt1[k] = DG_NO_BIN;
            cb[k] = -1;
            start[k] = 0;
        }
        double2 cr[DG_C], cw[DG_C];
        double cz[DG_C];
        int s = ibeg;  // row of the next step

        for (int c = 0; c < nchunk; c++) {
            dg_mbar_wait(bar0 + 8 * stage, (phase >> stage) & 1u);
            phase ^= 1u << stage;
            // row record of the next step; this lane's column records
            const unsigned char *rp = dg_smem + (wid * 2 + stage) * DG_STAGE_BYTES;
            const unsigned char *cp = rp + DG_ROW_BYTES + lane * DG_REC;
            if (c == 0) {
#pragma unroll
                for (int k = 0; k < DG_C - 1; k++) {
                    cr[k] = dg_lds128(cp + k * DG_PLANE_BYTES);
                    cw[k] = dg_lds128(cp + k * DG_PLANE_BYTES + 16);
                    cz[k] = dg_lds64(cp + k * DG_PLANE_BYTES + 32);
                }
            }
            const int nit = min(DG_R, nrows - c * DG_R) / DG_C;
            for (int it = 0; it < nit; it++) {
#pragma unroll
                for (int uu = 0; uu < DG_C; uu++) {
                    const double2 r1 = dg_lds128(rp + uu * DG_REC);       // (rc1, dm1), broadcast
                    const double2 w1 = dg_lds128(rp + uu * DG_REC + 16);  // (w1, delta1 w1)
                    const double z1 = dg_lds64(rp + uu * DG_REC + 32);    // z1 / 2
                    {
                        // the new column of this row: row + D0 + DG_C * lane + DG_C - 1
                        const int pl = (uu + DG_C - 1) % DG_C;
                        const unsigned char *at = cp + pl * DG_PLANE_BYTES + ((uu + DG_C - 1) / DG_C) * DG_REC;
                        cr[pl] = dg_lds128(at);
                        cw[pl] = dg_lds128(at + 16);
                        cz[pl] = dg_lds64(at + 32);
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
                        const int lp = __double2loint(__fma_rd(x, kpf, DG_MAGIC));
                        const int lt = __double2loint(__fma_rd(t[k], ktf, DG_MAGIC));
                        chg[k] = (unsigned)(lp - bp1[k]) >= 65534u || (unsigned)(lt - bt1[k]) >= 65534u;
                        any = any || chg[k];
                    }
                    // phase 2: run changes (one branch per row in the common case)
                    if (any) {
#pragma unroll
                        for (int k = 0; k < DG_C; k++) {
                            if (chg[k]) {
                                const int sl = (uu + k) % DG_C;
                                // ---- diagonal k left its run: add the run to its bin
                                if (cb[k] >= 0)
                                    dg_emit(DG_BIN_ADDR(srow, cb[k]), sidx - start[k], a0[k], a1[k],
                                            a2[k] * ch, a3[k] * sh, a4[k]);
                                // ---- the new run.  The low words hold floor(65536 x K): bin in
                                // the upper, a 16-bit fraction in the lower half.  (Recomputed
                                // behind an opaque copy: keeping phase 1's values alive for this
                                // path costs register moves on every pair.)
                                double x = FOLD ? v[k] : sub_rn(mul_rn(v[k], ch), P.r_par_min);
                                double tt = t[k];
                                asm volatile("" : "+d"(x), "+d"(tt));
                                const double ut = __fma_rd(tt, ktf, DG_MAGIC);
                                const int lp = __double2loint(__fma_rd(x, kpf, DG_MAGIC));
                                const int lt = __double2loint(ut);
                                // dummy pixels (distance 1e300) leave the high word off 2^52 + 2^51
                                const bool fmt = __double2hiint(ut) == DG_MAGIC_HI;
                                // new window per dimension: 65536 bin + 1 (bin 0 also accepts a
                                // zero fraction: its lower edge is x = 0 itself, which pixels of
                                // a common wavelength grid hit exactly with d = 0)
                                const int hp = lp & (int)0xffff0000, ht = lt & (int)0xffff0000;
                                int nb1p = hp + (hp != 0), nb1t = ht + (ht != 0);
                                const unsigned bp = (unsigned)lp >> 16, bt = (unsigned)lt >> 16;
                                int ncb = (int)(bp * nt_r + bt);
                                if ((unsigned)lp >= np16 || (unsigned)lt >= nt16) ncb = -1;
                                if (!fmt) {  // dummy pixel: no bin, and quiet while it lasts
                                    nb1p = lp - 1;
                                    nb1t = lt - 1;
                                    ncb = -1;
                                }
                                // outside its own window: a fraction of 65535 (65534), or of 0
                                // above bin 0 -> within 2^-16 of a bin edge: the reference
                                // expression decides, and the pair is a run of its own
                                if (max((unsigned)(lp - nb1p), (unsigned)(lt - nb1t)) >= 65534u) {
                                    ncb = dg_exact_bin(P, r1.x, r1.y, cr[sl].x, cr[sl].y, ang, ch, sh);
                                    nb1p = lp ^ DG_NO_BIN;
                                    nb1t = lt ^ DG_NO_BIN;
                                }
                                bp1[k] = nb1p;
                                bt1[k] = nb1t;
                                cb[k] = ncb;
                                start[k] = sidx;
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
                s += DG_C;
                rp += DG_C * DG_REC;
                cp += DG_REC;
            }
            // the stage is free again: refill it with the chunk after the next one
            __syncwarp();
            request(stage);
            stage ^= 1u;
        }
        // ---- last runs of the block's diagonals
#pragma unroll
        for (int k = 0; k < DG_C; k++) {
            if (cb[k] >= 0)
                dg_emit(DG_BIN_ADDR(srow, cb[k]), s - start[k], a0[k], a1[k], a2[k] * ch, a3[k] * sh,
                        a4[k]);
        }
        }  // blocks of the forest pair
    }
}

// scratch [row][bin][8] -> the caller's [row][6][bin] (accumulated; slot 5 is an int64 count)
__global__ void pb2_xi_diag_fold(const double *__restrict__ scr, double *__restrict__ out,
                                 long long n_rows, int nb)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * nb) return;
    const long long row = idx / nb;
    const int bin = (int)(idx - row * nb);
    const double2 *src = reinterpret_cast<const double2 *>(scr + idx * 8);
    const double2 v01 = src[0], v23 = src[1], v45 = src[2];
    double *dst = out + row * 6 * (long long)nb + bin;
    dst[0] += v01.x;
    dst[(size_t)nb] += v01.y;
    dst[2 * (size_t)nb] += v23.x;
    dst[3 * (size_t)nb] += v23.y;
    dst[4 * (size_t)nb] += v45.x;
    long long *cnt = reinterpret_cast<long long *>(dst + 5 * (size_t)nb);
    *cnt += __double_as_longlong(v45.y);
}

template <bool ABS, bool FOLD>
static int32_t dg_launch(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                         const pb2_pairs *pairs, const DiagConst &C, const int32_t *d_out_row,
                         double *d_scr, unsigned long long *d_ctr, int blocks, cudaStream_t s)
{
    const size_t smem = (size_t)DG_WARPS * 2 * DG_STAGE_BYTES;
    PB2_CUDA(cudaFuncSetAttribute(pb2_xi_auto_diag<ABS, FOLD>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pb2_xi_auto_diag<ABS, FOLD><<<blocks, DG_THREADS, smem, s>>>(*c1, *c2, *par, *pairs, C,
                                                                  d_out_row, d_scr, d_ctr);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_xi_auto_diag");
}

// can this launch use the diagonal-lane kernel?  (the caller has checked the binning mode)
bool pb2_xi_diag_eligible(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                          int64_t n_rows)
{
    if (!c1->dg_rec || !c2->dg_rec || !c2->il_rec || c1->dg_lanes != DG_C || c2->dg_lanes != DG_C)
        return false;
    if (!c1->dg_ok || !c2->dg_ok) return false;
    if ((c1->dg_max_pix + c2->dg_max_pix + DG_C) / DG_BLOCK + 2 > DG_GMAX) return false;
    const double nb6 = 6. * par->num_bins_r_par * par->num_bins_r_trans;
    if ((double)n_rows * nb6 >= 2147483648.) return false;  // 32-bit bin arithmetic
    // low-word bins: |x K| must stay far below 2^31 for every real pair
    const double kp = (double)par->num_bins_r_par / (par->r_par_max - par->r_par_min);
    const double kt = (double)par->num_bins_r_trans / par->r_trans_max;
    const double reach = c1->dg_reach + c2->dg_reach;
    const double rmin = par->r_par_min < 0 ? -par->r_par_min : par->r_par_min;
    // floor(65536 x K) must stay far below 2^31 for every real pair: bins < 8192
    if (!((reach + rmin) * kp < 8192.) || !(reach * kt < 8192.)) return false;
    if (par->num_bins_r_par > 4096 || par->num_bins_r_trans > 4096) return false;
    return true;
}

int32_t pb2_launch_xi_diag(const pb2_catalog *c1, const pb2_catalog *c2, const pb2_params *par,
                           const pb2_pairs *pairs, const int32_t *d_out_row, int64_t n_rows,
                           double *d_out, cudaStream_t s)
{
    DiagConst C;
    const double kp = (double)par->num_bins_r_par / (par->r_par_max - par->r_par_min);
    const double kt = (double)par->num_bins_r_trans / par->r_trans_max;
    C.kp16 = kp * 65536.;
    C.kt16 = kt * 65536.;
    int dev = 0, sms = 0;
    PB2_CUDA(cudaGetDevice(&dev));
    PB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long want = (pairs->n_pairs + DG_CHUNK - 1) / DG_CHUNK;
    int blocks = (int)(want < sms ? want : sms);
    if (blocks < 1) blocks = 1;
    // scratch histogram [row][bin][8], folded into d_out after the pair kernel
    const int nb = par->num_bins_r_par * par->num_bins_r_trans;
    const size_t scr_bytes = (size_t)n_rows * nb * 8 * sizeof(double);
    double *d_scr = nullptr;
    {
        // keep the stream-ordered pool's memory across calls (the default releases it at every
        // synchronisation, which costs ~0.5 s per call for a few hundred MB)
        cudaMemPool_t pool;
        PB2_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;
        PB2_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    // + the work counter behind the histogram (zeroed by the same memset)
    PB2_CUDA(cudaMallocAsync((void **)&d_scr, scr_bytes + 256, s));
    PB2_CUDA(cudaMemsetAsync(d_scr, 0, scr_bytes + 256, s));
    unsigned long long *d_ctr = reinterpret_cast<unsigned long long *>(
        reinterpret_cast<unsigned char *>(d_scr) + scr_bytes);
    int32_t rc;
    if (par->x_correlation)
        rc = dg_launch<false, false>(c1, c2, par, pairs, C, d_out_row, d_scr, d_ctr, blocks, s);
    else if (par->r_par_min != 0.)
        rc = dg_launch<true, false>(c1, c2, par, pairs, C, d_out_row, d_scr, d_ctr, blocks, s);
    else
        rc = dg_launch<true, true>(c1, c2, par, pairs, C, d_out_row, d_scr, d_ctr, blocks, s);
    if (rc == 0) {
        const long long total = (long long)n_rows * nb;
        pb2_xi_diag_fold<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d_scr, d_out, n_rows, nb);
        pb2_count_launch(1);
        rc = pb2_check_launch("pb2_xi_diag_fold");
    }
    cudaFreeAsync(d_scr, s);
    return rc;
}
