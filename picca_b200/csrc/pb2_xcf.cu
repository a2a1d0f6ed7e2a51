// Forest x object (quasar) pixel-pair histogram: replaces xcf.compute_xi's loop and
// xcf.compute_xi_forest_pairs_fast (reference py/picca/xcf.py:149-213, 223-322).
//
// Product path (pb2_xi_cross_chunk): for a fixed object the pixels of a forest walk monotonically
// through the (r_par, r_trans) bins in runs of ~7, and every sum of the reference over a run
// factorises into (object constants) x (sums over the run's pixels of w, delta w, r_comov w,
// dist_m w, z w and a count).  Those come from per-forest PREFIX SUMS (pb2_build_prefix), so a
// run costs two record loads instead of a loop over its pixels.  One warp takes a forest; lane l
// first prepares neighbouring object l of a batch of 32 (its constants and pixel window, by
// binary search); then the warp handles the 32 objects in turn, 32 consecutive pixels at a time:
// lane = pixel evaluates the pixel's bin (coalesced loads, same proven-or-exact bins as below), a
// ballot marks where the bin changes, and the last lane of every run adds the run to its bin
// with six native reductions.  The run boundaries are therefore exact, pixel by pixel.
//
// General path (pb2_xi_cross_kernel; every mode, and the validation variant):
// One warp per forest; lane l owns one neighbouring object (its r_comov, dist_m, z, weight and
// cos/sin of half the separation stay in registers) and the warp sweeps the forest's pixels, every
// lane reading the same pixel (broadcast loads).  For a fixed object the pixels walk monotonically
// through the (r_par, r_trans) bins, so each lane accumulates a run of same-bin pixels in
// registers, in factored form (object constants applied at flush time), and flushes it with
// native fp64 global reductions when its bin changes.  Bins are the reference's, bit for bit:
// same sandwich test as the auto kernel, exact re-evaluation with true divisions when in doubt.
#include "pb2_common.cuh"

#define PB2_MAGIC 6755399441055744.0  // 2^52 + 2^51

struct XcfFast {
    double kp_lo, kp_hi, kt_lo, kt_hi, magic;
    int fast;
};

struct XRun {
    double sw, sdw, srp, srt, szw;  // sums over pixels of w1, delta1*w1, r_par*w1, r_trans*w1, z1*w1
    int cnt;
    int key;
};

__device__ __forceinline__ void xrun_flush(XRun &r, double wq, double zq, double *__restrict__ orow,
                                           int nb)
{
    if (r.key >= 0) {
        const double we = wq * r.sw;
        atomic_add_f64(orow + 0 * (size_t)nb + r.key, we);                       // xcf.py:318
        atomic_add_f64(orow + 1 * (size_t)nb + r.key, wq * r.sdw);               // xcf.py:308,317
        atomic_add_f64(orow + 2 * (size_t)nb + r.key, wq * r.srp);               // xcf.py:319
        atomic_add_f64(orow + 3 * (size_t)nb + r.key, wq * r.srt);               // xcf.py:320
        atomic_add_f64(orow + 4 * (size_t)nb + r.key, 0.5 * (wq * r.szw + zq * we));  // :286,:321
        atomic_add_i64(orow + 5 * (size_t)nb + r.key, (long long)r.cnt);
    }
    r.sw = r.sdw = r.srp = r.srt = r.szw = 0.;
    r.cnt = 0;
    r.key = -1;
}

// first index in the non-decreasing a[0..n) with a[i] > v (strict) / a[i] >= v
__device__ __forceinline__ int lane_upper_bound(const double *__restrict__ a, int n, double v,
                                                bool strict)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double x = __ldg(a + mid);
        const bool left = strict ? (x <= v) : (x < v);
        if (left) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

template <bool FAST>
__global__ void __launch_bounds__(256)
pb2_xi_cross_kernel(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, XcfFast F,
                    const int32_t *__restrict__ out_row, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const unsigned np_u = (unsigned)P.num_bins_r_par, nt_u = (unsigned)P.num_bins_r_trans;
    const bool zcut = P.has_z_min_pairs || P.has_z_max_pairs;
    const double magic = F.magic;

    for (long long k = warp; k < pr.n_f1; k += nwarps) {
        const int f1 = pr.f1_index[k];
        const long long a = c1.offset[f1];
        const int n1 = (int)(c1.offset[f1 + 1] - a);
        const long long e0 = pr.nb_offset[k], e1 = pr.nb_offset[k + 1];
        if (n1 == 0 || e1 == e0) continue;  // xcf.py:157
        double *__restrict__ orow = out + (size_t)out_row[k] * 6 * nb;
        const double *__restrict__ p_rc1 = c1.r_comov + a;
        const double *__restrict__ p_dm1 = c1.dist_m + a;
        const double *__restrict__ p_z1 = c1.z + a;
        const double *__restrict__ p_w1 = c1.weights + a;
        const double *__restrict__ p_dw1 = c1.delta_w + a;

        for (long long eb = e0; eb < e1; eb += 32) {
            const long long e = eb + lane;
            const bool have = e < e1;
            double rcq = 0., dmq = 0., zq = 0., wq = 0., ang = 0., ch = 1., sh = 0.;
            if (have) {
                const int f2 = pr.nb_f2[e];
                const long long q = c2.offset[f2];
                rcq = c2.r_comov[q];
                dmq = c2.dist_m[q];
                zq = c2.z[q];
                wq = c2.weights[q];
                ang = pr.nb_ang[e];
                ch = pr.nb_cos[e];
                sh = pr.nb_sin[e];
            }
            const bool vq = have && (wq != 0.);  // xcf.py:283

            // ---- per-lane pixel window (superset), warp sweeps the union
            int ilo = 0, ihi = n1;
            if (FAST) {
                // r_par_min < (rc1 - rcq) ch < r_par_max  and  (dm1 + dmq) sh < r_trans_max
                const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
                const double hi_rc = rcq + P.r_par_max * inv_c;
                const double lo_rc = rcq + P.r_par_min * inv_c;
                ilo = lane_upper_bound(p_rc1, n1, lo_rc - fabs(lo_rc) * 1e-9 - 1e-9, false);
                ihi = lane_upper_bound(p_rc1, n1, hi_rc + fabs(hi_rc) * 1e-9 + 1e-9, true);
                const double tsum = P.r_trans_max * inv_s;
                if (isfinite(tsum))
                    ihi = min(ihi, lane_upper_bound(p_dm1, n1, (tsum - dmq) * (1. + 1e-9) + 1e-9, true));
            }
            if (!vq) {
                ilo = n1;
                ihi = 0;
            }
            int ILO = ilo, IHI = ihi;
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                ILO = min(ILO, __shfl_xor_sync(0xffffffffu, ILO, m));
                IHI = max(IHI, __shfl_xor_sync(0xffffffffu, IHI, m));
            }

            XRun run;
            run.key = -1;
            xrun_flush(run, 0., 0., orow, nb);
            for (int i = ILO; i < IHI; i++) {
                const double rc1 = __ldg(p_rc1 + i), dm1 = __ldg(p_dm1 + i);
                const double w1 = __ldg(p_w1 + i), dw1 = __ldg(p_dw1 + i), z1 = __ldg(p_z1 + i);
                const bool both = vq && (w1 != 0.) && (i >= ilo) && (i < ihi);  // xcf.py:279
                bool in = false;
                int bin = -1;
                double rp = 0., rt = 0.;
                if (FAST) {
                    rp = mul_rn(sub_rn(rc1, rcq), ch);
                    rt = mul_rn(add_rn(dm1, dmq), sh);
                    const double x = sub_rn(rp, P.r_par_min);
                    const int bpl = __double2loint(__fma_rd(x, F.kp_lo, magic));
                    const int bph = __double2loint(__fma_rd(x, F.kp_hi, magic));
                    const int btl = __double2loint(__fma_rd(rt, F.kt_lo, magic));
                    const int bth = __double2loint(__fma_rd(rt, F.kt_hi, magic));
                    // x == 0 exactly is rejected by the reference (r_par <= r_par_min, xcf.py:305)
                    const bool sure = (bpl == bph) && (btl == bth) && (x != 0.);
                    in = both && sure && ((unsigned)bpl < np_u) && ((unsigned)btl < nt_u);
                    bin = btl + (int)nt_u * bpl;
                    if (both && !sure) {
                        PairGeom g = pb2_pair_exact(P, rc1, dm1, rcq, dmq, ang, ch, sh, true, false);
                        in = g.bin >= 0;
                        bin = g.bin;
                    }
                } else if (both) {
                    PairGeom g = pb2_pair_exact(P, rc1, dm1, rcq, dmq, ang, ch, sh, true, false);
                    in = g.bin >= 0;
                    bin = g.bin;
                    rp = g.r_par;
                    rt = g.r_trans;
                }
                if (zcut && in) {
                    const double zm = div_rn(add_rn(z1, zq), 2.);  // xcf.py:286-291
                    if (P.has_z_min_pairs && zm < P.z_min_pairs) in = false;
                    if (P.has_z_max_pairs && zm > P.z_max_pairs) in = false;
                }
                if (in) {
                    if (bin != run.key) {
                        xrun_flush(run, wq, zq, orow, nb);
                        run.key = bin;
                    }
                    run.sw += w1;
                    run.sdw += dw1;
                    run.szw = fma(z1, w1, run.szw);
                    run.srp = fma(rp, w1, run.srp);
                    run.srt = fma(rt, w1, run.srt);
                    run.cnt += 1;
                }
            }
            xrun_flush(run, wq, zq, orow, nb);
        }
    }
}

// ------------------------------------------------------------------------------------------
// prefix records: entry offset[f] + f + i of px_rec = sums over pixels [0, i) of forest f of
// (w, delta w, (r_comov - r_comov[0]) w, (dist_m - dist_m[0]) w, z w, [w != 0]); n + 1 entries.
__global__ void pb2_build_prefix_kernel(pb2_catalog c, double *__restrict__ px)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= c.n_los) return;
    const long long a = c.offset[f];
    const int n = (int)(c.offset[f + 1] - a);
    double *o = px + 6 * (a + f);
    double s0 = 0., s1 = 0., s2 = 0., s3 = 0., s4 = 0., cnt = 0.;
    const double rc0 = n ? c.r_comov[a] : 0., dm0 = n ? c.dist_m[a] : 0.;
    for (int i = 0; i <= n; i++) {
        o[0] = s0; o[1] = s1; o[2] = s2; o[3] = s3; o[4] = s4; o[5] = cnt;
        o += 6;
        if (i < n) {
            const double w = c.weights[a + i];
            s0 += w;
            s1 += c.delta_w[a + i];
            s2 = fma(sub_rn(c.r_comov[a + i], rc0), w, s2);
            s3 = fma(sub_rn(c.dist_m[a + i], dm0), w, s3);
            s4 += c.z_w[a + i];
            cnt += (w != 0.) ? 1. : 0.;
        }
    }
}

// first index in the non-decreasing a[0..n).x (.y) with value > v (strict) / >= v; shared memory
__device__ __forceinline__ int smem_upper_bound(const double2 *a, int n, double v, bool strict,
                                                bool second)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double x = second ? a[mid].y : a[mid].x;
        const bool left = strict ? (x <= v) : (x < v);
        if (left) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// One CTA per forest: its (r_comov, dist_m) pairs and prefix records are staged in shared memory
// once, then the warps take batches of 32 neighbouring objects.
__global__ void __launch_bounds__(256)
pb2_xi_cross_chunk(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, XcfFast F,
                   const int32_t *__restrict__ out_row, double *__restrict__ out,
                   unsigned long long *__restrict__ counter)
{
    extern __shared__ __align__(16) double2 xs[];  // [n1] (rc, dm), then 3 (n1 + 1) prefix halves
    __shared__ unsigned long long s_k;
    __shared__ unsigned s_batch;
    const int lane = threadIdx.x & 31;
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const unsigned np_u = (unsigned)P.num_bins_r_par, nt_u = (unsigned)P.num_bins_r_trans;
    const double magic = F.magic;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_k = atomicAdd(counter, 1ull);
            s_batch = 0;
        }
        __syncthreads();
        const unsigned long long k = s_k;
        if ((long long)k >= pr.n_f1) break;
        const int f1 = pr.f1_index[k];
        const long long a = c1.offset[f1];
        const int n1 = (int)(c1.offset[f1 + 1] - a);
        const long long e0 = pr.nb_offset[k], e1 = pr.nb_offset[k + 1];
        if (n1 == 0 || e1 == e0) continue;  // xcf.py:157
        double *__restrict__ orow = out + (size_t)out_row[k] * 6 * nb;
        double2 *const s_px = xs + n1;
        {
            const double *__restrict__ g_rc = c1.r_comov + a;
            const double *__restrict__ g_dm = c1.dist_m + a;
            const double2 *__restrict__ g_px = reinterpret_cast<const double2 *>(c1.px_rec) + 3 * (a + f1);
            for (int i = threadIdx.x; i < n1; i += blockDim.x) xs[i] = make_double2(__ldg(g_rc + i), __ldg(g_dm + i));
            for (int i = threadIdx.x; i < 3 * (n1 + 1); i += blockDim.x) s_px[i] = __ldg(g_px + i);
        }
        __syncthreads();
        const double rc0 = xs[0].x, dm0 = xs[0].y;

        for (;;) {
            unsigned bidx = 0;
            if (lane == 0) bidx = atomicAdd(&s_batch, 1u);
            bidx = __shfl_sync(0xffffffffu, bidx, 0);
            const long long eb = e0 + 32ll * bidx;
            if (eb >= e1) break;
            // ---- lane l prepares object eb + l: constants and pixel window (a superset)
            const long long e = eb + lane;
            const bool have = e < e1;
            double rcq = 0., dmq = 0., zq = 0., wq = 0., ang = 0., ch = 1., sh = 0.;
            if (have) {
                const int f2 = pr.nb_f2[e];
                const long long q = c2.offset[f2];
                rcq = c2.r_comov[q];
                dmq = c2.dist_m[q];
                zq = c2.z[q];
                wq = c2.weights[q];
                ang = pr.nb_ang[e];
                ch = pr.nb_cos[e];
                sh = pr.nb_sin[e];
            }
            int ilo = n1, ihi = 0;
            if (have && wq != 0.) {  // xcf.py:283
                // r_par_min < (rc1 - rcq) ch < r_par_max  and  (dm1 + dmq) sh < r_trans_max
                const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
                const double hi_rc = rcq + P.r_par_max * inv_c;
                const double lo_rc = rcq + P.r_par_min * inv_c;
                ilo = smem_upper_bound(xs, n1, lo_rc - fabs(lo_rc) * 1e-9 - 1e-9, false, false);
                ihi = smem_upper_bound(xs, n1, hi_rc + fabs(hi_rc) * 1e-9 + 1e-9, true, false);
                const double tsum = P.r_trans_max * inv_s;
                if (isfinite(tsum))
                    ihi = min(ihi, smem_upper_bound(xs, n1, (tsum - dmq) * (1. + 1e-9) + 1e-9, true, true));
            }
            unsigned todo = __ballot_sync(0xffffffffu, ihi > ilo);

            // ---- the warp takes the prepared objects in turn
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const double q_rc = __shfl_sync(0xffffffffu, rcq, src);
                const double q_dm = __shfl_sync(0xffffffffu, dmq, src);
                const double q_z = __shfl_sync(0xffffffffu, zq, src);
                const double q_w = __shfl_sync(0xffffffffu, wq, src);
                const double q_ang = __shfl_sync(0xffffffffu, ang, src);
                const double q_ch = __shfl_sync(0xffffffffu, ch, src);
                const double q_sh = __shfl_sync(0xffffffffu, sh, src);
                const int q_lo = __shfl_sync(0xffffffffu, ilo, src);
                const int q_hi = __shfl_sync(0xffffffffu, ihi, src);
                for (int base = q_lo & ~31; base < q_hi; base += 32) {
                    const int i = base + lane;
                    // ---- bin of pixel i (lane = pixel): -1 = rejected / outside the window
                    int key = -1;
                    if (i >= q_lo && i < q_hi) {
                        const double2 p1 = xs[i];
                        const double rp = mul_rn(sub_rn(p1.x, q_rc), q_ch);
                        const double rt = mul_rn(add_rn(p1.y, q_dm), q_sh);
                        const double x = sub_rn(rp, P.r_par_min);
                        const int bpl = __double2loint(__fma_rd(x, F.kp_lo, magic));
                        const int bph = __double2loint(__fma_rd(x, F.kp_hi, magic));
                        const int btl = __double2loint(__fma_rd(rt, F.kt_lo, magic));
                        const int bth = __double2loint(__fma_rd(rt, F.kt_hi, magic));
                        // x == 0 exactly is rejected by the reference (r_par <= r_par_min, xcf.py:305)
                        if ((bpl == bph) && (btl == bth) && (x != 0.)) {
                            if (((unsigned)bpl < np_u) && ((unsigned)btl < nt_u)) key = btl + (int)nt_u * bpl;
                        } else {
                            key = pb2_pair_exact(P, p1.x, p1.y, q_rc, q_dm, q_ang, q_ch, q_sh, true, false).bin;
                        }
                    }
                    // ---- runs of equal bins inside the chunk: a lane starts a run when its bin
                    // differs from its left neighbour's; the last lane of a run adds it up
                    const int left = __shfl_up_sync(0xffffffffu, key, 1);
                    const bool starts = (lane == 0) || (key != left);
                    const unsigned smask = __ballot_sync(0xffffffffu, starts);
                    const bool last = (lane == 31) || ((smask >> ((lane + 1) & 31)) & 1u);
                    if (last && key >= 0) {
                        const int first = 31 - __clz(smask & (0xffffffffu >> (31 - lane)));
                        const double2 *pa = s_px + 3 * (base + first);
                        const double2 *pb = s_px + 3 * (i + 1);
                        const double2 a01 = pa[0], a23 = pa[1], a45 = pa[2];
                        const double2 b01 = pb[0], b23 = pb[1], b45 = pb[2];
                        const int cnt = (int)(b45.y - a45.y);
                        if (cnt > 0) {
                            const double sw = b01.x - a01.x, sdw = b01.y - a01.y;
                            const double src2 = b23.x - a23.x, sdm = b23.y - a23.y, szw = b45.x - a45.x;
                            const double we = q_w * sw;
                            double *dst = orow + key;
                            atomic_add_f64(dst, we);                                       // xcf.py:318
                            atomic_add_f64(dst + (size_t)nb, q_w * sdw);                   // :308,:317
                            atomic_add_f64(dst + 2 * (size_t)nb,
                                           q_w * q_ch * fma(rc0 - q_rc, sw, src2));        // :319
                            atomic_add_f64(dst + 3 * (size_t)nb,
                                           q_w * q_sh * fma(dm0 + q_dm, sw, sdm));         // :320
                            atomic_add_f64(dst + 4 * (size_t)nb, 0.5 * (q_w * szw + q_z * we));  // :286,:321
                            atomic_add_i64(dst + 5 * (size_t)nb, (long long)cnt);
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// pb2_xi_cross_chunk_t: the same runs, but the reductions of a step are TRANSPOSED.  fp64 global
// reductions cost LSU time per (instruction x distinct sector), not per lane
// (scripts/micro/red_coalesce.cu: five lanes issuing six REDs each sustain 3.2e10 runs/s on a
// B200, thirty lanes issuing ONE instruction, six adjacent lanes per 64-byte line, 8.3e10), and
// the kernel above keeps the L1 data pipe 89 % busy with exactly that.  Here the last lane of a run
// only publishes (bin, first pixel, end pixel) in a per-warp slot; then lane 6 g + m forms
// component m of run g -- prefix[end][m] - prefix[first][m], scaled by the two per-object
// constants of that component -- and ONE reduction instruction adds five runs into a bin-major
// scratch histogram [row][bin][8] (slot 5 = the pair count, as a double: exact).  A run that
// crosses a 32-pixel step stays open (warp-uniform carry), so runs are no longer cut at step
// boundaries.  pb2_xi_cross_fold adds the scratch into the caller's [row][6][bin] blocks.
__device__ __forceinline__ void xcf_red(double *p, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// (a minimum-blocks launch bound makes ptxas produce 20 % slower code here, at 3, 4 and 5 CTAs per SM
// alike: profiles/r02_xcf_variants.log; three CTAs are resident at 77 registers)
#define XCF_T_GRID 4
__global__ void __launch_bounds__(256)
pb2_xi_cross_chunk_t(pb2_catalog c1, pb2_catalog c2, pb2_params P, pb2_pairs pr, XcfFast F,
                     const int32_t *__restrict__ out_row, double *__restrict__ scr,
                     unsigned long long *__restrict__ counter)
{
    extern __shared__ __align__(16) double2 xs[];  // [n1] (rc, dm), then 3 (n1 + 1) prefix halves
    __shared__ unsigned long long s_k;
    __shared__ unsigned s_batch;
    __shared__ int2 s_run[8][33];   // per warp: (bin, first pixel << 16 | end pixel) of a run
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int nb = P.num_bins_r_par * P.num_bins_r_trans;
    const unsigned np_u = (unsigned)P.num_bins_r_par, nt_u = (unsigned)P.num_bins_r_trans;
    const double magic = F.magic;
    const int g_of = lane / 6, m_of = lane - 6 * g_of;   // emission role: run g, component m
    int2 *const runs = s_run[wid];

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_k = atomicAdd(counter, 1ull);
            s_batch = 0;
        }
        __syncthreads();
        const unsigned long long k = s_k;
        if ((long long)k >= pr.n_f1) break;
        const int f1 = pr.f1_index[k];
        const long long a = c1.offset[f1];
        const int n1 = (int)(c1.offset[f1 + 1] - a);
        const long long e0 = pr.nb_offset[k], e1 = pr.nb_offset[k + 1];
        if (n1 == 0 || e1 == e0) continue;  // xcf.py:157
        double *__restrict__ srow = scr + (size_t)out_row[k] * nb * 8;
        double2 *const s_px = xs + n1;
        const double *const s_pxd = reinterpret_cast<const double *>(s_px);
        {
            const double *__restrict__ g_rc = c1.r_comov + a;
            const double *__restrict__ g_dm = c1.dist_m + a;
            const double2 *__restrict__ g_px = reinterpret_cast<const double2 *>(c1.px_rec) + 3 * (a + f1);
            for (int i = threadIdx.x; i < n1; i += blockDim.x) xs[i] = make_double2(__ldg(g_rc + i), __ldg(g_dm + i));
            for (int i = threadIdx.x; i < 3 * (n1 + 1); i += blockDim.x) s_px[i] = __ldg(g_px + i);
        }
        __syncthreads();
        const double rc0 = xs[0].x, dm0 = xs[0].y;
        // objects per claim: about eight claims per warp, so that the eight warps of the CTA
        // finish a forest together (config 3: ~900 neighbouring objects per forest; ncu showed
        // 15 % of the stall samples at the CTA barrier with claims of 32)
        const int bsz = (int)min(32ll, max(4ll, (e1 - e0 + 63) / 64));

        for (;;) {
            unsigned bidx = 0;
            if (lane == 0) bidx = atomicAdd(&s_batch, 1u);
            bidx = __shfl_sync(0xffffffffu, bidx, 0);
            const long long eb = e0 + (long long)bsz * bidx;
            if (eb >= e1) break;
            // ---- lane l prepares object eb + l: constants and pixel window (a superset)
            const long long e = eb + lane;
            const bool have = lane < bsz && e < e1;
            double rcq = 0., dmq = 0., zq = 0., wq = 0., ang = 0., ch = 1., sh = 0.;
            if (have) {
                const int f2 = pr.nb_f2[e];
                const long long q = c2.offset[f2];
                rcq = c2.r_comov[q];
                dmq = c2.dist_m[q];
                zq = c2.z[q];
                wq = c2.weights[q];
                ang = pr.nb_ang[e];
                ch = pr.nb_cos[e];
                sh = pr.nb_sin[e];
            }
            int ilo = n1, ihi = 0;
            if (have && wq != 0.) {  // xcf.py:283
                const double inv_c = 1.0 / ch, inv_s = 1.0 / sh;
                const double hi_rc = rcq + P.r_par_max * inv_c;
                const double lo_rc = rcq + P.r_par_min * inv_c;
                ilo = smem_upper_bound(xs, n1, lo_rc - fabs(lo_rc) * 1e-9 - 1e-9, false, false);
                ihi = smem_upper_bound(xs, n1, hi_rc + fabs(hi_rc) * 1e-9 + 1e-9, true, false);
                const double tsum = P.r_trans_max * inv_s;
                if (isfinite(tsum))
                    ihi = min(ihi, smem_upper_bound(xs, n1, (tsum - dmq) * (1. + 1e-9) + 1e-9, true, true));
            }
            unsigned todo = __ballot_sync(0xffffffffu, ihi > ilo);

            // ---- the warp takes the prepared objects in turn
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const double q_rc = __shfl_sync(0xffffffffu, rcq, src);
                const double q_dm = __shfl_sync(0xffffffffu, dmq, src);
                const double q_z = __shfl_sync(0xffffffffu, zq, src);
                const double q_w = __shfl_sync(0xffffffffu, wq, src);
                const double q_ang = __shfl_sync(0xffffffffu, ang, src);
                const double q_ch = __shfl_sync(0xffffffffu, ch, src);
                const double q_sh = __shfl_sync(0xffffffffu, sh, src);
                const int q_lo = __shfl_sync(0xffffffffu, ilo, src);
                const int q_hi = __shfl_sync(0xffffffffu, ihi, src);
                // component m of a run = A (prefix_m[end] - prefix_m[first]) + B (sum of w):
                // xcf.py:318 (m = 0), :308,:317 (1), :319 (2), :320 (3), :286,:321 (4), count (5)
                double cA = q_w, cB = 0.;
                if (m_of == 2) { cA = q_w * q_ch; cB = cA * (rc0 - q_rc); }
                if (m_of == 3) { cA = q_w * q_sh; cB = cA * (dm0 + q_dm); }
                if (m_of == 4) { cA = 0.5 * q_w; cB = 0.5 * (q_z * q_w); }
                if (m_of == 5) cA = 1.;
                int open_key = -2;    // bin of the run that is open at the start of a step
                int open_first = 0;   // ... and its first pixel (warp-uniform)
                for (int base = q_lo & ~31; base < q_hi; base += 32) {
                    const int i = base + lane;
                    // ---- bin of pixel i (lane = pixel): -1 = rejected / outside the window
                    int key = -1;
                    if (i >= q_lo && i < q_hi) {
                        const double2 p1 = xs[i];
                        const double rp = mul_rn(sub_rn(p1.x, q_rc), q_ch);
                        const double rt = mul_rn(add_rn(p1.y, q_dm), q_sh);
                        const double x = sub_rn(rp, P.r_par_min);
                        const int bpl = __double2loint(__fma_rd(x, F.kp_lo, magic));
                        const int bph = __double2loint(__fma_rd(x, F.kp_hi, magic));
                        const int btl = __double2loint(__fma_rd(rt, F.kt_lo, magic));
                        const int bth = __double2loint(__fma_rd(rt, F.kt_hi, magic));
                        // x == 0 exactly is rejected by the reference (r_par <= r_par_min, xcf.py:305)
                        if ((bpl == bph) && (btl == bth) && (x != 0.)) {
                            if (((unsigned)bpl < np_u) && ((unsigned)btl < nt_u)) key = btl + (int)nt_u * bpl;
                        } else {
                            key = pb2_pair_exact(P, p1.x, p1.y, q_rc, q_dm, q_ang, q_ch, q_sh, true, false).bin;
                        }
                    }
                    // ---- runs of equal bins: a lane starts a run when its bin differs from its
                    // left neighbour's (lane 0: from the run left open by the previous step); the
                    // last lane of a run publishes it, except the run still open at lane 31
                    int left = __shfl_up_sync(0xffffffffu, key, 1);
                    if (lane == 0) left = open_key;
                    const bool starts = key != left;
                    const unsigned smask = __ballot_sync(0xffffffffu, starts);
                    const bool final_step = base + 32 >= q_hi;
                    const bool last = lane == 31 ? final_step : ((smask >> (lane + 1)) & 1u);
                    const unsigned emask = __ballot_sync(0xffffffffu, last && key >= 0);
                    const unsigned mine = smask & (0xffffffffu >> (31 - lane));   // starts at or left of me
                    const int first = mine ? base + 31 - __clz(mine) : open_first;
                    if (last && key >= 0)
                        runs[__popc(emask & ((1u << lane) - 1u))] = make_int2(key, (first << 16) | (i + 1));
                    // the run left open by the previous step ended at its last pixel when lane 0
                    // starts another one: no lane of this step is its last, lane 0 publishes it
                    const bool carried_ends = (smask & 1u) && open_key >= 0;   // warp-uniform
                    if (carried_ends && lane == 0)
                        runs[__popc(emask)] = make_int2(open_key, (open_first << 16) | base);
                    // the run open after this step
                    open_key = __shfl_sync(0xffffffffu, key, 31);
                    open_first = __shfl_sync(0xffffffffu, first, 31);
                    __syncwarp();
                    const int nrun = __popc(emask) + (carried_ends ? 1 : 0);
                    for (int r0 = 0; r0 < nrun; r0 += 5) {
                        if (g_of < 5 && r0 + g_of < nrun) {
                            const int2 rr = runs[r0 + g_of];
                            const int pa = 6 * (rr.y >> 16), pb = 6 * (rr.y & 0xffff);
                            const double dm_ = s_pxd[pb + m_of] - s_pxd[pa + m_of];
                            const double d0 = s_pxd[pb] - s_pxd[pa];
                            xcf_red(srow + (size_t)rr.x * 8 + m_of, fma(cA, dm_, cB * d0));
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }
}

// scratch [row][bin][8] (slot 5: the pair count as a double) -> the caller's [row][6][bin]
__global__ void pb2_xi_cross_fold(const double *__restrict__ scr, double *__restrict__ out,
                                  long long n_rows, int nb)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * nb) return;
    const long long row = idx / nb;
    const int bin = (int)(idx - row * nb);
    const double2 *src = reinterpret_cast<const double2 *>(scr + idx * 8);
    const double2 v01 = src[0], v23 = src[1], v45 = src[2];
    if (v45.y == 0.) return;   // no pair in this bin
    double *dst = out + row * 6 * (long long)nb + bin;
    dst[0] += v01.x;
    dst[(size_t)nb] += v01.y;
    dst[2 * (size_t)nb] += v23.x;
    dst[3 * (size_t)nb] += v23.y;
    dst[4 * (size_t)nb] += v45.x;
    long long *cnt = reinterpret_cast<long long *>(dst + 5 * (size_t)nb);
    *cnt += (long long)v45.y;
}

extern "C" {

int32_t pb2_xi_cross(const pb2_catalog *cat1, const pb2_catalog *objs, const pb2_params *par,
                     const pb2_pairs *pairs, const int32_t *d_out_row, int64_t n_rows,
                     double *d_out, int32_t variant, void *stream)
{
    if (!cat1 || !objs || !par || !pairs || !d_out_row || !d_out) {
        pb2_set_error("pb2_xi_cross: null pointer argument");
        return PB2_EINVAL;
    }
    if (pairs->n_pairs <= 0 || pairs->n_f1 <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    XcfFast F;
    const double kp = (double)par->num_bins_r_par / (par->r_par_max - par->r_par_min);
    const double kt = (double)par->num_bins_r_trans / par->r_trans_max;
    const double eps = 9.094947017729282e-13;  // 2^-40
    F.kp_lo = kp * (1. - eps);
    F.kp_hi = kp * (1. + eps);
    F.kt_lo = kt * (1. - eps);
    F.kt_hi = kt * (1. + eps);
    F.magic = PB2_MAGIC;
    // variant 1 = validation: every pair through the reference expression, no windows
    F.fast = (variant != 1 && !par->rmu_binning && !par->ang_correlation && cat1->sorted &&
              par->num_bins_r_par <= 4096 && par->num_bins_r_trans <= 4096 &&
              par->r_par_max > par->r_par_min && par->r_trans_max > 0.) ? 1 : 0;
    long long blocks = (pairs->n_f1 + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    pb2_timing_begin(s);
    // variant 0: prefix-sum kernel when the catalogue carries prefix records and no per-pair
    // z cut is set (the cut is per pixel pair, xcf.py:286-291); variant 2 forces the lane = object
    // kernel
    const size_t chunk_smem = (size_t)64 * (cat1->max_pix + 1);
    const bool chunk_ok = F.fast && cat1->px_rec && !par->has_z_min_pairs && !par->has_z_max_pairs &&
                          chunk_smem <= 100 * 1024;
    // variant 0: transposed reductions into a bin-major scratch (pixel indices of a run travel in
    // 16 bits; the scratch is 64 bytes per bin and HEALPix row); variant 4: the per-lane reductions
    const double scr_bytes_d = (double)n_rows * par->num_bins_r_par * par->num_bins_r_trans * 64.;
    if (variant == 0 && chunk_ok && cat1->max_pix < 65535 && n_rows > 0 && scr_bytes_d < 16e9) {
        const int nb = par->num_bins_r_par * par->num_bins_r_trans;
        const size_t scr_bytes = (size_t)n_rows * nb * 8 * sizeof(double);
        int dev = 0;
        PB2_CUDA(cudaGetDevice(&dev));
        {
            cudaMemPool_t pool;   // keep the stream-ordered pool's memory across calls
            PB2_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
            uint64_t keep = UINT64_MAX;
            PB2_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        double *d_scr = nullptr;
        PB2_CUDA(cudaMallocAsync((void **)&d_scr, scr_bytes + 256, s));
        PB2_CUDA(cudaMemsetAsync(d_scr, 0, scr_bytes + 256, s));
        unsigned long long *d_ctr = reinterpret_cast<unsigned long long *>(
            reinterpret_cast<unsigned char *>(d_scr) + scr_bytes);
        PB2_CUDA(cudaFuncSetAttribute(pb2_xi_cross_chunk_t, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)chunk_smem));
        long long ctas = pairs->n_f1 < 148 * XCF_T_GRID ? pairs->n_f1 : 148 * XCF_T_GRID;
        pb2_xi_cross_chunk_t<<<(unsigned)ctas, 256, chunk_smem, s>>>(*cat1, *objs, *par, *pairs, F,
                                                                     d_out_row, d_scr, d_ctr);
        pb2_count_launch(1);
        int32_t rc2 = pb2_check_launch("pb2_xi_cross_chunk_t");
        pb2_timing_end(s);
        if (rc2 == 0) {
            const long long total = (long long)n_rows * nb;
            pb2_xi_cross_fold<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d_scr, d_out, n_rows, nb);
            pb2_count_launch(1);
            rc2 = pb2_check_launch("pb2_xi_cross_fold");
        }
        cudaFreeAsync(d_scr, s);
        return rc2;
    }
    if ((variant == 0 || variant == 4) && chunk_ok) {
        // forest-claim counter: allocated in stream order for this call only (concurrent calls
        // on other streams / threads each get their own)
        unsigned long long *d_ctr = nullptr;
        PB2_CUDA(cudaMallocAsync((void **)&d_ctr, sizeof(unsigned long long), s));
        PB2_CUDA(cudaMemsetAsync(d_ctr, 0, sizeof(unsigned long long), s));
        PB2_CUDA(cudaFuncSetAttribute(pb2_xi_cross_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)chunk_smem));
        long long ctas = pairs->n_f1 < 148 * 4 ? pairs->n_f1 : 148 * 4;
        pb2_xi_cross_chunk<<<(unsigned)ctas, 256, chunk_smem, s>>>(*cat1, *objs, *par, *pairs, F,
                                                                   d_out_row, d_out, d_ctr);
        pb2_count_launch(1);
        int32_t rc2 = pb2_check_launch("pb2_xi_cross_chunk");
        pb2_timing_end(s);
        cudaFreeAsync(d_ctr, s);
        return rc2;
    }
    if (F.fast)
        pb2_xi_cross_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(*cat1, *objs, *par, *pairs, F,
                                                                    d_out_row, d_out);
    else
        pb2_xi_cross_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(*cat1, *objs, *par, *pairs, F,
                                                                     d_out_row, d_out);
    pb2_count_launch(1);
    int32_t rc = pb2_check_launch("pb2_xi_cross_kernel");
    pb2_timing_end(s);
    return rc;
}

int32_t pb2_build_prefix(const pb2_catalog *cat, double *d_px_rec, void *stream)
{
    if (!cat || !d_px_rec) {
        pb2_set_error("pb2_build_prefix: null pointer argument");
        return PB2_EINVAL;
    }
    if (cat->n_los <= 0) return 0;
    pb2_build_prefix_kernel<<<(unsigned)((cat->n_los + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        *cat, d_px_rec);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_build_prefix");
}

}  // extern "C"
