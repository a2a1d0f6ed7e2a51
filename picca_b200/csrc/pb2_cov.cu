// Sub-sample covariance of the per-HEALPix correlation blocks and its smoothing -- the consumer of
// the [n_healpix][nb] WE/DA blocks the pair kernels produce.
//
//   pb2_cov_subsample : utils.compute_cov   (reference py/picca/utils.py:100-128)
//   pb2_cov_smooth    : utils.smooth_cov    (reference py/picca/utils.py:153-249)
//
// compute_cov is  C = M^T M / (W W^T),  M[s,i] = w[s,i] (xi[s,i] - <xi>_i),  W_i = sum_s w[s,i]:
// an fp64 rank-n_s update of an nb x nb matrix (2500^2 x 1400 x 2 = 1.8e10 flop at config 2).
// B200 runs FP64 through the ordinary DFMA pipe (the fp64 tensor path is no faster on sm_100a), so
// this is a register-tiled DFMA kernel:
//   * pb2_cov_colstats: one thread per bin adds the sub-samples in order -- the same association
//     as NumPy's axis-0 reduction, so <xi> and W are bit-equal to the reference's;
//     pb2_cov_build_m writes M once into a zero-padded scratch [ks][ld] (ld, ks multiples of the
//     tile sizes: no bounds checks anywhere in the contraction);
//   * pb2_cov_syrk: upper-triangular 64x64 tiles, 128 threads x (8 x 4) accumulators, K chunks of
//     16 sub-samples staged by TMA bulk copies (cp.async.bulk + mbarrier, 3 stages, one producer
//     thread), epilogue divides by W_i W_j where positive and writes the tile and its mirror.
// smooth_cov averages the correlation coefficient over bin pairs with equal
// (round(|dr_par|/delta), round(|dr_trans|/delta)) -- a Python double loop over nb^2/2 pairs in the
// reference.  Here: one thread per (i, j > i), native red.global.add into an L2-resident key table,
// then a second kernel writes the smoothed matrix.
#include "pb2_common.cuh"

#define CV_TILE 64
#define CV_KC 16
#define CV_STAGES 3
#define CV_THREADS 128
#define CV_STAGE_BYTES (2 * CV_KC * CV_TILE * 8)

__device__ __forceinline__ unsigned cv_saddr(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cv_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cv_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void cv_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void cv_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CV_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CV_DONE;\n"
        "bra CV_WAIT;\n"
        "CV_DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// ---- column statistics (utils.py:113-116): one thread per bin, the sub-samples added in order
// like NumPy's axis-0 reduction (bit-equal means and weight sums).  The loads of 16 sub-samples are
// issued together and one batch ahead, only the additions are sequential.
#define CV_BATCH 16
__global__ void pb2_cov_colstats(int n_s, int nb, const double *__restrict__ xi,
                                 const double *__restrict__ we, double *__restrict__ mean_xi,
                                 double *__restrict__ sum_w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    double sx = 0., sw = 0.;
    int s = 0;
    // software pipeline: the loads of batch k+1 are in flight while batch k is added
    double x[CV_BATCH], w[CV_BATCH], xn[CV_BATCH], wn[CV_BATCH];
    const int n_full = n_s / CV_BATCH;
    if (n_full > 0) {
#pragma unroll
        for (int k = 0; k < CV_BATCH; ++k) {
            x[k] = __ldg(xi + (size_t)k * nb + i);
            w[k] = __ldg(we + (size_t)k * nb + i);
        }
    }
    for (int b = 0; b < n_full; ++b, s += CV_BATCH) {
        if (b + 1 < n_full) {
#pragma unroll
            for (int k = 0; k < CV_BATCH; ++k) {
                xn[k] = __ldg(xi + (size_t)(s + CV_BATCH + k) * nb + i);
                wn[k] = __ldg(we + (size_t)(s + CV_BATCH + k) * nb + i);
            }
        }
#pragma unroll
        for (int k = 0; k < CV_BATCH; ++k) {
            sx = add_rn(sx, mul_rn(x[k], w[k]));  // one rounded product per term, no FMA
            sw = add_rn(sw, w[k]);
        }
#pragma unroll
        for (int k = 0; k < CV_BATCH; ++k) x[k] = xn[k], w[k] = wn[k];
    }
    for (; s < n_s; ++s) {
        const double w = we[(size_t)s * nb + i];
        sx = add_rn(sx, mul_rn(xi[(size_t)s * nb + i], w));
        sw = add_rn(sw, w);
    }
    if (sw > 0.) sx = div_rn(sx, sw);  // mean_xi[w] /= sum_weights[w]
    mean_xi[i] = sx;
    sum_w[i] = sw;
}

// ---- M[s][i] = weights * (xi - mean_xi) (utils.py:118) into the zero-padded scratch [ks][ld]
__global__ void pb2_cov_build_m(int n_s, int nb, int ld, const double *__restrict__ xi,
                                const double *__restrict__ we, const double *__restrict__ mean_xi,
                                double *__restrict__ M)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (i >= ld) return;
    double v = 0.;
    if (i < nb && s < n_s)
        v = mul_rn(we[(size_t)s * nb + i], sub_rn(xi[(size_t)s * nb + i], mean_xi[i]));
    M[(size_t)s * ld + i] = v;
}

// ---- one 64x64 tile of A^T B: A is [ks][lda], B is [ks][ldb] (K-major, zero padded), K chunks of
// 16 rows staged by TMA bulk copies into a 3-stage ring.  Thread (ty, tx) ends with
// acc[a][b] = sum_k A[k][i0 + ty*8 + a] * B[k][j0 + (b>>1)*32 + tx*2 + (b&1)].
__device__ __forceinline__ void cv_tile_atb(const double *__restrict__ gA, size_t lda,
                                            const double *__restrict__ gB, size_t ldb, int n_chunks,
                                            unsigned char *smem, unsigned long long *s_bar,
                                            double (&acc)[8][4])
{
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < CV_STAGES; ++s) cv_mbar_init(cv_saddr(&s_bar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned smem0 = cv_saddr(smem);

    auto issue = [&](int chunk) {  // one thread: 2 x 16 rows of 512 bytes
        const int st = chunk % CV_STAGES;
        const unsigned bar = cv_saddr(&s_bar[st]);
        const unsigned dst = smem0 + st * CV_STAGE_BYTES;
        cv_mbar_expect_tx(bar, CV_STAGE_BYTES);
        const size_t row0 = (size_t)chunk * CV_KC;
#pragma unroll
        for (int k = 0; k < CV_KC; ++k) {
            cv_bulk_g2s(dst + k * CV_TILE * 8, gA + (row0 + k) * lda, CV_TILE * 8, bar);
            cv_bulk_g2s(dst + (CV_KC + k) * CV_TILE * 8, gB + (row0 + k) * ldb, CV_TILE * 8, bar);
        }
    };
    if (tid == 0)
        for (int c = 0; c < CV_STAGES && c < n_chunks; ++c) issue(c);

    // thread (ty, tx): rows ty*8 .. ty*8+7 of the A tile; columns tx*2, tx*2+1, 32+tx*2, 32+tx*2+1
    // of the B tile (16-byte stride between lanes: conflict-free LDS.128)
    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.;

    for (int c = 0; c < n_chunks; ++c) {
        const int st = c % CV_STAGES;
        cv_mbar_wait(cv_saddr(&s_bar[st]), (unsigned)((c / CV_STAGES) & 1));
        const double *As = reinterpret_cast<const double *>(smem + st * CV_STAGE_BYTES);
        const double *Bs = As + CV_KC * CV_TILE;
#pragma unroll
        for (int k = 0; k < CV_KC; ++k) {
            double av[8], bv[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const double2 v = *reinterpret_cast<const double2 *>(As + k * CV_TILE + ty * 8 + 2 * m);
                av[2 * m] = v.x;
                av[2 * m + 1] = v.y;
            }
            const double2 b0 = *reinterpret_cast<const double2 *>(Bs + k * CV_TILE + tx * 2);
            const double2 b1 = *reinterpret_cast<const double2 *>(Bs + k * CV_TILE + 32 + tx * 2);
            bv[0] = b0.x, bv[1] = b0.y, bv[2] = b1.x, bv[3] = b1.y;
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();  // every thread is done with stage st: refill it
        if (tid == 0 && c + CV_STAGES < n_chunks) issue(c + CV_STAGES);
    }
}

// ---- C = M^T M, upper-triangular tiles + mirror.  sum_w != NULL: C[w] /= (W W^T)[w]
// (utils.py:123-126); sum_w == NULL: C *= scale (np.cov's `c *= 1/(N-1)`).
__global__ void __launch_bounds__(CV_THREADS)
pb2_cov_syrk(int nb, int ld, int ks, int n_tiles, const double *__restrict__ M,
             const double *__restrict__ sum_w, double scale, double *__restrict__ cov)
{
    extern __shared__ __align__(128) unsigned char cv_smem[];
    __shared__ __align__(8) unsigned long long s_bar[CV_STAGES];
    // linear CTA index -> (bi <= bj)
    int t = blockIdx.x, bi = 0;
    while (t >= n_tiles - bi) {
        t -= n_tiles - bi;
        ++bi;
    }
    const int bj = bi + t;
    double acc[8][4];
    cv_tile_atb(M + (size_t)bi * CV_TILE, ld, M + (size_t)bj * CV_TILE, ld, ks / CV_KC, cv_smem,
                s_bar, acc);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int i = bi * CV_TILE + ty * 8 + a;
        if (i >= nb) continue;
        const double wi = sum_w ? sum_w[i] : 0.;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = bj * CV_TILE + (b >> 1) * 32 + tx * 2 + (b & 1);
            if (j >= nb) continue;
            double v = acc[a][b];
            if (sum_w) {
                const double den = mul_rn(sum_w[j], wi);
                if (den > 0.) v = div_rn(v, den);
            } else {
                v = mul_rn(v, scale);
            }
            cov[(size_t)i * nb + j] = v;
            if (bi != bj) cov[(size_t)j * nb + i] = v;
        }
    }
}

// ---- bootstrap covariance (utils.compute_cov_boot, utils.py:131-150)
// out[b][i] (leading dimension ldo) = sum_s A[s][b] B[s][i];  divide != 0: out = out_prev / that
__global__ void __launch_bounds__(CV_THREADS)
pb2_cov_gemm_atb(int m, int n, int lda, int ldb, int ldo, int ks, const double *__restrict__ A,
                 const double *__restrict__ B, int divide, double *__restrict__ out)
{
    extern __shared__ __align__(128) unsigned char cv_smem[];
    __shared__ __align__(8) unsigned long long s_bar[CV_STAGES];
    const int bi = blockIdx.y, bj = blockIdx.x;
    double acc[8][4];
    cv_tile_atb(A + (size_t)bi * CV_TILE, lda, B + (size_t)bj * CV_TILE, ldb, ks / CV_KC, cv_smem,
                s_bar, acc);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int i = bi * CV_TILE + ty * 8 + a;
        if (i >= m) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = bj * CV_TILE + (b >> 1) * 32 + tx * 2 + (b & 1);
            if (j >= n) continue;
            double *dst = out + (size_t)i * ldo + j;
            *dst = divide ? div_rn(*dst, acc[a][b]) : acc[a][b];
        }
    }
}

// multiplicity of sub-sample s in bootstrap realisation b: cnt[s][b] (leading dimension ldc)
__global__ void pb2_cov_boot_counts(long long total, int n_s, int ldc, const int *__restrict__ idx,
                                    double *__restrict__ cnt)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int b = (int)(t / n_s);
    atomicAdd(cnt + (size_t)idx[t] * ldc + b, 1.0);
}

// P[s][i] = weights * xi, Wm[s][i] = weights, zero padded [ks][ld]
__global__ void pb2_cov_boot_operands(int n_s, int nb, int ld, const double *__restrict__ xi,
                                      const double *__restrict__ we, double *__restrict__ P,
                                      double *__restrict__ Wm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (i >= ld) return;
    double p = 0., w = 0.;
    if (i < nb && s < n_s) {
        w = we[(size_t)s * nb + i];
        p = mul_rn(w, xi[(size_t)s * nb + i]);
    }
    P[(size_t)s * ld + i] = p;
    Wm[(size_t)s * ld + i] = w;
}

// np.cov: X -= X.mean over the realisations, in place (boot is [kb][ld], rows >= n_boot are zero).
// Block = 32 bins x 16 row groups; row group y sums rows y, y+16, ...
__global__ void pb2_cov_boot_center(int n_boot, int nb, int ld, double *__restrict__ boot)
{
    __shared__ double part[16][33];
    const int i = blockIdx.x * 32 + threadIdx.x;
    const int y = threadIdx.y;
    double acc = 0.;
    if (i < nb)
        for (int b = y; b < n_boot; b += 16) acc += boot[(size_t)b * ld + i];
    part[y][threadIdx.x] = acc;
    __syncthreads();
    double tot = 0.;
#pragma unroll
    for (int k = 0; k < 16; ++k) tot += part[k][threadIdx.x];
    const double mean = tot / (double)n_boot;
    if (i < nb)
        for (int b = y; b < n_boot; b += 16) boot[(size_t)b * ld + i] -= mean;
}

// ---- smooth_cov (utils.py:185-247)
struct CvSmooth {
    int nb, per_r_par, n_dp, n_dt, rp_lo, n_rp;
    double delta_r_par, delta_r_trans;
};

__device__ __forceinline__ int cv_key(const CvSmooth &S, double rp_i, double rt_i, double rp_j,
                                      double rt_j)
{
    // round() of a Python float is round-half-to-even == rint
    const int dp = (int)rint(div_rn(fabs(sub_rn(rp_j, rp_i)), S.delta_r_par));
    const int dt = (int)rint(div_rn(fabs(sub_rn(rt_i, rt_j)), S.delta_r_trans));
    if (dp < 0 || dp >= S.n_dp || dt < 0 || dt >= S.n_dt) return -1;
    int key = dp * S.n_dt + dt;
    if (S.per_r_par) {
        const int irp = (int)div_rn(rp_i, S.delta_r_par) - S.rp_lo;  // int(): towards zero
        if (irp < 0 || irp >= S.n_rp) return -1;
        key += irp * S.n_dp * S.n_dt;
    }
    return key;
}

__global__ void pb2_cov_smooth_accumulate(CvSmooth S, const double *__restrict__ cov,
                                          const double *__restrict__ r_par,
                                          const double *__restrict__ r_trans,
                                          double *__restrict__ tab_sum,
                                          unsigned long long *__restrict__ tab_cnt,
                                          int *__restrict__ bad)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= S.nb || j <= i) return;
    const double var_i = cov[(size_t)i * S.nb + i], var_j = cov[(size_t)j * S.nb + j];
    // correlation = covariance / np.sqrt(var * var[:, None]) (utils.py:192)
    const double corr = div_rn(cov[(size_t)i * S.nb + j], sqrt(mul_rn(var_j, var_i)));
    const int key = cv_key(S, r_par[i], r_trans[i], r_par[j], r_trans[j]);
    if (key < 0) {
        atomicExch(bad, 1);
        return;
    }
    atomicAdd(tab_sum + key, corr);
    atomicAdd(tab_cnt + key, 1ull);
}

__global__ void pb2_cov_smooth_apply(CvSmooth S, const double *__restrict__ cov,
                                     const double *__restrict__ r_par,
                                     const double *__restrict__ r_trans,
                                     const double *__restrict__ tab_sum,
                                     const unsigned long long *__restrict__ tab_cnt,
                                     double *__restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= S.nb || j < i) return;
    const double var_i = cov[(size_t)i * S.nb + i], var_j = cov[(size_t)j * S.nb + j];
    const double scale = sqrt(mul_rn(var_j, var_i));  // np.sqrt(var * var[:, None]) (utils.py:246)
    if (j == i) {
        out[(size_t)i * S.nb + i] = scale;  // correlation_smooth[index, index] = 1.
        return;
    }
    const int key = cv_key(S, r_par[i], r_trans[i], r_par[j], r_trans[j]);
    double v = 0.;
    if (key >= 0) v = mul_rn(div_rn(tab_sum[key], (double)tab_cnt[key]), scale);
    out[(size_t)i * S.nb + j] = v;
    out[(size_t)j * S.nb + i] = v;
}

extern "C" {

int64_t pb2_cov_scratch_bytes(int64_t n_samples, int32_t nb)
{
    const int64_t ld = (nb + CV_TILE - 1) / CV_TILE * CV_TILE;
    const int64_t ks = (n_samples + CV_KC - 1) / CV_KC * CV_KC;
    return ld * (ks > 0 ? ks : CV_KC) * 8;
}

int32_t pb2_cov_subsample(int64_t n_samples, int32_t nb, const double *d_xi, const double *d_weights,
                          double *d_cov, double *d_mean_xi, double *d_sum_weights, void *d_scratch,
                          int64_t scratch_bytes, void *stream)
{
    if (!d_xi || !d_weights || !d_cov || !d_mean_xi || !d_sum_weights || !d_scratch) {
        pb2_set_error("pb2_cov_subsample: null pointer argument");
        return PB2_EINVAL;
    }
    if (nb <= 0 || n_samples < 0 || n_samples > 0x7fffffff) {
        pb2_set_error("pb2_cov_subsample: bad sizes (n_samples %lld, nb %d)", (long long)n_samples, nb);
        return PB2_EINVAL;
    }
    if (scratch_bytes < pb2_cov_scratch_bytes(n_samples, nb)) {
        pb2_set_error("pb2_cov_subsample: scratch too small");
        return PB2_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int ld = (nb + CV_TILE - 1) / CV_TILE * CV_TILE;
    int ks = (int)((n_samples + CV_KC - 1) / CV_KC * CV_KC);
    if (ks == 0) ks = CV_KC;
    const int n_tiles = ld / CV_TILE;
    double *M = (double *)d_scratch;
    pb2_timing_begin(s);
    pb2_cov_colstats<<<(nb + 31) / 32, 32, 0, s>>>((int)n_samples, nb, d_xi, d_weights, d_mean_xi,
                                                   d_sum_weights);
    pb2_count_launch(1);
    int32_t rc = pb2_check_launch("pb2_cov_colstats");
    if (rc) return rc;
    pb2_cov_build_m<<<dim3((ld + 255) / 256, ks), 256, 0, s>>>((int)n_samples, nb, ld, d_xi,
                                                                d_weights, d_mean_xi, M);
    pb2_count_launch(1);
    rc = pb2_check_launch("pb2_cov_build_m");
    if (rc) return rc;
    const int smem = CV_STAGES * CV_STAGE_BYTES;
    PB2_CUDA(cudaFuncSetAttribute(pb2_cov_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pb2_cov_syrk<<<n_tiles * (n_tiles + 1) / 2, CV_THREADS, smem, s>>>(nb, ld, ks, n_tiles, M,
                                                                      d_sum_weights, 1., d_cov);
    pb2_count_launch(1);
    rc = pb2_check_launch("pb2_cov_syrk");
    pb2_timing_end(s);
    return rc;
}

static inline int64_t cv_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

int64_t pb2_cov_boot_scratch_bytes(int64_t n_samples, int32_t nb, int32_t n_boot)
{
    const int64_t ld = cv_up(nb, CV_TILE), ks = cv_up(n_samples > 0 ? n_samples : 1, CV_KC);
    const int64_t lc = cv_up(n_boot, CV_TILE), kb = cv_up(n_boot > 0 ? n_boot : 1, CV_KC);
    // counts [ks][lc] + P [ks][ld] + W [ks][ld] + boot_xis [max(kb, lc)][ld]
    return 8 * (ks * lc + 2 * ks * ld + (kb > lc ? kb : lc) * ld);
}

/* utils.compute_cov_boot (py/picca/utils.py:131-150): d_idx [n_boot][n_samples] int32 holds the
 * host-drawn `rng.choice(nhpx, size=nhpx)` of every realisation (the reference's RNG stream);
 * boot_xis[b] = sum_k w[idx_k] xi[idx_k] / sum_k w[idx_k] as two A^T B contractions with the
 * multiplicity matrix, then np.cov(boot_xis, rowvar=False). */
int32_t pb2_cov_boot(int64_t n_samples, int32_t nb, int32_t n_boot, const double *d_xi,
                     const double *d_weights, const int32_t *d_idx, double *d_cov, void *d_scratch,
                     int64_t scratch_bytes, void *stream)
{
    if (!d_xi || !d_weights || !d_idx || !d_cov || !d_scratch) {
        pb2_set_error("pb2_cov_boot: null pointer argument");
        return PB2_EINVAL;
    }
    if (nb <= 0 || n_samples <= 0 || n_samples > 0x7fffffff || n_boot < 2) {
        pb2_set_error("pb2_cov_boot: bad sizes (n_samples %lld, nb %d, n_boot %d)",
                      (long long)n_samples, nb, n_boot);
        return PB2_EINVAL;
    }
    if (scratch_bytes < pb2_cov_boot_scratch_bytes(n_samples, nb, n_boot)) {
        pb2_set_error("pb2_cov_boot: scratch too small");
        return PB2_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int ld = (int)cv_up(nb, CV_TILE), ks = (int)cv_up(n_samples, CV_KC);
    const int lc = (int)cv_up(n_boot, CV_TILE), kb = (int)cv_up(n_boot, CV_KC);
    const int rows_boot = kb > lc ? kb : lc;
    double *cnt = (double *)d_scratch;
    double *P = cnt + (size_t)ks * lc;
    double *Wm = P + (size_t)ks * ld;
    double *boot = Wm + (size_t)ks * ld;
    PB2_CUDA(cudaMemsetAsync(cnt, 0, (size_t)ks * lc * 8, s));
    PB2_CUDA(cudaMemsetAsync(boot, 0, (size_t)rows_boot * ld * 8, s));
    pb2_timing_begin(s);
    const long long total = (long long)n_boot * n_samples;
    pb2_cov_boot_counts<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(total, (int)n_samples, lc,
                                                                        d_idx, cnt);
    pb2_cov_boot_operands<<<dim3((ld + 255) / 256, ks), 256, 0, s>>>((int)n_samples, nb, ld, d_xi,
                                                                    d_weights, P, Wm);
    pb2_count_launch(2);
    int32_t rc = pb2_check_launch("pb2_cov_boot_operands");
    if (rc) return rc;
    const int smem = CV_STAGES * CV_STAGE_BYTES;
    PB2_CUDA(cudaFuncSetAttribute(pb2_cov_gemm_atb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    PB2_CUDA(cudaFuncSetAttribute(pb2_cov_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid(ld / CV_TILE, lc / CV_TILE);
    pb2_cov_gemm_atb<<<grid, CV_THREADS, smem, s>>>(n_boot, nb, lc, ld, ld, ks, cnt, P, 0, boot);
    pb2_cov_gemm_atb<<<grid, CV_THREADS, smem, s>>>(n_boot, nb, lc, ld, ld, ks, cnt, Wm, 1, boot);
    pb2_cov_boot_center<<<(nb + 31) / 32, dim3(32, 16), 0, s>>>(n_boot, nb, ld, boot);
    pb2_count_launch(3);
    rc = pb2_check_launch("pb2_cov_boot_center");
    if (rc) return rc;
    const int n_tiles = ld / CV_TILE;
    pb2_cov_syrk<<<n_tiles * (n_tiles + 1) / 2, CV_THREADS, smem, s>>>(
        nb, ld, kb, n_tiles, boot, nullptr, 1. / (double)(n_boot - 1), d_cov);
    pb2_count_launch(1);
    rc = pb2_check_launch("pb2_cov_syrk(boot)");
    pb2_timing_end(s);
    return rc;
}

int32_t pb2_cov_smooth(int32_t nb, const double *d_cov, const double *d_r_par,
                       const double *d_r_trans, double delta_r_par, double delta_r_trans,
                       int32_t per_r_par, int32_t n_dp, int32_t n_dt, int32_t rp_lo, int32_t n_rp,
                       double *d_table_sum, uint64_t *d_table_count, int32_t *d_bad,
                       double *d_cov_smooth, void *stream)
{
    if (!d_cov || !d_r_par || !d_r_trans || !d_table_sum || !d_table_count || !d_bad ||
        !d_cov_smooth) {
        pb2_set_error("pb2_cov_smooth: null pointer argument");
        return PB2_EINVAL;
    }
    if (nb <= 0 || n_dp <= 0 || n_dt <= 0 || (per_r_par && n_rp <= 0) || !(delta_r_par > 0.) ||
        !(delta_r_trans > 0.)) {
        pb2_set_error("pb2_cov_smooth: bad sizes");
        return PB2_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    CvSmooth S;
    S.nb = nb, S.per_r_par = per_r_par, S.n_dp = n_dp, S.n_dt = n_dt, S.rp_lo = rp_lo;
    S.n_rp = per_r_par ? n_rp : 1;
    S.delta_r_par = delta_r_par, S.delta_r_trans = delta_r_trans;
    const size_t keys = (size_t)S.n_rp * n_dp * n_dt;
    PB2_CUDA(cudaMemsetAsync(d_table_sum, 0, keys * 8, s));
    PB2_CUDA(cudaMemsetAsync(d_table_count, 0, keys * 8, s));
    PB2_CUDA(cudaMemsetAsync(d_bad, 0, 4, s));
    dim3 grid((nb + 255) / 256, nb);
    pb2_timing_begin(s);
    pb2_cov_smooth_accumulate<<<grid, 256, 0, s>>>(S, d_cov, d_r_par, d_r_trans, d_table_sum,
                                                   (unsigned long long *)d_table_count, d_bad);
    pb2_count_launch(1);
    int32_t rc = pb2_check_launch("pb2_cov_smooth_accumulate");
    if (rc) return rc;
    pb2_cov_smooth_apply<<<grid, 256, 0, s>>>(S, d_cov, d_r_par, d_r_trans, d_table_sum,
                                              (const unsigned long long *)d_table_count,
                                              d_cov_smooth);
    pb2_count_launch(1);
    rc = pb2_check_launch("pb2_cov_smooth_apply");
    pb2_timing_end(s);
    return rc;
}

}  // extern "C"
