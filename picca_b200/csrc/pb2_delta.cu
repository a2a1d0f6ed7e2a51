// Delta loader: FITS delta files -> the SoA CSR buffers of the pair kernels, without a per-forest
// Python detour.  Replaces the per-HDU work of io.read_delta_file /
// Delta.from_fitsio (reference py/picca/io.py:338-381, py/picca/data.py:375-474) and the
// per-forest loop of io.read_deltas (py/picca/io.py:493-510: z, distances, weight evolution,
// projection data.py:622-655).
//
//   host  pb2_fits_scan / pb2_fits_cards : walk the HDUs of a (decompressed) FITS buffer and pull
//         the header cards the loader needs -- plain C, no Python per HDU;
//   dev   pb2_delta_unpack  : big-endian BinTable rows -> native fp64 SoA (byte swap +
//         de-interleave), the raw file bytes are uploaded as they are;
//   dev   pb2_delta_prepare : warp per forest: z = 10^loglam / lambda_abs - 1, linear
//         interpolation of r_comov / dist_m on the cosmology table exactly as
//         scipy.interpolate.interp1d evaluates it, weights *= ((1+z)/(1+z_ref))^(alpha-1), the
//         projection of data.py:622-655, per-forest z range.
// The byte traffic is one read of the raw rows and one write of the SoA: HBM-bound, trivially.
#include <stdlib.h>
#include <string.h>

#include "pb2_common.cuh"

// ------------------------------------------------------------------------------------- host: FITS
#define FITS_BLOCK 2880
#define FITS_CARD 80

static inline bool fits_key_is(const uint8_t *card, const char *key8)
{
    return memcmp(card, key8, 8) == 0;
}

// value field of a card ("KEY     = value / comment"); returns false if the card has no value
static inline bool fits_has_value(const uint8_t *card) { return card[8] == '=' && card[9] == ' '; }

static long long fits_int(const uint8_t *card)
{
    char tmp[72];
    memcpy(tmp, card + 10, 70);
    tmp[70] = 0;
    return strtoll(tmp, nullptr, 10);
}

extern "C" {

/* Walk the HDUs of a FITS buffer.  info[h] = {header_off, data_off, data_bytes, bitpix, naxis,
 * naxis1, naxis2, tfields}.  Returns the number of HDUs (<= max_hdu), or -1 on a malformed file. */
int64_t pb2_fits_scan(const uint8_t *buf, int64_t len, int64_t max_hdu, int64_t *info)
{
    int64_t pos = 0, n = 0;
    while (pos + FITS_BLOCK <= len && n < max_hdu) {
        int64_t bitpix = 0, naxis = 0, pcount = 0, gcount = 1, tfields = 0;
        int64_t axes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int64_t header_off = pos;
        bool end = false, first = true;
        while (!end) {
            if (pos + FITS_BLOCK > len) {
                pb2_set_error("pb2_fits_scan: header of HDU %lld runs past the end of the file",
                              (long long)n);
                return -1;
            }
            for (int c = 0; c < FITS_BLOCK / FITS_CARD && !end; ++c) {
                const uint8_t *card = buf + pos + c * FITS_CARD;
                if (first) {
                    if (!fits_key_is(card, "SIMPLE  ") && !fits_key_is(card, "XTENSION")) {
                        // trailing padding after the last HDU
                        if (n > 0) return n;
                        pb2_set_error("pb2_fits_scan: not a FITS file");
                        return -1;
                    }
                    first = false;
                }
                if (fits_key_is(card, "END     ")) {
                    end = true;
                } else if (fits_has_value(card)) {
                    if (fits_key_is(card, "BITPIX  ")) bitpix = fits_int(card);
                    else if (fits_key_is(card, "NAXIS   ")) naxis = fits_int(card);
                    else if (fits_key_is(card, "PCOUNT  ")) pcount = fits_int(card);
                    else if (fits_key_is(card, "GCOUNT  ")) gcount = fits_int(card);
                    else if (fits_key_is(card, "TFIELDS ")) tfields = fits_int(card);
                    else if (memcmp(card, "NAXIS", 5) == 0 && card[5] >= '1' && card[5] <= '8' &&
                             card[6] == ' ')
                        axes[card[5] - '1'] = fits_int(card);
                }
            }
            pos += FITS_BLOCK;
        }
        int64_t cells = naxis > 0 ? 1 : 0;
        for (int a = 0; a < naxis && a < 8; ++a) cells *= axes[a];
        int64_t bytes = (bitpix < 0 ? -bitpix : bitpix) / 8 * gcount * (pcount + cells);
        if (bytes < 0 || pos + bytes > len) {
            pb2_set_error("pb2_fits_scan: data of HDU %lld runs past the end of the file",
                          (long long)n);
            return -1;
        }
        int64_t *row = info + 8 * n;
        row[0] = header_off, row[1] = pos, row[2] = bytes, row[3] = bitpix, row[4] = naxis;
        row[5] = axes[0], row[6] = axes[1], row[7] = tfields;
        pos += (bytes + FITS_BLOCK - 1) / FITS_BLOCK * FITS_BLOCK;
        ++n;
    }
    return n;
}

/* Header cards of many HDUs at once.  keys: n_keys x 8 characters (blank padded).  For HDU h and
 * key k (index h*n_keys + k): kind = 0 missing, 1 number, 2 string, 3 logical; num = value as
 * double (logical: 1/0); inum = value as int64 when the literal is an integer (else 0);
 * str (24 characters, NUL padded) = string value without quotes and trailing blanks. */
int32_t pb2_fits_cards(const uint8_t *buf, int64_t len, int64_t n_hdu, const int64_t *header_off,
                       int32_t n_keys, const char *keys, int32_t *kind, double *num, int64_t *inum,
                       char *str)
{
    for (int64_t h = 0; h < n_hdu; ++h) {
        int32_t *kd = kind + h * n_keys;
        double *nm = num + h * n_keys;
        int64_t *in = inum + h * n_keys;
        char *st = str + h * n_keys * 24;
        for (int k = 0; k < n_keys; ++k) kd[k] = 0, nm[k] = 0., in[k] = 0;
        memset(st, 0, (size_t)n_keys * 24);
        int64_t pos = header_off[h];
        bool end = false;
        while (!end) {
            if (pos + FITS_BLOCK > len) {
                pb2_set_error("pb2_fits_cards: header %lld runs past the end of the file",
                              (long long)h);
                return PB2_EINVAL;
            }
            for (int c = 0; c < FITS_BLOCK / FITS_CARD && !end; ++c) {
                const uint8_t *card = buf + pos + c * FITS_CARD;
                if (fits_key_is(card, "END     ")) {
                    end = true;
                    break;
                }
                if (!fits_has_value(card)) continue;
                for (int k = 0; k < n_keys; ++k) {
                    if (kd[k] || memcmp(card, keys + 8 * k, 8) != 0) continue;
                    int p = 10;
                    while (p < FITS_CARD && card[p] == ' ') ++p;
                    if (p >= FITS_CARD) break;
                    if (card[p] == '\'') {  // string: '' is an escaped quote
                        int o = 0;
                        ++p;
                        while (p < FITS_CARD) {
                            if (card[p] == '\'') {
                                if (p + 1 < FITS_CARD && card[p + 1] == '\'') {
                                    if (o < 23) st[24 * k + o++] = '\'';
                                    p += 2;
                                    continue;
                                }
                                break;
                            }
                            if (o < 23) st[24 * k + o++] = (char)card[p];
                            ++p;
                        }
                        while (o > 0 && st[24 * k + o - 1] == ' ') st[24 * k + --o] = 0;
                        kd[k] = 2;
                    } else if ((card[p] == 'T' || card[p] == 'F') &&
                               (p + 1 >= FITS_CARD || card[p + 1] == ' ' || card[p + 1] == '/')) {
                        kd[k] = 3;
                        nm[k] = card[p] == 'T' ? 1. : 0.;
                        in[k] = card[p] == 'T';
                    } else {
                        char tmp[72];
                        int o = 0;
                        bool integer = true;
                        while (p < FITS_CARD && card[p] != ' ' && card[p] != '/' && o < 70) {
                            char ch = (char)card[p++];
                            if (ch == 'D' || ch == 'd') ch = 'E';  // Fortran double exponent
                            if (!((ch >= '0' && ch <= '9') || ch == '-' || ch == '+')) integer = false;
                            tmp[o++] = ch;
                        }
                        tmp[o] = 0;
                        if (o) {
                            kd[k] = 1;
                            nm[k] = strtod(tmp, nullptr);
                            in[k] = integer ? strtoll(tmp, nullptr, 10) : 0;
                        }
                    }
                    break;
                }
            }
            pos += FITS_BLOCK;
        }
    }
    return 0;
}

/* One long keyword in the ESO HIERARCH convention ("HIERARCH WAVE_SOLUTION = 'lin'", what fitsio
 * writes for keywords longer than 8 characters) of the header starting at header_off.
 * *kind = 0 missing / 1 number / 2 string; *num = numeric value; str24 = string value. */
int32_t pb2_fits_hierarch(const uint8_t *buf, int64_t len, int64_t header_off, const char *key,
                          int32_t *kind, double *num, char *str24)
{
    *kind = 0;
    *num = 0.;
    memset(str24, 0, 24);
    const size_t klen = strlen(key);
    int64_t pos = header_off;
    for (;;) {
        if (pos + FITS_BLOCK > len) {
            pb2_set_error("pb2_fits_hierarch: header runs past the end of the file");
            return PB2_EINVAL;
        }
        for (int c = 0; c < FITS_BLOCK / FITS_CARD; ++c) {
            const uint8_t *card = buf + pos + c * FITS_CARD;
            if (fits_key_is(card, "END     ")) return 0;
            if (memcmp(card, "HIERARCH", 8) != 0) continue;
            int p = 8;
            while (p < FITS_CARD && card[p] == ' ') ++p;
            if (p + (int)klen >= FITS_CARD || memcmp(card + p, key, klen) != 0) continue;
            p += (int)klen;
            while (p < FITS_CARD && card[p] == ' ') ++p;
            if (p >= FITS_CARD || card[p] != '=') continue;
            ++p;
            while (p < FITS_CARD && card[p] == ' ') ++p;
            if (p >= FITS_CARD) return 0;
            if (card[p] == '\'') {
                int o = 0;
                ++p;
                while (p < FITS_CARD && card[p] != '\'') {
                    if (o < 23) str24[o++] = (char)card[p];
                    ++p;
                }
                while (o > 0 && str24[o - 1] == ' ') str24[--o] = 0;
                *kind = 2;
            } else {
                char tmp[72];
                int o = 0;
                while (p < FITS_CARD && card[p] != ' ' && card[p] != '/' && o < 70) {
                    char ch = (char)card[p++];
                    if (ch == 'D' || ch == 'd') ch = 'E';
                    tmp[o++] = ch;
                }
                tmp[o] = 0;
                if (o) {
                    *kind = 1;
                    *num = strtod(tmp, nullptr);
                }
            }
            return 0;
        }
        pos += FITS_BLOCK;
    }
}

}  // extern "C"

// ---------------------------------------------------------------------------------- device: unpack
__device__ __forceinline__ double be64_to_double(const uint8_t *p)
{
    unsigned long long v;
    if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) {
        const uint2 w = *reinterpret_cast<const uint2 *>(p);
        v = ((unsigned long long)__byte_perm(w.x, 0, 0x0123) << 32) | __byte_perm(w.y, 0, 0x0123);
    } else {
        v = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) v = (v << 8) | p[b];
    }
    return __longlong_as_double((long long)v);
}

// one warp per forest; forest f: rows at raw + row0[f], row_bytes[f] apart; column byte offsets
// col_off[3 f + {0,1,2}] = wavelength (LOGLAM or LAMBDA), DELTA, WEIGHT
__global__ void pb2_delta_unpack_kernel(long long n_los, const uint8_t *__restrict__ raw,
                                        const long long *__restrict__ row0,
                                        const int *__restrict__ row_bytes,
                                        const int *__restrict__ col_off,
                                        const long long *__restrict__ offset,
                                        double *__restrict__ log_lambda,
                                        double *__restrict__ delta, double *__restrict__ weights)
{
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= n_los) return;
    const int lane = threadIdx.x & 31;
    const long long a = offset[f];
    const int n = (int)(offset[f + 1] - a);
    const uint8_t *base = raw + row0[f];
    const int rb = row_bytes[f];
    const int c0 = col_off[3 * f], c1 = col_off[3 * f + 1], c2 = col_off[3 * f + 2];
    for (int p = lane; p < n; p += 32) {
        const uint8_t *row = base + (long long)p * rb;
        log_lambda[a + p] = be64_to_double(row + c0);
        delta[a + p] = be64_to_double(row + c1);
        weights[a + p] = be64_to_double(row + c2);
    }
}

// ImageHDU flavour (Delta.from_image, data.py:519-620): one common wavelength grid of n_lambda
// pixels and 2-D images [n_forest][n_lambda]; forest f is image row rows[f] and keeps the pixels
// with WEIGHT > 0 (data.py:572, :604-611).  Pass 1 counts them, pass 2 compacts them (warp per
// forest, ballot prefix) into the CSR arrays.
__global__ void pb2_delta_image_count_kernel(long long n_los, const uint8_t *__restrict__ raw,
                                             long long weight_off, int n_lambda,
                                             const int *__restrict__ rows, int *__restrict__ count)
{
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= n_los) return;
    const int lane = threadIdx.x & 31;
    const uint8_t *w = raw + weight_off + (long long)rows[f] * n_lambda * 8;
    int n = 0;
    for (int p0 = 0; p0 < n_lambda; p0 += 32) {
        const int p = p0 + lane;
        const bool keep = p < n_lambda && be64_to_double(w + 8ll * p) > 0.;
        n += __popc(__ballot_sync(0xffffffffu, keep));
    }
    if (lane == 0) count[f] = n;
}

__global__ void pb2_delta_image_fill_kernel(long long n_los, const uint8_t *__restrict__ raw,
                                            long long lambda_off, long long delta_off,
                                            long long weight_off, int n_lambda,
                                            const int *__restrict__ rows,
                                            const long long *__restrict__ offset,
                                            double *__restrict__ log_lambda,
                                            double *__restrict__ delta, double *__restrict__ weights)
{
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= n_los) return;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)rows[f] * n_lambda * 8;
    const uint8_t *w = raw + weight_off + row, *d = raw + delta_off + row, *l = raw + lambda_off;
    long long out = offset[f];
    for (int p0 = 0; p0 < n_lambda; p0 += 32) {
        const int p = p0 + lane;
        double wv = 0.;
        if (p < n_lambda) wv = be64_to_double(w + 8ll * p);
        const bool keep = p < n_lambda && wv > 0.;
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const long long at = out + __popc(mask & ((1u << lane) - 1u));
            weights[at] = wv;
            delta[at] = be64_to_double(d + 8ll * p);
            log_lambda[at] = be64_to_double(l + 8ll * p);
        }
        out += __popc(mask);
    }
}

// ---------------------------------------------------------------------------------- device: rebin
// Delta.rebin (data.py:657-686): the pixels of a forest are summed into bins of `factor` original
// pixels: wave = 10**log_lambda; start = wave.min() - dwave/2; num_bins = ceil(((wave[-1] -
// wave[0])/dwave + 1)/factor); edges[k] = k*dwave*factor + start; a pixel goes to bin
// searchsorted(edges, wave) (left); bins 1 .. num_bins-1 with a non-zero weight sum survive, at the
// mid-points of their edges.  One warp per forest, lane = output bin; the pixels of a bin are a
// contiguous range of the (ascending) wavelengths, found by binary search and added in order --
// the association of np.bincount, so the sums are bit-equal.  Pass 1 (fill == false) counts the
// surviving bins, pass 2 compacts them (ballot prefix) into the new CSR arrays.
__device__ __forceinline__ double rb_edge(int k, double dwave, double factor, double start)
{
    // np.arange(num_bins) * dwave * factor + start: left to right
    return add_rn(mul_rn(mul_rn((double)k, dwave), factor), start);
}

// first pixel with wave > v
__device__ __forceinline__ int rb_upper(const double *__restrict__ w, int n, double v)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (w[mid] <= v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

template <bool FILL>
__global__ void pb2_delta_rebin_kernel(long long n_los, const long long *__restrict__ offset,
                                       const double *__restrict__ wave,
                                       const double *__restrict__ delta,
                                       const double *__restrict__ weights,
                                       const double *__restrict__ dwave_of, int factor_i,
                                       int *__restrict__ count, int *__restrict__ status,
                                       const long long *__restrict__ new_offset,
                                       double *__restrict__ new_wave, double *__restrict__ new_delta,
                                       double *__restrict__ new_weights)
{
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= n_los) return;
    const int lane = threadIdx.x & 31;
    const long long a = offset[f];
    const int n = (int)(offset[f + 1] - a);
    if (n == 0) {
        if (!FILL && lane == 0) count[f] = 0;
        return;
    }
    const double *__restrict__ w = wave + a;
    const double dwave = dwave_of[f], factor = (double)factor_i;
    // wave.min() and the ascending order the bin ranges rely on
    double wmin = 1e300;
    bool sorted = true;
    for (int p = lane; p < n; p += 32) {
        wmin = fmin(wmin, w[p]);
        if (p > 0 && w[p] < w[p - 1]) sorted = false;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) wmin = fmin(wmin, __shfl_xor_sync(0xffffffffu, wmin, m));
    if (!__all_sync(0xffffffffu, sorted)) {
        if (lane == 0) atomicExch(status, 2);
        if (!FILL && lane == 0) count[f] = 0;
        return;
    }
    const double start = sub_rn(wmin, div_rn(dwave, 2.));
    const int num_bins = (int)ceil(div_rn(add_rn(div_rn(sub_rn(w[n - 1], w[0]), dwave), 1.), factor));
    int kept = 0;
    long long out = FILL ? new_offset[f] : 0;
    for (int k0 = 1; k0 < num_bins; k0 += 32) {  // [1:-1] of the bincount: bins 1 .. num_bins-1
        const int k = k0 + lane;
        double sw = 0., sdw = 0., lo_e = 0., hi_e = 0.;
        if (k < num_bins) {
            lo_e = rb_edge(k - 1, dwave, factor, start);
            hi_e = rb_edge(k, dwave, factor, start);
            const int p_lo = rb_upper(w, n, lo_e), p_hi = rb_upper(w, n, hi_e);
            for (int p = p_lo; p < p_hi; ++p) {  // edges[k-1] < wave <= edges[k]
                const double wp = weights[a + p];
                sdw = add_rn(sdw, mul_rn(delta[a + p], wp));
                sw = add_rn(sw, wp);
            }
        }
        const bool keep = k < num_bins && sw != 0.;  // mask = binned_weight != 0
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (FILL && keep) {
            const long long at = out + __popc(mask & ((1u << lane) - 1u));
            new_wave[at] = div_rn(add_rn(hi_e, lo_e), 2.);  // (edges[1:] + edges[:-1]) / 2
            new_delta[at] = div_rn(sdw, sw);
            new_weights[at] = sw;
        }
        out += __popc(mask);
        kept += __popc(mask);
    }
    if (!FILL && lane == 0) count[f] = kept;
}

// wave = 10**log_lambda (log_lambda = log10(lambda) first when the file stores LAMBDA)
__global__ void pb2_delta_wave_kernel(long long n_pix, int wave_is_lambda,
                                      double *__restrict__ log_lambda, double *__restrict__ wave)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pix) return;
    double ll = log_lambda[p];
    if (wave_is_lambda) {
        ll = log10(ll);
        log_lambda[p] = ll;
    }
    wave[p] = exp10(ll);
}

// --------------------------------------------------------------------------------- device: prepare
struct DeltaPrep {
    double lambda_abs, alpha_m1, one_plus_z_ref;
    int n_table, project, has_z_in, wave_is_lambda;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// scipy.interpolate.interp1d(kind='linear')._call_linear: idx = searchsorted(x, x_new) clipped to
// [1, n-1]; slope = (y_hi - y_lo) / (x_hi - x_lo); y = slope * (x_new - x_lo) + y_lo, each
// operation rounded separately
__device__ __forceinline__ int table_index(const double *__restrict__ x, int n, double v)
{
    int lo = 0, hi = n;  // first index with x[idx] >= v (side='left')
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (x[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return min(max(lo, 1), n - 1);
}
__device__ __forceinline__ double table_eval(const double *__restrict__ x,
                                             const double *__restrict__ y, int idx, double v)
{
    const double x_lo = x[idx - 1], x_hi = x[idx], y_lo = y[idx - 1], y_hi = y[idx];
    const double slope = div_rn(sub_rn(y_hi, y_lo), sub_rn(x_hi, x_lo));
    return add_rn(mul_rn(slope, sub_rn(v, x_lo)), y_lo);
}

__global__ void pb2_delta_prepare_kernel(long long n_los, DeltaPrep S,
                                         const long long *__restrict__ offset,
                                         const int *__restrict__ order,
                                         const double *__restrict__ tab_z,
                                         const double *__restrict__ tab_r_comov,
                                         const double *__restrict__ tab_dist_m,
                                         const double *__restrict__ z_in,
                                         double *__restrict__ log_lambda, double *__restrict__ delta,
                                         double *__restrict__ weights, double *__restrict__ z_out,
                                         double *__restrict__ r_comov, double *__restrict__ dist_m,
                                         double *__restrict__ z_range, int *__restrict__ status)
{
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= n_los) return;
    const int lane = threadIdx.x & 31;
    const long long a = offset[f];
    const int n = (int)(offset[f + 1] - a);
    double zmin = 1e300, zmax = -1e300, sw = 0., swd = 0., swl = 0.;
    for (int p = lane; p < n; p += 32) {
        double ll = log_lambda[a + p];
        if (S.wave_is_lambda) {  // data.py:411-412
            ll = log10(ll);
            log_lambda[a + p] = ll;
        }
        // z = 10**log_lambda / lambda_abs - 1 (io.py:496); parity mode supplies the host's power
        const double z = S.has_z_in ? z_in[a + p] : sub_rn(div_rn(exp10(ll), S.lambda_abs), 1.0);
        z_out[a + p] = z;
        zmin = fmin(zmin, z);
        zmax = fmax(zmax, z);
        if (tab_z) {
            if (!(z >= tab_z[0] && z <= tab_z[S.n_table - 1])) {
                atomicExch(status, 1);  // interp1d raises ValueError outside the table
                r_comov[a + p] = dist_m[a + p] = nan("");
            } else {
                const int idx = table_index(tab_z, S.n_table, z);
                r_comov[a + p] = table_eval(tab_z, tab_r_comov, idx, z);
                dist_m[a + p] = table_eval(tab_z, tab_dist_m, idx, z);
            }
        }
        // delta.weights *= ((1 + z) / (1 + z_ref))**(alpha - 1)  (io.py:503)
        const double w = mul_rn(weights[a + p],
                                pow(div_rn(add_rn(1.0, z), S.one_plus_z_ref), S.alpha_m1));
        weights[a + p] = w;
        sw += w;
        swd += w * delta[a + p];
        swl += w * ll;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        zmin = fmin(zmin, __shfl_xor_sync(0xffffffffu, zmin, m));
        zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, m));
    }
    if (lane == 0) {
        z_range[2 * f] = zmin;
        z_range[2 * f + 1] = zmax;
    }
    if (!S.project) return;
    // ---- Delta.project (data.py:622-655)
    sw = warp_sum(sw);
    if (!(sw > 0.0)) return;  // :636-640
    swd = warp_sum(swd);
    swl = warp_sum(swl);
    const double mean_delta = swd / sw;  // np.average(delta, weights=weights)
    const int ord = order[f];
    if (ord == 1 && n > 1) {
        const double mean_ll = swl / sw;
        double num = 0., den = 0.;
        for (int p = lane; p < n; p += 32) {
            const double dl = log_lambda[a + p] - mean_ll;
            const double w = weights[a + p];
            num += w * delta[a + p] * dl;
            den += w * dl * dl;
        }
        num = warp_sum(num);
        den = warp_sum(den);
        const double coef = num / den;
        for (int p = lane; p < n; p += 32)
            delta[a + p] -= mean_delta + coef * (log_lambda[a + p] - mean_ll);
    } else if (ord == 1) {
        for (int p = lane; p < n; p += 32) delta[a + p] -= mean_delta + delta[a + p];  // :651-652
    } else {
        for (int p = lane; p < n; p += 32) delta[a + p] -= mean_delta;
    }
}

extern "C" {

int32_t pb2_delta_unpack(int64_t n_los, const uint8_t *d_raw, const int64_t *d_row0,
                         const int32_t *d_row_bytes, const int32_t *d_col_off,
                         const int64_t *d_offset, double *d_log_lambda, double *d_delta,
                         double *d_weights, void *stream)
{
    if (n_los <= 0) return 0;
    if (!d_raw || !d_row0 || !d_row_bytes || !d_col_off || !d_offset || !d_log_lambda || !d_delta ||
        !d_weights) {
        pb2_set_error("pb2_delta_unpack: null pointer argument");
        return PB2_EINVAL;
    }
    const int warps = 8;
    pb2_delta_unpack_kernel<<<(unsigned)((n_los + warps - 1) / warps), warps * 32, 0,
                              (cudaStream_t)stream>>>(
        n_los, d_raw, (const long long *)d_row0, d_row_bytes, d_col_off, (const long long *)d_offset,
        d_log_lambda, d_delta, d_weights);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_delta_unpack");
}

int32_t pb2_delta_image_count(int64_t n_los, const uint8_t *d_raw, int64_t weight_off,
                              int32_t n_lambda, const int32_t *d_rows, int32_t *d_count,
                              void *stream)
{
    if (n_los <= 0) return 0;
    if (!d_raw || !d_rows || !d_count || n_lambda <= 0) {
        pb2_set_error("pb2_delta_image_count: bad argument");
        return PB2_EINVAL;
    }
    const int warps = 8;
    pb2_delta_image_count_kernel<<<(unsigned)((n_los + warps - 1) / warps), warps * 32, 0,
                                   (cudaStream_t)stream>>>(n_los, d_raw, weight_off, n_lambda, d_rows,
                                                           d_count);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_delta_image_count");
}

int32_t pb2_delta_image_unpack(int64_t n_los, const uint8_t *d_raw, int64_t lambda_off,
                               int64_t delta_off, int64_t weight_off, int32_t n_lambda,
                               const int32_t *d_rows, const int64_t *d_offset, double *d_log_lambda,
                               double *d_delta, double *d_weights, void *stream)
{
    if (n_los <= 0) return 0;
    if (!d_raw || !d_rows || !d_offset || !d_log_lambda || !d_delta || !d_weights || n_lambda <= 0) {
        pb2_set_error("pb2_delta_image_unpack: bad argument");
        return PB2_EINVAL;
    }
    const int warps = 8;
    pb2_delta_image_fill_kernel<<<(unsigned)((n_los + warps - 1) / warps), warps * 32, 0,
                                  (cudaStream_t)stream>>>(
        n_los, d_raw, lambda_off, delta_off, weight_off, n_lambda, d_rows,
        (const long long *)d_offset, d_log_lambda, d_delta, d_weights);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_delta_image_unpack");
}

int32_t pb2_delta_wave(int64_t n_pix, int32_t wave_is_lambda, double *d_log_lambda, double *d_wave,
                       void *stream)
{
    if (n_pix <= 0) return 0;
    if (!d_log_lambda || !d_wave) {
        pb2_set_error("pb2_delta_wave: null pointer argument");
        return PB2_EINVAL;
    }
    pb2_delta_wave_kernel<<<(unsigned)((n_pix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n_pix, wave_is_lambda, d_log_lambda, d_wave);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_delta_wave");
}

int32_t pb2_delta_rebin(int64_t n_los, const int64_t *d_offset, const double *d_wave,
                        const double *d_delta, const double *d_weights, const double *d_dwave,
                        int32_t factor, int32_t *d_count, int32_t *d_status,
                        const int64_t *d_new_offset, double *d_new_wave, double *d_new_delta,
                        double *d_new_weights, void *stream)
{
    if (n_los <= 0) return 0;
    if (!d_offset || !d_wave || !d_delta || !d_weights || !d_dwave || !d_status || factor < 1 ||
        (!d_new_offset && !d_count) ||
        (d_new_offset && (!d_new_wave || !d_new_delta || !d_new_weights))) {
        pb2_set_error("pb2_delta_rebin: bad argument");
        return PB2_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((n_los + 7) / 8);
    if (!d_new_offset)
        pb2_delta_rebin_kernel<false><<<blocks, 256, 0, s>>>(
            n_los, (const long long *)d_offset, d_wave, d_delta, d_weights, d_dwave, factor, d_count,
            d_status, nullptr, nullptr, nullptr, nullptr);
    else
        pb2_delta_rebin_kernel<true><<<blocks, 256, 0, s>>>(
            n_los, (const long long *)d_offset, d_wave, d_delta, d_weights, d_dwave, factor, nullptr,
            d_status, (const long long *)d_new_offset, d_new_wave, d_new_delta, d_new_weights);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_delta_rebin");
}

int32_t pb2_delta_prepare(int64_t n_los, const int64_t *d_offset, const int32_t *d_order,
                          double lambda_abs, double alpha, double z_ref, int32_t n_table,
                          const double *d_tab_z, const double *d_tab_r_comov,
                          const double *d_tab_dist_m, int32_t project, int32_t wave_is_lambda,
                          const double *d_z_in, double *d_log_lambda, double *d_delta,
                          double *d_weights, double *d_z, double *d_r_comov, double *d_dist_m,
                          double *d_z_range, int32_t *d_status, void *stream)
{
    if (n_los <= 0) return 0;
    if (!d_offset || !d_order || !d_log_lambda || !d_delta || !d_weights || !d_z || !d_z_range ||
        !d_status || (d_tab_z && (!d_tab_r_comov || !d_tab_dist_m || !d_r_comov || !d_dist_m ||
                                  n_table < 2))) {
        pb2_set_error("pb2_delta_prepare: null pointer argument");
        return PB2_EINVAL;
    }
    DeltaPrep S;
    S.lambda_abs = lambda_abs;
    S.alpha_m1 = alpha - 1.;
    S.one_plus_z_ref = 1. + z_ref;
    S.n_table = n_table;
    S.project = project;
    S.has_z_in = d_z_in != nullptr;
    S.wave_is_lambda = wave_is_lambda;
    cudaStream_t s = (cudaStream_t)stream;
    PB2_CUDA(cudaMemsetAsync(d_status, 0, 4, s));
    const int warps = 8;
    pb2_delta_prepare_kernel<<<(unsigned)((n_los + warps - 1) / warps), warps * 32, 0, s>>>(
        n_los, S, (const long long *)d_offset, d_order, d_tab_z, d_tab_r_comov, d_tab_dist_m, d_z_in,
        d_log_lambda, d_delta, d_weights, d_z, d_r_comov, d_dist_m, d_z_range, d_status);
    pb2_count_launch(1);
    return pb2_check_launch("pb2_delta_prepare");
}

}  // extern "C"
