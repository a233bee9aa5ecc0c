// b200fit_prepass.cu -- the moment / signal-mask pre-pass of cubefit_gen on the GPU.
//
// Reference: mufasa/signals.py  get_moments :329-396  (v_estimate :251-272, get_v_mask :296-326, the disk(3)
// dilation :386, spectral_cube moment0/1/2) -- the step that turns the cube into the moment maps
// moment_guess.moment_guesses consumes (multi_v_fit.py:244-297).  The reference does it with skimage / scipy
// binary morphology on [nchan, ny, nx] boolean cubes; here a boolean cube is a BIT CUBE bits[ny*nx][W]
// (bit b of word w <-> channel 32 w + b, W = nchan_stride / 32) and every structuring element becomes a handful of
// word-parallel shifts:
//   ball(2)   33 voxels: centre column dz -2..2, the 8 in-plane neighbours at distance <= sqrt(2) dz -1..1,
//             the 4 at distance 2 dz 0                                  (v_estimate's binary_erosion)
//   cross     7 voxels (scipy generate_binary_structure(3, 1))          (binary_opening = erosion, then dilation)
//   disk(3)   29 pixels in the plane, dz 0                              (signal-mask dilation)
// Border conventions follow the mirror in mufasa_b200/signals.py: erosion sees True outside the cube (skimage),
// dilation sees False.  Everything is integer / comparison work except the moments (fp64 sums).
#include <cuda_runtime.h>

#include "b200fit_ctx.h"

namespace b200fit {

constexpr unsigned FULLW = 0xffffffffu;

// channels of word w that exist (bit set), given nchan
__device__ __forceinline__ unsigned valid_bits(int w, int nchan) {
    const int lo = w << 5;
    if (lo >= nchan) return 0u;
    const int n = nchan - lo;
    return (n >= 32) ? FULLW : ((1u << n) - 1u);
}

// word w of pixel (y, x); outside the plane or the spectral range every bit reads as ``border``
__device__ __forceinline__ unsigned load_word(const unsigned* __restrict__ in, int y, int x, int w, int ny, int nx, int W,
                                              int nchan, unsigned border) {
    if (y < 0 || y >= ny || x < 0 || x >= nx || w < 0 || w >= W) return border;
    const unsigned vb = valid_bits(w, nchan);
    const unsigned v = in[((size_t)y * nx + x) * W + w];
    return (v & vb) | (border & ~vb);
}

// the 32 channels 32 w + b + dz (b = 0..31) of pixel (y, x)
__device__ __forceinline__ unsigned shifted_word(const unsigned* __restrict__ in, int y, int x, int w, int dz, int ny,
                                                 int nx, int W, int nchan, unsigned border) {
    const unsigned c = load_word(in, y, x, w, ny, nx, W, nchan, border);
    if (dz == 0) return c;
    if (dz > 0) {
        const unsigned hi = load_word(in, y, x, w + 1, ny, nx, W, nchan, border);
        return (c >> dz) | (hi << (32 - dz));
    }
    const int d = -dz;
    const unsigned lo = load_word(in, y, x, w - 1, ny, nx, W, nchan, border);
    return (c << d) | (lo >> (32 - d));
}

enum { MORPH_ERODE_BALL2 = 0, MORPH_ERODE_CROSS = 1, MORPH_DILATE_CROSS = 2, MORPH_DILATE_DISK3 = 3 };

__global__ void __launch_bounds__(256)
    bits_morph_kernel(int op, int nchan, int ny, int nx, int W, const unsigned* __restrict__ in,
                      unsigned* __restrict__ out) {
    const long long total = (long long)ny * nx * W;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(t % W);
        const long long p = t / W;
        const int y = (int)(p / nx), x = (int)(p % nx);
        unsigned r;
        if (op == MORPH_ERODE_BALL2) {
            r = FULLW;
            for (int dy = -2; dy <= 2; ++dy)
                for (int dx = -2; dx <= 2; ++dx) {
                    const int r2 = dy * dy + dx * dx;
                    if (r2 > 4) continue;
                    const int dzmax = (r2 == 0) ? 2 : ((r2 <= 2) ? 1 : 0);
                    for (int dz = -dzmax; dz <= dzmax; ++dz)
                        r &= shifted_word(in, y + dy, x + dx, w, dz, ny, nx, W, nchan, FULLW);
                }
        } else if (op == MORPH_ERODE_CROSS) {
            r = shifted_word(in, y, x, w, 0, ny, nx, W, nchan, FULLW);
            r &= shifted_word(in, y, x, w, -1, ny, nx, W, nchan, FULLW) & shifted_word(in, y, x, w, 1, ny, nx, W, nchan, FULLW);
            r &= shifted_word(in, y - 1, x, w, 0, ny, nx, W, nchan, FULLW) & shifted_word(in, y + 1, x, w, 0, ny, nx, W, nchan, FULLW);
            r &= shifted_word(in, y, x - 1, w, 0, ny, nx, W, nchan, FULLW) & shifted_word(in, y, x + 1, w, 0, ny, nx, W, nchan, FULLW);
        } else if (op == MORPH_DILATE_CROSS) {
            r = shifted_word(in, y, x, w, 0, ny, nx, W, nchan, 0u);
            r |= shifted_word(in, y, x, w, -1, ny, nx, W, nchan, 0u) | shifted_word(in, y, x, w, 1, ny, nx, W, nchan, 0u);
            r |= shifted_word(in, y - 1, x, w, 0, ny, nx, W, nchan, 0u) | shifted_word(in, y + 1, x, w, 0, ny, nx, W, nchan, 0u);
            r |= shifted_word(in, y, x - 1, w, 0, ny, nx, W, nchan, 0u) | shifted_word(in, y, x + 1, w, 0, ny, nx, W, nchan, 0u);
        } else {   // MORPH_DILATE_DISK3
            r = 0u;
            for (int dy = -3; dy <= 3; ++dy)
                for (int dx = -3; dx <= 3; ++dx)
                    if (dy * dy + dx * dx <= 9) r |= load_word(in, y + dy, x + dx, w, ny, nx, W, nchan, 0u);
        }
        out[t] = r & valid_bits(w, nchan);
    }
}

// bits = (y > snr_cut * rms[pix]) per channel; NaN data or NaN rms compare false (LazyComparisonMask on raw data).
// One warp per pixel, one ballot per word.
__global__ void __launch_bounds__(256)
    signal_bits_kernel(int nchan, long long npix, int W, const float* __restrict__ spectra, long long stride,
                       const double* __restrict__ rms, double snr_cut, unsigned* __restrict__ bits) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long pix = warp0; pix < npix; pix += nwarps) {
        const float* y = spectra + pix * stride;
        const double thr = rms[pix] * snr_cut;
        for (int w = 0; w < W; ++w) {
            const int i = (w << 5) + lane;
            const bool on = (i < nchan) && ((double)y[i] > thr);
            const unsigned b = __ballot_sync(FULLW, on);
            if (lane == 0) bits[pix * W + w] = b;
        }
    }
}

// out = in & (v_center - hw < V_i < v_center + hw); v_center per pixel (d_vcenter) or one value for the cube.
// NaN centre -> empty window (comparisons false), as in numpy.
__global__ void __launch_bounds__(256)
    bits_window_kernel(int nchan, long long npix, int W, double v0, double dv, const unsigned* __restrict__ in,
                       const double* __restrict__ vcenter, double vcenter_all, double hw, unsigned* __restrict__ out) {
    const long long total = npix * W;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(t % W);
        const long long pix = t / W;
        const double vc = vcenter ? vcenter[pix] : vcenter_all;
        const double hi = vc + hw, lo = vc - hw;
        unsigned m = 0u;
        for (int b = 0; b < 32; ++b) {
            const int i = (w << 5) + b;
            if (i >= nchan) break;
            const double v = v0 + dv * (double)i;
            if (v < hi && v > lo) m |= 1u << b;
        }
        out[t] = in[t] & m;
    }
}

// get_v_at_peak: velocity of the first maximum of the data over the channels set in bits_a (& bits_b) that are
// finite and whose pixel is in the planemask; NaN if there is none.  One warp per pixel.
__global__ void __launch_bounds__(256)
    v_at_peak_kernel(int nchan, long long npix, int W, const float* __restrict__ spectra, long long stride,
                     const unsigned* __restrict__ bits_a, const unsigned* __restrict__ bits_b,
                     const unsigned char* __restrict__ planemask, double v0, double dv, double* __restrict__ vpeak) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long pix = warp0; pix < npix; pix += nwarps) {
        const float* y = spectra + pix * stride;
        const bool plane = planemask ? (planemask[pix] != 0) : true;
        float best = -INFINITY;
        int besti = 0x7fffffff;
        for (int w = 0; w < W; ++w) {
            unsigned m = bits_a[pix * W + w];
            if (bits_b) m &= bits_b[pix * W + w];
            const int i = (w << 5) + lane;
            if (plane && i < nchan && ((m >> lane) & 1u)) {
                const float v = y[i];
                if ((v - v == 0.f) && v > best) {   // finite; strict > keeps the first maximum within a lane
                    best = v;
                    besti = i;
                }
            }
        }
        // warp arg-max, ties -> lowest channel index (np.argmax returns the first)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(FULLW, best, o);
            const int oi = __shfl_xor_sync(FULLW, besti, o);
            if (ob > best || (ob == best && oi < besti)) {
                best = ob;
                besti = oi;
            }
        }
        if (lane == 0) vpeak[pix] = (besti == 0x7fffffff) ? NAN : (v0 + dv * (double)besti);
    }
}

// spectral_cube moment0 / moment1 / moment2 over the channels set in ``bits`` that are finite and in the planemask
// (signals.moments): m0 = sum(I) |dv|, m1 = sum(I v) / sum(I), m2 = sum(I (v - m1)^2) / sum(I); NaN without a channel.
__global__ void __launch_bounds__(256)
    masked_moments_kernel(int nchan, long long npix, int W, const float* __restrict__ spectra, long long stride,
                          const unsigned* __restrict__ bits, const unsigned char* __restrict__ planemask, double v0,
                          double dv, double* __restrict__ m0, double* __restrict__ m1, double* __restrict__ m2) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long pix = warp0; pix < npix; pix += nwarps) {
        const float* y = spectra + pix * stride;
        const bool plane = planemask ? (planemask[pix] != 0) : true;
        double s0 = 0.0, s1 = 0.0;
        int n = 0;
        for (int w = 0; w < W; ++w) {
            const unsigned m = bits[pix * W + w];
            const int i = (w << 5) + lane;
            if (plane && i < nchan && ((m >> lane) & 1u)) {
                const float v = y[i];
                if (v - v == 0.f) {
                    s0 += (double)v;
                    s1 += (double)v * (v0 + dv * (double)i);
                    ++n;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(FULLW, s0, o);
            s1 += __shfl_xor_sync(FULLW, s1, o);
            n += __shfl_xor_sync(FULLW, n, o);
        }
        const double mean = s1 / s0;
        double s2 = 0.0;
        for (int w = 0; w < W; ++w) {
            const unsigned m = bits[pix * W + w];
            const int i = (w << 5) + lane;
            if (plane && i < nchan && ((m >> lane) & 1u)) {
                const float v = y[i];
                if (v - v == 0.f) {
                    const double d = (v0 + dv * (double)i) - mean;
                    s2 += (double)v * d * d;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(FULLW, s2, o);
        if (lane == 0) {
            m0[pix] = (n > 0) ? s0 * fabs(dv) : NAN;
            m1[pix] = (n > 0) ? mean : NAN;
            m2[pix] = (n > 0) ? s2 / s0 : NAN;
        }
    }
}

}  // namespace b200fit

using namespace b200fit;

static unsigned grid_for(const b200fit_ctx* ctx, long long work_items, int per_block) {
    long long blocks = (work_items + per_block - 1) / per_block;
    const long long cap = (long long)ctx->num_sms * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}

extern "C" {

int b200fit_signal_bits(b200fit_ctx* ctx, int32_t nchan, int64_t npix, const float* d_spectra, int64_t nchan_stride,
                        const double* d_rms, double snr_cut, uint32_t* d_bits, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (npix == 0) return B200FIT_OK;
    if (!d_spectra || !d_rms || !d_bits || nchan < 1 || nchan > B200FIT_MAX_NCHAN || nchan_stride < nchan ||
        (nchan_stride & 31) || npix < 0)
        return fail(ctx, B200FIT_E_ARG, "signal_bits: bad argument");
    CK(cudaSetDevice(ctx->device));
    signal_bits_kernel<<<grid_for(ctx, npix, 8), 256, 0, (cudaStream_t)stream>>>(
        nchan, npix, (int)(nchan_stride / 32), d_spectra, nchan_stride, d_rms, snr_cut, d_bits);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_bits_morph(b200fit_ctx* ctx, int32_t op, int32_t nchan, int32_t ny, int32_t nx, int64_t nchan_stride,
                       const uint32_t* d_in, uint32_t* d_out, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (ny == 0 || nx == 0) return B200FIT_OK;
    if (!d_in || !d_out || d_in == d_out || op < 0 || op > 3 || nchan < 1 || nchan > B200FIT_MAX_NCHAN || ny < 0 ||
        nx < 0 || nchan_stride < nchan || (nchan_stride & 31))
        return fail(ctx, B200FIT_E_ARG, "bits_morph: bad argument");
    CK(cudaSetDevice(ctx->device));
    const int W = (int)(nchan_stride / 32);
    bits_morph_kernel<<<grid_for(ctx, (long long)ny * nx * W, 256), 256, 0, (cudaStream_t)stream>>>(op, nchan, ny, nx, W,
                                                                                                    d_in, d_out);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_bits_window(b200fit_ctx* ctx, int32_t nchan, int64_t npix, int64_t nchan_stride, double v0, double dv,
                        const uint32_t* d_in, const double* d_vcenter, double vcenter_all, double hwidth,
                        uint32_t* d_out, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (npix == 0) return B200FIT_OK;
    if (!d_in || !d_out || nchan < 1 || nchan > B200FIT_MAX_NCHAN || npix < 0 || nchan_stride < nchan ||
        (nchan_stride & 31))
        return fail(ctx, B200FIT_E_ARG, "bits_window: bad argument");
    CK(cudaSetDevice(ctx->device));
    const int W = (int)(nchan_stride / 32);
    bits_window_kernel<<<grid_for(ctx, npix * W, 256), 256, 0, (cudaStream_t)stream>>>(nchan, npix, W, v0, dv, d_in,
                                                                                       d_vcenter, vcenter_all, hwidth, d_out);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_v_at_peak(b200fit_ctx* ctx, int32_t nchan, int64_t npix, const float* d_spectra, int64_t nchan_stride,
                      const uint32_t* d_bits_a, const uint32_t* d_bits_b, const unsigned char* d_planemask, double v0,
                      double dv, double* d_vpeak, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (npix == 0) return B200FIT_OK;
    if (!d_spectra || !d_bits_a || !d_vpeak || nchan < 1 || nchan > B200FIT_MAX_NCHAN || npix < 0 ||
        nchan_stride < nchan || (nchan_stride & 31))
        return fail(ctx, B200FIT_E_ARG, "v_at_peak: bad argument");
    CK(cudaSetDevice(ctx->device));
    v_at_peak_kernel<<<grid_for(ctx, npix, 8), 256, 0, (cudaStream_t)stream>>>(
        nchan, npix, (int)(nchan_stride / 32), d_spectra, nchan_stride, d_bits_a, d_bits_b, d_planemask, v0, dv, d_vpeak);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_masked_moments(b200fit_ctx* ctx, int32_t nchan, int64_t npix, const float* d_spectra, int64_t nchan_stride,
                           const uint32_t* d_bits, const unsigned char* d_planemask, double v0, double dv, double* d_m0,
                           double* d_m1, double* d_m2, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (npix == 0) return B200FIT_OK;
    if (!d_spectra || !d_bits || !d_m0 || !d_m1 || !d_m2 || nchan < 1 || nchan > B200FIT_MAX_NCHAN || npix < 0 ||
        nchan_stride < nchan || (nchan_stride & 31))
        return fail(ctx, B200FIT_E_ARG, "masked_moments: bad argument");
    CK(cudaSetDevice(ctx->device));
    masked_moments_kernel<<<grid_for(ctx, npix, 8), 256, 0, (cudaStream_t)stream>>>(
        nchan, npix, (int)(nchan_stride / 32), d_spectra, nchan_stride, d_bits, d_planemask, v0, dv, d_m0, d_m1, d_m2);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

}  // extern "C"
