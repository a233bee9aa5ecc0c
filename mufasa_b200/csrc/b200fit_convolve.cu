// b200fit_convolve.cu -- the convolved-cube stage of iter_2comp_fit on the GPU.
//
// Reference: mufasa/convolve_tools.py  convolve_sky_byfactor :37-87 -> convolve_sky :90-138 ->
// spectral_cube ``convolve_to(beam)`` (every channel plane convolved with the beam that takes the cube's own beam to
// ``factor`` x that beam, astropy ``convolve`` defaults: kernel normalised, NaNs interpolated over, zeros beyond the
// image), then ``reproject(downsampled header, order='bilinear')``.  [RECALLED: spectral_cube / radio_beam /
// astropy.convolution / reproject are absent from the image; their documented behaviour is restated in
// oracle/convolve.py, which is what the parity tests compare against.]
//
//   out(c, y, x) = sum_k K(k) v(c, y + dy_k, x + dx_k) / sum_k K(k) f(c, y + dy_k, x + dx_k)
//     v = value (0 where NaN or masked), f = 1 where the voxel is finite and not masked, and v = 0, f = 1 beyond the
//     image (boundary='fill', fill_value=0: the zeros count in the normalisation);
//   out = NaN where the voxel itself is NaN or masked (the convolved cube keeps the mask of the input cube) or no
//   finite voxel is under the kernel.
//
// One block convolves one 32x32 tile of one channel plane out of shared memory.  A circular beam (the usual case:
// BMAJ == BMIN) gives a rank-1 kernel and takes the separable path -- 3K fused multiply-adds per output and
// accumulator, which leaves the stage bound by the HBM stream of the cube (one read, one write); a rotated
// elliptical beam takes the direct K x K path (compute bound on the fp32 pipe, K^2 per output).
#include <cuda_runtime.h>

#include <cstdint>

#include "b200fit_ctx.h"

namespace b200fit {

constexpr int CT = 32;            // output tile edge
constexpr int CONV_THREADS = 256;
constexpr int CONV_KMAX = 65;     // largest kernel edge (odd)

// stage the (CT + K - 1)^2 input window of the tile: values with NaN / masked -> 0, finite flags as 0 / 1
__device__ __forceinline__ void load_window(const float* __restrict__ plane, const unsigned char* __restrict__ include,
                                            int ny, int nx, int y0, int x0, int K, float* __restrict__ val,
                                            float* __restrict__ fin, int pitch) {
    const int h = K >> 1, W = CT + K - 1;
    for (int e = threadIdx.x; e < W * W; e += CONV_THREADS) {
        const int r = e / W, c = e - r * W;
        const int y = y0 + r - h, x = x0 + c - h;
        float v = 0.f, f = 1.f;   // beyond the image: a finite zero
        if (y >= 0 && y < ny && x >= 0 && x < nx) {
            v = plane[(size_t)y * nx + x];
            const bool ok = (v == v) && (!include || include[(size_t)y * nx + x]);
            f = ok ? 1.f : 0.f;
            v = ok ? v : 0.f;
        }
        val[r * pitch + c] = v;
        fin[r * pitch + c] = f;
    }
}

__device__ __forceinline__ void store_out(float* __restrict__ out, const unsigned char* __restrict__ include, int ny,
                                          int nx, int y, int x, float num, float den, float self_ok) {
    if (y >= ny || x >= nx) return;
    const bool keep = self_ok > 0.f && den > 0.f;   // the cube keeps its own mask: a NaN / excluded voxel stays NaN
    out[(size_t)y * nx + x] = keep ? num / den : __int_as_float(0x7fc00000);
}

// separable kernel: K(dy, dx) = ky[dy] kx[dx]
__global__ void __launch_bounds__(CONV_THREADS)
    convolve_sep_kernel(const float* __restrict__ cube, float* __restrict__ out, const unsigned char* __restrict__ include,
                        int ny, int nx, int K, const float* __restrict__ kx, const float* __restrict__ ky) {
    extern __shared__ float smem[];
    const int W = CT + K - 1, pitch = W + 1;
    float* val = smem;                       // [W][pitch]
    float* fin = val + W * pitch;            // [W][pitch]
    float* mnum = fin + W * pitch;           // [W][CT + 1]   rows of the window, horizontally convolved
    float* mden = mnum + W * (CT + 1);
    float* skx = mden + W * (CT + 1);        // [K]
    float* sky = skx + K;
    const size_t plane = (size_t)ny * nx;
    const float* in = cube + blockIdx.z * plane;
    const int y0 = blockIdx.y * CT, x0 = blockIdx.x * CT;
    for (int k = threadIdx.x; k < K; k += CONV_THREADS) {
        skx[k] = kx[k];
        sky[k] = ky[k];
    }
    load_window(in, include, ny, nx, y0, x0, K, val, fin, pitch);
    __syncthreads();
    // horizontal pass: W rows x CT columns
    for (int e = threadIdx.x; e < W * CT; e += CONV_THREADS) {
        const int r = e / CT, c = e - r * CT;
        const float* v = val + r * pitch + c;
        const float* f = fin + r * pitch + c;
        float num = 0.f, den = 0.f;
        for (int k = 0; k < K; ++k) {
            const float w = skx[k];
            num = fmaf(w, v[k], num);
            den = fmaf(w, f[k], den);
        }
        mnum[r * (CT + 1) + c] = num;
        mden[r * (CT + 1) + c] = den;
    }
    __syncthreads();
    // vertical pass: CT x CT outputs
    for (int e = threadIdx.x; e < CT * CT; e += CONV_THREADS) {
        const int r = e / CT, c = e - r * CT;
        float num = 0.f, den = 0.f;
        for (int k = 0; k < K; ++k) {
            const float w = sky[k];
            num = fmaf(w, mnum[(r + k) * (CT + 1) + c], num);
            den = fmaf(w, mden[(r + k) * (CT + 1) + c], den);
        }
        store_out(out + blockIdx.z * plane, include, ny, nx, y0 + r, x0 + c, num, den,
                  fin[(r + (K >> 1)) * pitch + c + (K >> 1)]);
    }
}

// general kernel k2[K][K] (a rotated elliptical beam); each thread owns 4 outputs of one column, 8 rows apart
__global__ void __launch_bounds__(CONV_THREADS)
    convolve_2d_kernel(const float* __restrict__ cube, float* __restrict__ out, const unsigned char* __restrict__ include,
                       int ny, int nx, int K, const float* __restrict__ k2) {
    extern __shared__ float smem[];
    const int W = CT + K - 1, pitch = W + 1;
    float* val = smem;
    float* fin = val + W * pitch;
    float* sk = fin + W * pitch;             // [K * K]
    const size_t plane = (size_t)ny * nx;
    const float* in = cube + blockIdx.z * plane;
    const int y0 = blockIdx.y * CT, x0 = blockIdx.x * CT;
    for (int k = threadIdx.x; k < K * K; k += CONV_THREADS) sk[k] = k2[k];
    load_window(in, include, ny, nx, y0, x0, K, val, fin, pitch);
    __syncthreads();
    const int c = threadIdx.x & 31, r0 = threadIdx.x >> 5;
    float num[4] = {0.f, 0.f, 0.f, 0.f}, den[4] = {0.f, 0.f, 0.f, 0.f};
    for (int dy = 0; dy < K; ++dy) {
        for (int dx = 0; dx < K; ++dx) {
            const float w = sk[dy * K + dx];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int o = (r0 + 8 * q + dy) * pitch + c + dx;
                num[q] = fmaf(w, val[o], num[q]);
                den[q] = fmaf(w, fin[o], den[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        store_out(out + blockIdx.z * plane, include, ny, nx, y0 + r0 + 8 * q, x0 + c, num[q], den[q],
                  fin[(r0 + 8 * q + (K >> 1)) * pitch + c + (K >> 1)]);
}

// bilinear resampling of every plane onto a grid whose pixel (oy, ox) sits at input pixel
// (ay oy + by, ax ox + bx): reproject's order='bilinear' for two headers related by downsample_header
// (utils/fits_utils.py:11-60).  A position within half a pixel of the image is interpolated with the edge
// replicated; beyond that, and wherever one of the four neighbours is NaN, the result is NaN.
__global__ void __launch_bounds__(256)
    resample_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, int nchan, int ny, int nx, int oy_n,
                             int ox_n, double ay, double by, double ax, double bx) {
    const long long total = (long long)nchan * oy_n * ox_n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(e % ox_n);
        const long long t = e / ox_n;
        const int oy = (int)(t % oy_n);
        const int ch = (int)(t / oy_n);
        double y = fma(ay, (double)oy, by), x = fma(ax, (double)ox, bx);
        float r = __int_as_float(0x7fc00000);
        if (y >= -0.5 && y <= ny - 0.5 && x >= -0.5 && x <= nx - 0.5) {
            y = fmin(fmax(y, 0.0), (double)(ny - 1));
            x = fmin(fmax(x, 0.0), (double)(nx - 1));
            const int yi = min((int)floor(y), ny - 1), xi = min((int)floor(x), nx - 1);
            const int yj = min(yi + 1, ny - 1), xj = min(xi + 1, nx - 1);
            const double fy = y - yi, fx = x - xi;
            const float* p = in + (size_t)ch * ny * nx;
            const double v00 = p[(size_t)yi * nx + xi], v01 = p[(size_t)yi * nx + xj];
            const double v10 = p[(size_t)yj * nx + xi], v11 = p[(size_t)yj * nx + xj];
            r = (float)((1.0 - fy) * ((1.0 - fx) * v00 + fx * v01) + fy * ((1.0 - fx) * v10 + fx * v11));
        }
        out[e] = r;
    }
}

}  // namespace b200fit

using namespace b200fit;

extern "C" {

int b200fit_convolve_planes(b200fit_ctx* ctx, const float* d_cube, int32_t nchan, int32_t ny, int32_t nx,
                            const unsigned char* d_include, int32_t ksize, const float* d_kx, const float* d_ky,
                            const float* d_k2, float* d_out, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (nchan == 0 || ny == 0 || nx == 0) return B200FIT_OK;
    const bool sep = d_kx && d_ky && !d_k2, full = d_k2 && !d_kx && !d_ky;
    if (!d_cube || !d_out || d_cube == d_out || nchan < 0 || ny < 0 || nx < 0 || ksize < 1 || ksize > CONV_KMAX ||
        !(ksize & 1) || !(sep || full) || nchan > 65535)
        return fail(ctx, B200FIT_E_ARG, "convolve_planes: bad argument");
    CK(cudaSetDevice(ctx->device));
    const int W = CT + ksize - 1;
    const dim3 grid((nx + CT - 1) / CT, (ny + CT - 1) / CT, nchan);
    if (sep) {
        const size_t smem = sizeof(float) * (2 * (size_t)W * (W + 1) + 2 * (size_t)W * (CT + 1) + 2 * ksize);
        CK(cudaFuncSetAttribute(convolve_sep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        convolve_sep_kernel<<<grid, CONV_THREADS, smem, (cudaStream_t)stream>>>(d_cube, d_out, d_include, ny, nx, ksize,
                                                                                d_kx, d_ky);
    } else {
        const size_t smem = sizeof(float) * (2 * (size_t)W * (W + 1) + (size_t)ksize * ksize);
        CK(cudaFuncSetAttribute(convolve_2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        convolve_2d_kernel<<<grid, CONV_THREADS, smem, (cudaStream_t)stream>>>(d_cube, d_out, d_include, ny, nx, ksize,
                                                                               d_k2);
    }
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_resample_bilinear(b200fit_ctx* ctx, const float* d_in, int32_t nchan, int32_t ny, int32_t nx, float* d_out,
                              int32_t out_ny, int32_t out_nx, double ay, double by, double ax, double bx, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (nchan == 0 || out_ny == 0 || out_nx == 0) return B200FIT_OK;
    if (!d_in || !d_out || d_in == d_out || nchan < 0 || ny < 1 || nx < 1 || out_ny < 0 || out_nx < 0)
        return fail(ctx, B200FIT_E_ARG, "resample_bilinear: bad argument");
    CK(cudaSetDevice(ctx->device));
    const long long total = (long long)nchan * out_ny * out_nx;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)ctx->num_sms * 32;
    if (blocks > cap) blocks = cap;
    resample_bilinear_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, nchan, ny, nx, out_ny, out_nx,
                                                                                 ay, by, ax, bx);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

}  // extern "C"
