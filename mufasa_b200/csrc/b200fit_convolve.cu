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
// One block convolves one 32 x 64 tile of one channel plane out of shared memory.  A circular beam (the usual case:
// BMAJ == BMIN) gives a rank-1 kernel and takes the separable path -- 3K fused multiply-adds per output and
// accumulator, which leaves the stage bound by the HBM stream of the cube (one read, one write); a rotated
// elliptical beam takes the direct K x K path (compute bound on the fp32 pipe, K^2 per output).
#include <cuda_runtime.h>

#include <cstdint>

#include "b200fit_ctx.h"

namespace b200fit {

constexpr int TX = 32, TY = 64;   // output tile: 32 columns x 64 rows
constexpr int NOUT = 8;           // outputs per thread along the convolution axis
constexpr int CONV_THREADS = 256;
constexpr int CONV_KMAX = 65;     // largest kernel edge (odd)
constexpr int MP = TX + 1;        // pitch of the horizontally convolved rows

// stage the (TY + K - 1) x (TX + K - 1) input window of the tile: values with NaN / masked -> 0, finite flags as
// 0 / 1.  A warp takes a window row at a time (coalesced, no index division).  Returns true when every voxel this
// thread staged is finite (zeros beyond the image count as finite: boundary='fill').
__device__ __forceinline__ bool load_window(const float* __restrict__ plane, const unsigned char* __restrict__ include,
                                            int ny, int nx, int y0, int x0, int K, float* __restrict__ val,
                                            float* __restrict__ fin, int pitch) {
    const int h = K >> 1, WX = TX + K - 1, WY = TY + K - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool all_ok = true;
    // every load is unconditional (clamped address) and 4 rows are in flight per thread: the staging is a latency
    // problem, ~600 cycles per dependent global access against ~20 cycles of work per value
    constexpr int NW = CONV_THREADS / 32, UR = 4;
    for (int rb = warp; rb < WY; rb += NW * UR) {
#pragma unroll 1
        for (int c = lane; c < WX; c += 32) {
            const int x = x0 + c - h;
            const bool col_in = (x >= 0 && x < nx);
            const int xc = min(max(x, 0), nx - 1);
            float v[UR];
            unsigned char m[UR];
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const int yc = min(max(y0 + rb + u * NW - h, 0), ny - 1);
                v[u] = __ldg(plane + (size_t)yc * nx + xc);
                m[u] = include ? __ldg(include + (size_t)yc * nx + xc) : (unsigned char)1;
            }
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const int r = rb + u * NW, y = y0 + r - h;
                if (r < WY) {
                    const bool inside = col_in && y >= 0 && y < ny;
                    const bool ok = !inside || ((v[u] == v[u]) && m[u]);   // beyond the image: a finite zero
                    val[r * pitch + c] = (inside && ok) ? v[u] : 0.f;
                    fin[r * pitch + c] = ok ? 1.f : 0.f;
                    all_ok = all_ok && ok;
                }
            }
        }
    }
    return all_ok;
}

__device__ __forceinline__ void store_out(float* __restrict__ out, int ny, int nx, int y, int x, float num, float den,
                                          float self_ok) {
    if (y >= ny || x >= nx) return;
    const bool keep = self_ok > 0.f && den > 0.f;   // the cube keeps its own mask: a NaN / excluded voxel stays NaN
    out[(size_t)y * nx + x] = keep ? num / den : __int_as_float(0x7fc00000);
}

// Sliding register window: NOUT consecutive outputs along the convolution axis share every value read from shared
// memory (1 load per NOUT FMAs; the one-output-per-load loop was bound by shared-memory bandwidth, then by the
// loop's own index arithmetic -- profiles/r1n).  ``w`` holds the K weights followed by NOUT - 1 zeros.
template <bool WITH_DEN>
__device__ __forceinline__ void slide(const float* __restrict__ v, const float* __restrict__ f, int stride, int K,
                                      const float* __restrict__ w, float (&num)[NOUT], float (&den)[NOUT]) {
    float wr[NOUT];   // wr[o] = w[j - o]: output o sees input j through weight j - o
#pragma unroll
    for (int o = 0; o < NOUT; ++o) wr[o] = 0.f;
#pragma unroll 8
    for (int j = 0; j < K + NOUT - 1; ++j) {
#pragma unroll
        for (int o = NOUT - 1; o > 0; --o) wr[o] = wr[o - 1];
        wr[0] = w[j];
        const float x = v[j * stride];
#pragma unroll
        for (int o = 0; o < NOUT; ++o) num[o] = fmaf(wr[o], x, num[o]);
        if (WITH_DEN) {
            const float y = f[j * stride];
#pragma unroll
            for (int o = 0; o < NOUT; ++o) den[o] = fmaf(wr[o], y, den[o]);
        }
    }
}

// copy a weight vector into shared memory with its zero tail; returns (in every thread of warp 0) its sum
__device__ __forceinline__ void stage_weights(const float* __restrict__ src, int K, float* __restrict__ dst) {
    for (int k = threadIdx.x; k < K + NOUT - 1; k += CONV_THREADS) dst[k] = (k < K) ? src[k] : 0.f;
}

// separable kernel: K(dy, dx) = ky[dy] kx[dx].  When the whole window of a tile is finite the normalisation is
// the constant sum(ky) sum(kx) -- the zeros beyond the image count as finite -- and its two passes are skipped.
template <bool WITH_DEN>
__device__ __forceinline__ void sep_passes(const float* val, const float* fin, float* mnum, float* mden,
                                           const float* skx, const float* sky, int K, int pitch, float (&num)[NOUT],
                                           float (&den)[NOUT]) {
    const int WY = TY + K - 1;
    // horizontal: WY rows x 4 groups of 8 columns
    for (int e = threadIdx.x; e < WY * (TX / NOUT); e += CONV_THREADS) {
        const int r = e >> 2, c0 = (e & 3) * NOUT;
        float n[NOUT], d[NOUT];
#pragma unroll
        for (int o = 0; o < NOUT; ++o) n[o] = d[o] = 0.f;
        slide<WITH_DEN>(val + r * pitch + c0, fin + r * pitch + c0, 1, K, skx, n, d);
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            mnum[r * MP + c0 + o] = n[o];
            if (WITH_DEN) mden[r * MP + c0 + o] = d[o];
        }
    }
    __syncthreads();
    // vertical: thread = column (lane) x group of 8 rows (warp)
    const int c = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * NOUT;
    slide<WITH_DEN>(mnum + r0 * MP + c, mden + r0 * MP + c, MP, K, sky, num, den);
}

__global__ void __launch_bounds__(CONV_THREADS)
    convolve_sep_kernel(const float* __restrict__ cube, float* __restrict__ out, const unsigned char* __restrict__ include,
                        int ny, int nx, int K, const float* __restrict__ kx, const float* __restrict__ ky, float ksum) {
    extern __shared__ float smem[];
    const int WX = TX + K - 1, WY = TY + K - 1, pitch = WX + 1;
    float* val = smem;                       // [WY][pitch]
    float* fin = val + WY * pitch;           // [WY][pitch]
    float* mnum = fin + WY * pitch;          // [WY][MP]   rows of the window, horizontally convolved
    float* mden = mnum + WY * MP;
    float* skx = mden + WY * MP;             // [K + NOUT - 1]
    float* sky = skx + K + NOUT - 1;
    const size_t plane = (size_t)ny * nx;
    const float* in = cube + blockIdx.z * plane;
    const int y0 = blockIdx.y * TY, x0 = blockIdx.x * TX;
    stage_weights(kx, K, skx);
    stage_weights(ky, K, sky);
    const bool mine = load_window(in, include, ny, nx, y0, x0, K, val, fin, pitch);
    const bool all_finite = __syncthreads_and(mine);   // also the barrier after the loads
    float num[NOUT], den[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) num[o] = den[o] = 0.f;
    if (all_finite) {
        sep_passes<false>(val, fin, mnum, mden, skx, sky, K, pitch, num, den);
#pragma unroll
        for (int o = 0; o < NOUT; ++o) den[o] = ksum;
    } else {
        sep_passes<true>(val, fin, mnum, mden, skx, sky, K, pitch, num, den);
    }
    const int c = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * NOUT, h = K >> 1;
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
        store_out(out + blockIdx.z * plane, ny, nx, y0 + r0 + o, x0 + c, num[o], den[o],
                  fin[(r0 + o + h) * pitch + c + h]);
}

// general kernel k2[K][K] (a rotated elliptical beam): each thread owns 8 vertically adjacent outputs and slides
// down its column once per kernel column dx (weights staged transposed, so a kernel column is contiguous)
__global__ void __launch_bounds__(CONV_THREADS)
    convolve_2d_kernel(const float* __restrict__ cube, float* __restrict__ out, const unsigned char* __restrict__ include,
                       int ny, int nx, int K, const float* __restrict__ k2, float ksum) {
    extern __shared__ float smem[];
    const int WX = TX + K - 1, WY = TY + K - 1, pitch = WX + 1, KP = K + NOUT - 1;
    float* val = smem;
    float* fin = val + WY * pitch;
    float* skT = fin + WY * pitch;           // [K dx][KP dy, zero tail]
    const size_t plane = (size_t)ny * nx;
    const float* in = cube + blockIdx.z * plane;
    const int y0 = blockIdx.y * TY, x0 = blockIdx.x * TX;
    for (int e = threadIdx.x; e < K * KP; e += CONV_THREADS) {
        const int dx = e / KP, dy = e - dx * KP;
        skT[e] = (dy < K) ? k2[dy * K + dx] : 0.f;
    }
    const bool mine = load_window(in, include, ny, nx, y0, x0, K, val, fin, pitch);
    const bool all_finite = __syncthreads_and(mine);
    const int c = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * NOUT, h = K >> 1;
    float num[NOUT], den[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) num[o] = den[o] = 0.f;
    if (all_finite) {
        for (int dx = 0; dx < K; ++dx) slide<false>(val + r0 * pitch + c + dx, fin, pitch, K, skT + dx * KP, num, den);
#pragma unroll
        for (int o = 0; o < NOUT; ++o) den[o] = ksum;
    } else {
        for (int dx = 0; dx < K; ++dx)
            slide<true>(val + r0 * pitch + c + dx, fin + r0 * pitch + c + dx, pitch, K, skT + dx * KP, num, den);
    }
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
        store_out(out + blockIdx.z * plane, ny, nx, y0 + r0 + o, x0 + c, num[o], den[o],
                  fin[(r0 + o + h) * pitch + c + h]);
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
                            const float* d_k2, double kernel_sum, float* d_out, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (nchan == 0 || ny == 0 || nx == 0) return B200FIT_OK;
    const bool sep = d_kx && d_ky && !d_k2, full = d_k2 && !d_kx && !d_ky;
    if (!d_cube || !d_out || d_cube == d_out || nchan < 0 || ny < 0 || nx < 0 || ksize < 1 || ksize > CONV_KMAX ||
        !(ksize & 1) || !(sep || full) || nchan > 65535 || !(kernel_sum > 0.0))
        return fail(ctx, B200FIT_E_ARG, "convolve_planes: bad argument");
    CK(cudaSetDevice(ctx->device));
    const size_t WX = TX + ksize - 1, WY = TY + ksize - 1, KP = ksize + NOUT - 1;
    const dim3 grid((nx + TX - 1) / TX, (ny + TY - 1) / TY, nchan);
    if (sep) {
        const size_t smem = sizeof(float) * (2 * WY * (WX + 1) + 2 * WY * MP + 2 * KP);
        CK(cudaFuncSetAttribute(convolve_sep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        convolve_sep_kernel<<<grid, CONV_THREADS, smem, (cudaStream_t)stream>>>(d_cube, d_out, d_include, ny, nx, ksize,
                                                                                d_kx, d_ky, (float)kernel_sum);
    } else {
        const size_t smem = sizeof(float) * (2 * WY * (WX + 1) + (size_t)ksize * KP);
        CK(cudaFuncSetAttribute(convolve_2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        convolve_2d_kernel<<<grid, CONV_THREADS, smem, (cudaStream_t)stream>>>(d_cube, d_out, d_include, ny, nx, ksize,
                                                                               d_k2, (float)kernel_sum);
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
