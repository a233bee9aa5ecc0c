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

#include <algorithm>
#include <cstdint>
#include <cstdlib>

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

// one 32 x 64 tile of plane ``bz`` whose first output is (y0, x0); all threads of the block
__device__ __forceinline__ void sep_tile(const float* __restrict__ cube, float* __restrict__ out,
                                         const unsigned char* __restrict__ include, int ny, int nx, int K,
                                         const float* __restrict__ kx, const float* __restrict__ ky, float ksum, int bz,
                                         int y0, int x0, float* smem) {
    const int WX = TX + K - 1, WY = TY + K - 1, pitch = WX + 1;
    float* val = smem;                       // [WY][pitch]
    float* fin = val + WY * pitch;           // [WY][pitch]
    float* mnum = fin + WY * pitch;          // [WY][MP]   rows of the window, horizontally convolved
    float* mden = mnum + WY * MP;
    float* skx = mden + WY * MP;             // [K + NOUT - 1]
    float* sky = skx + K + NOUT - 1;
    const size_t plane = (size_t)ny * nx;
    const float* in = cube + bz * plane;
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
        store_out(out + bz * plane, ny, nx, y0 + r0 + o, x0 + c, num[o], den[o], fin[(r0 + o + h) * pitch + c + h]);
}

__global__ void __launch_bounds__(CONV_THREADS)
    convolve_sep_kernel(const float* __restrict__ cube, float* __restrict__ out, const unsigned char* __restrict__ include,
                        int ny, int nx, int K, const float* __restrict__ kx, const float* __restrict__ ky, float ksum) {
    extern __shared__ float smem[];
    sep_tile(cube, out, include, ny, nx, K, kx, ky, ksum, blockIdx.z, blockIdx.y * TY, blockIdx.x * TX, smem);
}

// ---------------------------------------------------------------------------------------------------------------
// Fast separable path: planes without a pixel mask whose rows are 16-byte aligned (nx % 4 == 0), K >= 9.
//
// The generic kernel above spends its time on everything except the arithmetic (ncu profiles/r1n: 336 thread
// instructions per voxel of which 70 are the FMAs; staging loop with clamped scalar loads, two stores per staged
// value, a division per output).  Here
//   * the (64 + K - 1) x (64 + 2 ceil4(K/2)) window of a 64 x 64 tile goes to shared memory RAW, by 16-byte
//     cp.async copies whose source size is 0 beyond the image (hardware zero fill = boundary='fill'): one
//     instruction per four values, no sanitising pass, no finite-flag array;
//   * the passes run on PAIRS: vertical first, a thread owning two adjacent columns x 8 rows (one 64-bit
//     shared-memory load feeds 8 packed FMAs, ``fma.rn.f32x2``), its result stored TRANSPOSED so that the horizontal
//     pass -- a thread owning rows r and r + 32 x 8 columns, a warp's lanes on consecutive rows -- reads it without
//     bank conflicts; weights sit in registers, 8 at a time, and every product the loops issue is a real one (the
//     ramp-in / ramp-out of the 8-output window is unrolled statically per K mod 8, no zero tail);
//   * NaNs are not looked for: a NaN under an output's kernel makes that output NaN, the write-out loop sees it and
//     the tile is appended to a work list that convolve_sep_redo_fast_kernel (interpolating normalisation: the same
//     passes on values and on finite flags) redoes afterwards.
// Each lane of a packed FMA is the IEEE fp32 operation; per output the products are summed in the same order as in
// the generic kernel's all-finite path (j ascending, vertical after horizontal there / before here -- the two
// orders differ in rounding only, inside the 2e-6 of the parity tests).
// Measured on B200 (512 x 512 x 1000, K = 23; profiles/r2k_*): 0.89 ms = 2.36 TB/s of algorithmic bytes against
// 3.14 ms of the generic kernel.  What bounds it now is instruction issue: ncu counts a packed FMA as TWO issue slots
// (smsp__inst_executed = source-level instructions + one per FFMA2), so the 46 FMAs per output of a separable
// 23 x 23 kernel -- x 1.19 for the halo columns of the vertical pass -- are a floor of 0.40 ms at 128 FMA slots per
// SM and cycle, and the kernel spends 0.36 of its slots on everything else (cp.async address arithmetic,
// shared-memory loads, the write-out).  A persistent variant with the next tile's window in flight (two window
// buffers, 2 blocks per SM) was measured at 1.15 ms: the loads are already hidden behind the other three blocks of
// an SM, and 16 warps per SM leave the packed-FMA chains short of independent work.
constexpr int FT = 64;            // fast tile edge
constexpr int FPT = FT + 4;       // pitch of the transposed intermediate and of the output staging

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src, unsigned src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// 8 consecutive outputs of a 1-D correlation with K = 8 nb + KR weights, on pairs.  ``ld(j)`` returns the packed pair of
// inputs at offset j (0 <= j <= K + 6) from the first output; output o accumulates w[j - o] x(j) for 0 <= j - o < K,
// j ascending.  ``w``: shared memory, 16-byte aligned, readable up to 8 nb + 7.
template <int KR, class Load>
__device__ __forceinline__ void slide_pairs(Load ld, const float* __restrict__ w, int nb, f32x2 (&acc)[NOUT]) {
    static_assert(NOUT == 8, "the weight ring is 8 deep");
    float wp[8], wc[8], wn[8];    // weights of the previous / this / the next block of 8 (loaded one block ahead)
    auto load8 = [&](int m) {
        const float4 a = *reinterpret_cast<const float4*>(w + 8 * m), b = *reinterpret_cast<const float4*>(w + 8 * m + 4);
        wn[0] = a.x, wn[1] = a.y, wn[2] = a.z, wn[3] = a.w, wn[4] = b.x, wn[5] = b.y, wn[6] = b.z, wn[7] = b.w;
    };
    auto shift = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) wp[i] = wc[i], wc[i] = wn[i];
    };
    load8(0);
    shift();
    load8(1);                                 // (nb >= 1; the array is readable up to 8 nb + 7)
#pragma unroll
    for (int u = 0; u < 8; ++u) {             // ramp-in: input u reaches outputs 0 .. u
        const f32x2 x = ld(u);
#pragma unroll
        for (int o = 0; o <= u; ++o) acc[o] = fma2(pk2(wc[u - o], wc[u - o]), x, acc[o]);
    }
#pragma unroll 1
    for (int m = 1; m < nb; ++m) {            // every input reaches all 8 outputs
        shift();
        load8(m + 1);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const f32x2 x = ld(8 * m + u);
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const float wv = (u >= o) ? wc[u - o] : wp[8 + u - o];
                acc[o] = fma2(pk2(wv, wv), x, acc[o]);
            }
        }
    }
    shift();                                  // wc: weights 8 nb .. K - 1 (what follows them is never used)
#pragma unroll
    for (int t = 0; t < KR + 7; ++t) {        // ramp-out: input 8 nb + t reaches outputs o >= t - KR + 1
        const f32x2 x = ld(8 * nb + t);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            if (o >= t - KR + 1) {
                const int rel = t - o;        // weight index relative to 8 nb: -7 .. KR - 1
                const float wv = (rel >= 0) ? wc[rel >= 0 ? rel : 0] : wp[rel < 0 ? 8 + rel : 0];
                acc[o] = fma2(pk2(wv, wv), x, acc[o]);
            }
        }
    }
}

// ---- the phases of a fast tile (all threads of the block) -------------------------------------------------------
struct FastGeom {
    int h, H4, WXA, WY, nb;
    __device__ __forceinline__ explicit FastGeom(int K) : h(K >> 1), H4((h + 3) & ~3), WXA(FT + 2 * H4), WY(FT + K - 1), nb(K >> 3) {}
};

// the window of the tile at (y0, x0) of plane ``in``, four values per copy (a lane owns one 16-byte segment of every
// row its warp stages); segments beyond the image are zero filled by the copy engine.  One commit group.
__device__ __forceinline__ void stage_window(const FastGeom& g, const float* __restrict__ in, int ny, int nx, int y0, int x0,
                                             float* raw) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = x0 - g.H4 + 4 * lane;
    const bool col_ok = x >= 0 && x < nx;                            // nx % 4 == 0: a segment is inside or outside
    const float* col = in + (col_ok ? x : 0);
    unsigned dst = (unsigned)__cvta_generic_to_shared(raw) + (unsigned)((warp * g.WXA + 4 * lane) * sizeof(float));
    if (4 * lane < g.WXA) {
        int y = y0 + warp - g.h, off = y * nx;                       // (a plane holds < 2^31 voxels: checked by the host)
        for (int r = warp; r < g.WY; r += CONV_THREADS / 32) {
            const bool ok = col_ok && y >= 0 && y < ny;
            cp_async16(dst, ok ? col + off : in, ok ? 16u : 0u);
            dst += (CONV_THREADS / 32) * g.WXA * (unsigned)sizeof(float);
            y += CONV_THREADS / 32;
            off += (CONV_THREADS / 32) * nx;
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// vertical pass: thread = (column pair, group of 8 rows), lanes on consecutive pairs; result transposed into ``mid``
template <int KR>
__device__ __forceinline__ void pass_vertical(const FastGeom& g, const float* raw, float* mid, const float* sky) {
    const int npair = g.WXA >> 1;
    for (int e = threadIdx.x; e < npair * (FT / NOUT); e += CONV_THREADS) {
        const int rg = e / npair, cp = e - rg * npair;
        const float* x0p = raw + (rg * NOUT) * g.WXA + 2 * cp;
        const int WXA = g.WXA;
        f32x2 acc[NOUT];
#pragma unroll
        for (int o = 0; o < NOUT; ++o) acc[o] = 0ull;
        slide_pairs<KR>([&](int j) { return *reinterpret_cast<const f32x2*>(x0p + j * WXA); }, sky, g.nb, acc);
        float lo[NOUT], hi[NOUT];
#pragma unroll
        for (int o = 0; o < NOUT; ++o) upk2(acc[o], lo[o], hi[o]);
        float* d0 = mid + (2 * cp) * FPT + rg * NOUT;
        *reinterpret_cast<float4*>(d0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4*>(d0 + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
        *reinterpret_cast<float4*>(d0 + FPT) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(d0 + FPT + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
    }
}

// horizontal pass: warp = group of 8 output columns, lane = rows (lane, lane + 32); acc[o] = {row lane, row lane + 32}
// of output column 8 warp + o
template <int KR>
__device__ __forceinline__ void pass_horizontal(const FastGeom& g, const float* mid, const float* skx, f32x2 (&acc)[NOUT]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* x0p = mid + (g.H4 - g.h + warp * NOUT) * FPT + lane;
#pragma unroll
    for (int o = 0; o < NOUT; ++o) acc[o] = 0ull;
    slide_pairs<KR>([&](int j) { return pk2(x0p[j * FPT], x0p[j * FPT + 32]); }, skx, g.nb, acc);
}

// this thread's 16 outputs into the [64][FPT] staging tile
__device__ __forceinline__ void stage_outputs(float* tile, const float (&lo)[NOUT], const float (&hi)[NOUT]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* d0 = tile + lane * FPT + warp * NOUT;
    *reinterpret_cast<float4*>(d0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<float4*>(d0 + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
    *reinterpret_cast<float4*>(d0 + 32 * FPT) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(d0 + 32 * FPT + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
}

// coalesced write-out of the staging tile; returns whether this thread saw a non-finite output
__device__ __forceinline__ bool write_tile(const float* tile, float* __restrict__ op, int ny, int nx, int y0, int x0) {
    bool bad = false;
#pragma unroll
    for (int i = 0; i < (FT * FT / 4) / CONV_THREADS; ++i) {
        const int idx = threadIdx.x + i * CONV_THREADS, row = idx >> 4, q = idx & 15;
        const float4 v = *reinterpret_cast<const float4*>(tile + row * FPT + 4 * q);
        const int y = y0 + row, x = x0 + 4 * q;
        if (y < ny && x < nx) {
            const float s = (v.x - v.x) + (v.y - v.y) + (v.z - v.z) + (v.w - v.w);   // 0 iff all four are finite
            bad = bad || !(s == 0.f);
            *reinterpret_cast<float4*>(op + (size_t)y * nx + x) = v;
        }
    }
    return bad;
}

template <int KR>
__global__ void __launch_bounds__(CONV_THREADS, 4)
    convolve_sep_fast_kernel(const float* __restrict__ cube, float* __restrict__ out, int ny, int nx, int K,
                             const float* __restrict__ kx, const float* __restrict__ ky, float inv_ksum,
                             unsigned* __restrict__ redo) {
    extern __shared__ __align__(16) float smem[];
    const FastGeom g(K);
    const int KW = (K + 15) & ~7;                      // weights + room for the last 8-wide load
    float* raw = smem;                                  // [WY][WXA] as it lies in the plane; later the output staging
    float* mid = raw + g.WY * g.WXA;                    // [WXA][FPT] vertically convolved, transposed
    float* skx = mid + g.WXA * FPT;                     // [KW]
    float* sky = skx + KW;
    const size_t plane = (size_t)ny * nx;
    const int y0 = blockIdx.y * FT, x0 = blockIdx.x * FT;
    for (int k = threadIdx.x; k < KW; k += CONV_THREADS) {
        skx[k] = (k < K) ? kx[k] : 0.f;
        sky[k] = (k < K) ? ky[k] : 0.f;
    }
    stage_window(g, cube + blockIdx.z * plane, ny, nx, y0, x0, raw);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    pass_vertical<KR>(g, raw, mid, sky);
    __syncthreads();
    f32x2 acc[NOUT];
    pass_horizontal<KR>(g, mid, skx, acc);
    float lo[NOUT], hi[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
        upk2(acc[o], lo[o], hi[o]);
        lo[o] *= inv_ksum;
        hi[o] *= inv_ksum;
    }
    stage_outputs(raw, lo, hi);                         // (the window is dead since the barrier above)
    __syncthreads();
    // any non-finite output sends the tile to the two-accumulator kernel below
    const bool bad = write_tile(raw, out + blockIdx.z * plane, ny, nx, y0, x0);
    if (__syncthreads_or(bad) && threadIdx.x == 0)
        redo[1 + atomicAdd(redo, 1u)] = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
}

// The tiles the fast kernel gave up on (a NaN under some output's kernel), with the interpolating normalisation of the
// generic kernel: out = sum K v / sum K f, v = value or 0, f = 1 where the voxel is finite (and beyond the image), NaN
// where the voxel itself is NaN or nothing finite is under the kernel.  The same passes run twice on the same buffers:
// once on the sanitised values, once on the finite flags -- whose window is on its way in (second cp.async group)
// while the first horizontal pass runs.  Persistent blocks over the work list.
template <int KR>
__global__ void __launch_bounds__(CONV_THREADS, 2)
    convolve_sep_redo_fast_kernel(const float* __restrict__ cube, float* __restrict__ out, int ny, int nx, int K,
                                  const float* __restrict__ kx, const float* __restrict__ ky,
                                  const unsigned* __restrict__ redo, int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) float smem[];
    const FastGeom g(K);
    const int KW = (K + 15) & ~7;
    float* raw = smem;
    float* mid = raw + g.WY * g.WXA;
    float* skx = mid + g.WXA * FPT;
    float* sky = skx + KW;
    const size_t plane = (size_t)ny * nx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < KW; k += CONV_THREADS) {
        skx[k] = (k < K) ? kx[k] : 0.f;
        sky[k] = (k < K) ? ky[k] : 0.f;
    }
    const unsigned n = redo[0];
    for (unsigned i = blockIdx.x; i < n; i += gridDim.x) {
        const unsigned t = redo[1 + i];
        const int bx = t % tiles_x, by = (t / tiles_x) % tiles_y, bz = t / (tiles_x * tiles_y);
        const int y0 = by * FT, x0 = bx * FT;
        const float* in = cube + bz * plane;
        stage_window(g, in, ny, nx, y0, x0, raw);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // is each of this thread's 16 outputs a finite voxel itself?  (read before the window is sanitised)
        unsigned selfok = 0;
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            const float a = raw[(lane + g.h) * g.WXA + g.H4 + warp * NOUT + o];
            const float b = raw[(lane + 32 + g.h) * g.WXA + g.H4 + warp * NOUT + o];
            selfok |= (a == a ? 1u : 0u) << o;
            selfok |= (b == b ? 1u : 0u) << (8 + o);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < g.WY * g.WXA; e += CONV_THREADS) {
            const float v = raw[e];
            raw[e] = (v == v) ? v : 0.f;
        }
        __syncthreads();
        pass_vertical<KR>(g, raw, mid, sky);
        __syncthreads();
        stage_window(g, in, ny, nx, y0, x0, raw);          // the same window again, for the flags (L2-resident)
        f32x2 num[NOUT];
        pass_horizontal<KR>(g, mid, skx, num);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                    // every thread is done with ``mid`` and the window is in
        for (int e = threadIdx.x; e < g.WY * g.WXA; e += CONV_THREADS) {
            const float v = raw[e];
            raw[e] = (v == v) ? 1.f : 0.f;                  // (zero fill beyond the image: a finite zero)
        }
        __syncthreads();
        pass_vertical<KR>(g, raw, mid, sky);
        __syncthreads();
        f32x2 den[NOUT];
        pass_horizontal<KR>(g, mid, skx, den);
        float lo[NOUT], hi[NOUT];
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
            float nl, nh, dl, dh;
            upk2(num[o], nl, nh);
            upk2(den[o], dl, dh);
            lo[o] = (((selfok >> o) & 1u) && dl > 0.f) ? nl / dl : __int_as_float(0x7fc00000);
            hi[o] = (((selfok >> (8 + o)) & 1u) && dh > 0.f) ? nh / dh : __int_as_float(0x7fc00000);
        }
        stage_outputs(raw, lo, hi);                         // (the flags are dead since the barrier above)
        __syncthreads();
        write_tile(raw, out + bz * plane, ny, nx, y0, x0);
        __syncthreads();                                    // the staging tile is read: the next window may land
    }
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
        !(ksize & 1) || !(sep || full) || nchan > 65535 || !(kernel_sum > 0.0) || (long long)ny * nx >= (1ll << 31))
        return fail(ctx, B200FIT_E_ARG, "convolve_planes: bad argument");
    CK(cudaSetDevice(ctx->device));
    const size_t WX = TX + ksize - 1, WY = TY + ksize - 1, KP = ksize + NOUT - 1;
    const dim3 grid((nx + TX - 1) / TX, (ny + TY - 1) / TY, nchan);
    if (sep) {
        const size_t smem = sizeof(float) * (2 * WY * (WX + 1) + 2 * WY * MP + 2 * KP);
        CK(cudaFuncSetAttribute(convolve_sep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const bool fast = !d_include && ksize >= 9 && (nx % 4) == 0 && ((uintptr_t)d_cube % 16) == 0 &&
                          ((uintptr_t)d_out % 16) == 0 && !getenv("B200FIT_CONV_GENERIC");
        if (fast) {
            // fast kernel over 64 x 64 tiles, then the generic one over the tiles it flagged (those holding NaNs)
            const int h = ksize >> 1, H4 = (h + 3) & ~3, WXA = FT + 2 * H4, WYF = FT + ksize - 1, KW = (ksize + 15) & ~7;
            const size_t fsmem = sizeof(float) * ((size_t)WYF * WXA + (size_t)WXA * FPT + 2 * KW);
            const dim3 fgrid((nx + FT - 1) / FT, (ny + FT - 1) / FT, nchan);
            const size_t ntiles = (size_t)fgrid.x * fgrid.y * fgrid.z;
            if (ctx->redo_cap < ntiles + 1) {          // work list of the context, grown on demand (convolutions of one
                if (ctx->d_redo) {                      // context are issued one at a time, on one stream)
                    CK(cudaDeviceSynchronize());
                    CK(cudaFree(ctx->d_redo));
                    ctx->d_redo = nullptr, ctx->redo_cap = 0;
                }
                CK(cudaMalloc((void**)&ctx->d_redo, (ntiles + 1) * sizeof(unsigned)));
                ctx->redo_cap = ntiles + 1;
            }
            unsigned* d_redo = ctx->d_redo;
            CK(cudaMemsetAsync(d_redo, 0, sizeof(unsigned), (cudaStream_t)stream));
            const float inv = (float)(1.0 / kernel_sum);
#define B200FIT_LAUNCH_FAST(KR)                                                                                          \
    do {                                                                                                                 \
        CK(cudaFuncSetAttribute(convolve_sep_fast_kernel<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem)); \
        convolve_sep_fast_kernel<KR><<<fgrid, CONV_THREADS, fsmem, (cudaStream_t)stream>>>(d_cube, d_out, ny, nx, ksize, \
                                                                                            d_kx, d_ky, inv, d_redo);   \
    } while (0)
            switch (ksize & 7) {
                case 1: B200FIT_LAUNCH_FAST(1); break;
                case 3: B200FIT_LAUNCH_FAST(3); break;
                case 5: B200FIT_LAUNCH_FAST(5); break;
                default: B200FIT_LAUNCH_FAST(7); break;
            }
#undef B200FIT_LAUNCH_FAST
            CK(cudaGetLastError());
#define B200FIT_LAUNCH_REDO(KR)                                                                                          \
    do {                                                                                                                 \
        CK(cudaFuncSetAttribute(convolve_sep_redo_fast_kernel<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                (int)fsmem));                                                                            \
        convolve_sep_redo_fast_kernel<KR><<<ctx->num_sms * 2, CONV_THREADS, fsmem, (cudaStream_t)stream>>>(              \
            d_cube, d_out, ny, nx, ksize, d_kx, d_ky, d_redo, (int)fgrid.x, (int)fgrid.y);                               \
    } while (0)
            switch (ksize & 7) {
                case 1: B200FIT_LAUNCH_REDO(1); break;
                case 3: B200FIT_LAUNCH_REDO(3); break;
                case 5: B200FIT_LAUNCH_REDO(5); break;
                default: B200FIT_LAUNCH_REDO(7); break;
            }
#undef B200FIT_LAUNCH_REDO
            CK(cudaGetLastError());
            ctx->launches += 2;
            return B200FIT_OK;
        }
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
