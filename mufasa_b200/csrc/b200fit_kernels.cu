// b200fit_kernels.cu -- sm_100a kernels + C ABI of the B200-native per-pixel spectral fitter.
//
// Hot path replaced (reference file:line):
//   pcube.fiteach -> Specfit -> SpectralModel.fitter -> mpfit       mufasa/multi_v_fit.py:374-393, :192-206
//   HyperfineModel.multi_v_spectrum / _single_spectrum_hf            mufasa/spec_models/HyperfineModel.py:22-109
//   pcube.get_modelcube, UltraCube.get_rss / expand_mask / AICc ...  mufasa/UltraCube.py:322-325, 1384-1475, 1658-1698
//
// Design:
//   * the fit is a ROUND LOOP of small uniform kernels over a compacted list of the pixels still iterating
//     (fit_prologue / fit_eval / fit_solve / fit_finalize below): a warp (late rounds: a block) evaluates the fp32
//     model and its analytic Jacobian of one pixel over the chunks a hyperfine window touches and accumulates
//     J^T J / J^T r / chi2 (packed fp32 FMAs per lane, fp64 across lanes); P lanes advance the pixel's LM state
//     machine (fp64 Cholesky on the damped normal equations, mpfit-style limits; b200fit_lanes.cuh, serial
//     restatement in b200fit_common.h).  Several fits -- or parts of one -- run in flight together.
//   * model / aicc / mask / stat kernels: fp64 model in the reference's frequency-space operation order (the AICc
//     sample mask is ``model > 0``, decided by fp64 round-off), one warp per pixel.
//   * gather_kernel: FITS channel-major cube -> compacted pixel-major rows (smem tile transpose).
// No tensor cores: the per-pixel systems are 4x4 / 8x8.  No CPU fallback anywhere in this library.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b200fit_host.h"
#include "b200fit_lanes.cuh"

namespace b200fit {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------------------
// warp helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------
// The fit: a ROUND LOOP of two small, uniform kernels over a compacted list of the pixels still iterating.
//   fit_prologue_kernel  warp per pixel, once: weights (std of the finite channels), sum y^2, LM initialisation,
//                        active list 0
//   fit_eval_kernel      warp per ACTIVE pixel: fp32 model + analytic Jacobian over the hyperfine windows,
//                        J^T J / J^T r / chi2 (fp32 per lane, fp64 across lanes) at the pixel's trial point
//   fit_solve_kernel     P lanes per active pixel: one step of the LM state machine (fp64 Cholesky trust-region
//                        solve, mpfit limits); finished pixels leave the list, the rest join the next one
//   fit_finalize_kernel  once: parameter errors from (J^T J)^-1 and the result rows
// Why not one fused persistent kernel (rounds 1a-1c, profiles/): the solver is ~5k instructions of branchy fp64
// scalar code per step.  Fused, it ran on 1 (v1) or ~4 (v2, pixel groups) lanes of a warp, and evaluation +
// solver code together (107-263 KB) thrashed the 32 KB instruction cache -- >50 % of all stall samples were
// "no instruction".  Split, every warp of a kernel runs the same few KB of code, the solver runs 32 pixels per
// warp, pixels that converge early free their lanes at once (iteration counts differ by >10x), and the state
// that crosses kernels (~1.1 KB per pixel and round) is noise next to the arithmetic: the path is compute bound.
// What ONE evaluation reads of a pixel's state, contiguous and 16-byte aligned so that the evaluation kernel stages
// it in shared memory with a single bulk-async copy next to the pixel's spectrum row (fit_eval_kernel).
template <int NC>
struct alignas(16) EvalIn {
    CompCoef cc[NC];             // coefficients of the trial point st.pt (filled by the prologue / the solver)
    double v[NC];                // st.pt[4 c]: velocity of component c at the trial point
    double inv_sigma[NC];        // 1 / st.pt[4 c + 1]
    double sumsq, err;           // sum of y^2 over the finite channels; weight (std of the finite channels)
};
static_assert(sizeof(EvalIn<1>) % 16 == 0 && sizeof(EvalIn<2>) % 16 == 0, "bulk copies move multiples of 16 bytes");

template <int P>
struct alignas(16) PixState {
    EvalIn<P / 4> in;
    LMState<P> st;
    unsigned n_gauss, n_chunks;  // roofline accounting
    int fitted, pad_;            // 1 once the LM state is initialised (0: bad guess / no weight -> zeros, status 0)
};

struct DeviceLimits {            // box limits of the parameters, passed by value (prepare_guesses_kernel)
    double lo[8], hi[8];
};

struct FitQueues {               // head of the workspace
    unsigned int count[2];       // sizes of the ping-pong active lists
    unsigned int pad[62];
};

constexpr int EVAL_WARPS = 4;    // warps per block, evaluation / prologue kernels
// From this round on a block (not a warp) evaluates a pixel (fit_eval_kernel).  A team spends ~1.8x the warp-time of
// a single warp on a pixel (set-up, reduction and barriers do not shrink), so it pays once a round holds fewer
// pixels than warp slots / 4: measured on 65 536 NH3 spectra, round 96 for 1 component (evaluation 8.6 -> 7.5 ms),
// while 2-component fits keep ~3000 pixels (the ones running to maxiter) until round ~200, where teams from round
// 96 cost +4 %; they switch only for the final stretch.
template <int NC>
constexpr int fit_team_round() { return NC == 1 ? 96 : 288; }
constexpr int SOLVE_THREADS = 64;

// ---- packed fp32 pairs --------------------------------------------------------------------------------------
// sm_100 executes fma/add/mul on two fp32 lanes of a 64-bit register pair per instruction (SASS FFMA2 / FADD2 /
// FMUL2, PTX ``*.f32x2``).  Each lane is the IEEE fp32 operation, so results equal the scalar code's; the
// J^T J / J^T r updates of the evaluation kernel take 25 packed FMAs per channel instead of 45 scalar ones
// (2c); a scalar broadcast operand costs nothing (``R.F32`` modifier).  Measured: -3 % on the 2c evaluation.
// Packing the hyperfine gaussians two lines at a time as well was measured and dropped: odd line groups need a
// null line, and the halved instruction-level parallelism cost more than the saved issue slots (+5 % on 1c).
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

template <int NC>
struct alignas(16) EvalScratch {
    static constexpr int P = 4 * NC;
    static constexpr int NH = P * (P + 1) / 2;
    static constexpr int NACC = NH + P + 1;
    static constexpr int RED_ROWS = (NACC < 24) ? NACC : 24;
    alignas(16) LineCoef lines[NC][MAXHF_FIT];   // 16-byte records (float4 loads)
    float wlo[NC][MAXHF_FIT], whi[NC][MAXHF_FIT];
    float red[RED_ROWS][33];                     // cross-lane reduction scratch (+1 pad: conflict-free both ways)
};

__device__ __forceinline__ void load_fit_tables(const DeviceProblem& pr, FitTables& tb) {
    if (threadIdx.x < MAXHF_FIT) {
        const int k = threadIdx.x;
        tb.rho[k] = pr.rho[k];
        tb.c1mrho[k] = pr.c1mrho[k];
        tb.kdv[k] = pr.kdv[k];
        tb.lw[k] = pr.lw[k];
    }
    if (threadIdx.x <= MAXHF_FIT) tb.gstart[threadIdx.x] = pr.gstart[threadIdx.x];
    if (threadIdx.x < 8) {
        tb.lo[threadIdx.x] = pr.lo[threadIdx.x];
        tb.hi[threadIdx.x] = pr.hi[threadIdx.x];
    }
    __syncthreads();
}

// Coefficients of the trial point the next evaluation will use.
template <int NC>
__device__ __forceinline__ void prepare_trial(const DeviceProblem& pr, PixState<4 * NC>& ps) {
#pragma unroll 1
    for (int c = 0; c < NC; ++c) {
        CompCoef cc;
        comp_coefs(pr.T0c, pr.dT0, pr.Jbg0, pr.Jbg1, &ps.st.pt[4 * c], cc);
        ps.in.cc[c] = cc;
        ps.in.v[c] = ps.st.pt[4 * c];
        ps.in.inv_sigma[c] = 1.0 / ps.st.pt[4 * c + 1];
    }
}

template <int P>
__device__ __forceinline__ void write_result(const PixState<P>& ps, long long pix, const LMConfig& cf, bool fitted,
                                             double* __restrict__ params, double* __restrict__ perr,
                                             double* __restrict__ chi2, double* __restrict__ err_used,
                                             int* __restrict__ status, unsigned* __restrict__ counts) {
    const bool good = fitted && ps.st.status > 0;
    const double err2 = ps.in.err * ps.in.err;
    double pe[P];
    if (good) lm_errors<P>(ps.st, cf, err2, pe);
#pragma unroll 1
    for (int j = 0; j < P; ++j) {
        params[pix * P + j] = good ? ps.st.p[j] : 0.0;
        perr[pix * P + j] = good ? pe[j] : 0.0;
    }
    chi2[pix] = good ? lm_reported_chi2<P>(ps.st) / err2 : 0.0;
    err_used[pix] = ps.in.err;
    status[pix] = fitted ? ps.st.status : 0;
    if (counts) {
        counts[4 * pix + 0] = fitted ? (unsigned)ps.st.nevals : 0u;
        counts[4 * pix + 1] = fitted ? (unsigned)ps.st.nrej : 0u;
        counts[4 * pix + 2] = fitted ? ps.n_gauss : 0u;
        counts[4 * pix + 3] = fitted ? ps.n_chunks : 0u;
    }
}

// ---- prologue: warp per pixel -------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(EVAL_WARPS * 32)
    fit_prologue_kernel(const DeviceProblem pr, const float* __restrict__ spectra, long long stride,
                        const long long* __restrict__ row_index, const double* __restrict__ guesses,
                        const double* __restrict__ err_in, PixState<4 * NC>* __restrict__ states,
                        unsigned* __restrict__ list0, FitQueues* __restrict__ fq, double* __restrict__ params,
                        double* __restrict__ perr, double* __restrict__ chi2, double* __restrict__ err_used,
                        int* __restrict__ status, unsigned* __restrict__ counts) {
    constexpr int P = 4 * NC;
    __shared__ FitTables tb;
    load_fit_tables(pr, tb);
    const LMConfig cf{tb.lo, tb.hi, pr.ftol, pr.max_iter};
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * EVAL_WARPS + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * EVAL_WARPS;
    const int nchan = pr.nchan;
    for (long long px = warp0; px < pr.npix; px += nwarps) {
        // weights err = population std of the finite channels (Cube.fiteach), sum y^2, peak for signal_cut:
        // one coalesced pass, fp64 accumulators
        const long long row = row_index ? row_index[px] : px;
        const float* sp = spectra + row * stride;
        double s1 = 0.0, sq = 0.0, ymax = -INFINITY;
        int n = 0;
#pragma unroll 4
        for (int i = lane; i < nchan; i += 32) {
            const float y = __ldg(sp + i);
            if (y == y) {
                const double yd = (double)y;
                s1 += yd;
                sq = fma(yd, yd, sq);
                ++n;
                ymax = fmax(ymax, yd);
            }
        }
        s1 = warp_sum(s1);
        sq = warp_sum(sq);
        n = warp_sum_i(n);
        ymax = warp_max(ymax);
        if (lane == 0) {
            PixState<P>& ps = states[px];
            const double mean = (n > 0) ? s1 / n : 0.0;
            const double var = (n > 0) ? fmax(sq / n - mean * mean, 0.0) : NAN;
            double err = sqrt(var);
            if (pr.weight_mode == 1 && err_in != nullptr) err = err_in[px];
            ps.in.err = err;
            ps.in.sumsq = sq;
            ps.n_gauss = 0;
            ps.n_chunks = 0;
            ps.fitted = 0;
            bool go = (err > 0.0) && !isinf(err);
            if (go && pr.signal_cut > 0.0 && !(ymax / err >= pr.signal_cut)) go = false;
            const bool ok = go && lm_init<P>(ps.st, guesses + px * P, cf);
            if (ok) {
                ps.fitted = 1;
                prepare_trial<NC>(pr, ps);
                list0[atomicAdd(&fq->count[0], 1u)] = (unsigned)px;
            } else {   // not fitted: zeros, like pyspeckit's blank parcube
                write_result<P>(ps, px, cf, false, params, perr, chi2, err_used, status, counts);
            }
        }
    }
}

// ---- evaluation: warp per active pixel -----------------------------------------------------------------------
//   chi2 = sum_finite y^2 + sum_{channels inside a window} m (m - 2 y)      (exactly sum (y - m)^2: m == 0 elsewhere)
// Templated on the species so that the hyperfine GROUPS (runs of neighbouring lines that share one channel window)
// are compile-time: per 32-channel chunk and component the warp tests NG group bits and runs fully unrolled
// gaussian code for the groups that reach the chunk -- no loop control, no index arithmetic (the dynamic
// ``for k in [klo, khi)`` loop of the previous version was 15 % of all executed instructions, profiles/r1k).
template <int MODEL>
struct LineGroups;
template <>
struct LineGroups<B200FIT_MODEL_NH3_11> {   // sorted by velocity offset: -19.5 (2) | -7.5 (3) | main (8) | +7.5 (3) | +19.6 (2)
    static constexpr int NG = 5;
    __host__ __device__ static constexpr int start(int g) { return g == 0 ? 0 : g == 1 ? 2 : g == 2 ? 5 : g == 3 ? 13 : g == 4 ? 16 : 18; }
};
template <>
struct LineGroups<B200FIT_MODEL_N2HP_10> {  // -8.0 (1) | -0.6 .. 1.0 (3) | 5.5 .. 6.9 (3)
    static constexpr int NG = 3;
    __host__ __device__ static constexpr int start(int g) { return g == 0 ? 0 : g == 1 ? 1 : g == 2 ? 4 : 7; }
};

// hyperfine lines a chunk with group mask ``pk`` (bit 8c + g, c < 2) evaluates per channel: the sizes of its groups
// summed -- groups of equal size share one population count (roofline accounting only; ``pk`` is warp-uniform)
template <int MODEL, int NC>
__device__ __forceinline__ unsigned group_lines(unsigned pk);
template <>
__device__ __forceinline__ unsigned group_lines<B200FIT_MODEL_NH3_11, 1>(unsigned pk) {   // 2 | 3 | 8 | 3 | 2
    return 2u * __popc(pk & 0x11u) + 3u * __popc(pk & 0x0au) + 8u * __popc(pk & 0x04u);
}
template <>
__device__ __forceinline__ unsigned group_lines<B200FIT_MODEL_NH3_11, 2>(unsigned pk) {
    return 2u * __popc(pk & 0x1111u) + 3u * __popc(pk & 0x0a0au) + 8u * __popc(pk & 0x0404u);
}
template <>
__device__ __forceinline__ unsigned group_lines<B200FIT_MODEL_N2HP_10, 1>(unsigned pk) {  // 1 | 3 | 3
    return __popc(pk & 0x01u) + 3u * __popc(pk & 0x06u);
}
template <>
__device__ __forceinline__ unsigned group_lines<B200FIT_MODEL_N2HP_10, 2>(unsigned pk) {
    return __popc(pk & 0x0101u) + 3u * __popc(pk & 0x0606u);
}
static_assert(LineGroups<B200FIT_MODEL_NH3_11>::start(1) == 2 && LineGroups<B200FIT_MODEL_NH3_11>::start(2) == 5 &&
                  LineGroups<B200FIT_MODEL_NH3_11>::start(3) == 13 && LineGroups<B200FIT_MODEL_NH3_11>::start(4) == 16 &&
                  LineGroups<B200FIT_MODEL_NH3_11>::start(5) == 18 && LineGroups<B200FIT_MODEL_N2HP_10>::start(1) == 1 &&
                  LineGroups<B200FIT_MODEL_N2HP_10>::start(2) == 4 && LineGroups<B200FIT_MODEL_N2HP_10>::start(3) == 7,
              "group_lines hard-codes the group sizes");

//
// TEAM = warps per pixel.  TEAM = 1: a warp owns a pixel.  TEAM = EVAL_WARPS: a block owns a pixel and its warps take
// the active chunks round-robin -- a quarter of the evaluation latency, which is what the late rounds of a fit are
// made of (a few thousand pixels crawling to maxiter; profiles/r1m launch list).  The two variants sum the fp32
// per-lane accumulators in different groupings, so WHICH one evaluates a pixel must not depend on what else is being
// fitted: the host switches at a fixed round number (fit_team_round<NC>()), and since every active pixel is evaluated
// exactly once per round that is a function of the pixel's own evaluation count -- results stay independent of
// how the pixels are partitioned over calls or GPUs.
//
// Staging.  A unit (warp or team) is a small software pipeline over its pixels: while it evaluates pixel k, the
// spectrum row (stride x 4 bytes) and the EvalIn block of pixel k + 1 are already on their way into the other half of
// its shared-memory buffer -- two 1-D bulk-async copies (cp.async.bulk, the TMA engine) issued by one lane and
// completed on an mbarrier, so no thread waits on L2 / HBM latency inside the chunk loop (profiles/r1o: the scalar
// __ldg loads of the previous version were the second largest stall, long_scoreboard 1.03 cycles per instruction).
//
// Uniform control.  Which line groups reach a chunk is a per-chunk bit mask held by the lane that owns the chunk;
// the mask of the chunk being processed reaches every lane through redux.sync (warp-wide OR), whose result the
// compiler keeps in a UNIFORM register: the per-group tests are uniform-datapath instructions and plain branches.
// (The previous version proved uniformity to the compiler with one __any_sync vote per group and component -- 29 %
// of all executed warp instructions, profiles/r1o_eval2c_source_summary.txt.)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

template <int NC>
__host__ __device__ constexpr size_t eval_buf_bytes(long long stride) {   // one pipeline stage of one unit
    return (size_t)stride * sizeof(float) + sizeof(EvalIn<NC>);
}
template <int NC, int TEAM>
static size_t eval_dyn_smem(long long stride) {
    return ((TEAM > 1) ? 1 : EVAL_WARPS) * 2 * eval_buf_bytes<NC>(stride);
}

template <int NC, int MODEL, int TEAM>
// blocks per SM: 5 (96 registers) help the 1-component kernel; the 2-component one compiles to 96 registers with
// next to no spills as well, but runs 16 % slower there (24.1 vs 20.7 ms per fit) than at 4 blocks / 128 registers
__global__ void __launch_bounds__(EVAL_WARPS * 32, (NC == 1) ? 5 : 4)
    fit_eval_kernel(const DeviceProblem pr, const float* __restrict__ spectra, long long stride,
                    const long long* __restrict__ row_index, PixState<4 * NC>* __restrict__ states,
                    const unsigned* __restrict__ list, FitQueues* __restrict__ fq, int cur) {
    static_assert(TEAM == 1 || TEAM == EVAL_WARPS, "a pixel belongs to one warp or to one block");
    constexpr int P = 4 * NC;
    constexpr int NH = P * (P + 1) / 2;
    constexpr int NACC = NH + P + 1;
    constexpr int RR = EvalScratch<NC>::RED_ROWS;
    constexpr int UNITS = (TEAM > 1) ? 1 : EVAL_WARPS;   // pixels in flight per block
    using LG = LineGroups<MODEL>;
    constexpr int NG = LG::NG;
    extern __shared__ __align__(128) unsigned char stage_smem[];   // UNITS x 2 stages x (row, EvalIn)
    __shared__ FitTables tb;
    __shared__ EvalScratch<NC> scratch[EVAL_WARPS];
    __shared__ double part[(TEAM > 1) ? EVAL_WARPS : 1][NACC + 2];   // per-warp partial sums of a team
    __shared__ __align__(8) unsigned long long bars[UNITS][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int unit = (TEAM > 1) ? 0 : warp;
    const bool leader = (TEAM > 1) ? (threadIdx.x == 0) : (lane == 0);   // issues this unit's copies
    const unsigned count = fq->count[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) fq->count[cur ^ 1] = 0;   // the list this round's solver appends to
    // (the grid is sized for a full list: blocks without a pixel leave before any set-up)
    if (((TEAM > 1) ? blockIdx.x : blockIdx.x * EVAL_WARPS) >= count) return;
    const unsigned row_bytes = (unsigned)(stride * sizeof(float));
    const unsigned buf_bytes = (unsigned)eval_buf_bytes<NC>(stride);
    unsigned char* const ubase = stage_smem + (size_t)unit * 2 * buf_bytes;
    if (leader) {
        mbar_init(smem_u32(&bars[unit][0]), 1);
        mbar_init(smem_u32(&bars[unit][1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_fit_tables(pr, tb);                                // (ends with __syncthreads: the barriers are visible)
    EvalScratch<NC>& w = scratch[(TEAM > 1) ? 0 : warp];   // line records: per warp, or shared by the team
    EvalScratch<NC>& wr = scratch[warp];                    // reduction tile: always per warp
    const unsigned unit0 = (TEAM > 1) ? blockIdx.x : blockIdx.x * EVAL_WARPS + warp;
    const unsigned nunits = (TEAM > 1) ? gridDim.x : gridDim.x * EVAL_WARPS;
    const int nhf = pr.nhf;
    const int nchunks = (pr.nchan + 31) >> 5;
    const int nhalf = (nchunks > 32) ? 2 : 1;
    const float ic0 = (float)pr.ic0;

    // one lane starts the two bulk copies of a pixel into stage ``b``; they complete on that stage's barrier
    auto issue = [&](unsigned px, int b) {
        const long long row = row_index ? row_index[px] : (long long)px;
        const unsigned bar = smem_u32(&bars[unit][b]);
        const unsigned dst = smem_u32(ubase + (size_t)b * buf_bytes);
        mbar_expect_tx(bar, buf_bytes);
        bulk_g2s(dst, spectra + row * stride, row_bytes, bar);
        bulk_g2s(dst + row_bytes, &states[px].in, (unsigned)sizeof(EvalIn<NC>), bar);
    };
    // pixel indices run two iterations ahead of the evaluation (their loads never sit on the critical path)
    unsigned px_cur = (unit0 < count) ? list[unit0] : 0u;
    unsigned px_nxt = (unit0 + nunits < count) ? list[unit0 + nunits] : 0u;
    if (leader && unit0 < count) issue(px_cur, 0);

    int k = 0;
    for (unsigned idx = unit0; idx < count; idx += nunits, ++k) {
        const int sb = k & 1;
        const unsigned px = px_cur;
        if (leader && idx + nunits < count) issue(px_nxt, sb ^ 1);   // stage sb^1 was released by the sync below
        const unsigned px_far = (idx + 2 * nunits < count) ? list[idx + 2 * nunits] : 0u;
        PixState<P>& ps = states[px];
        mbar_wait(smem_u32(&bars[unit][sb]), (unsigned)((k >> 1) & 1));
        const float* __restrict__ srow = reinterpret_cast<const float*>(ubase + (size_t)sb * buf_bytes);
        const EvalIn<NC>& in = *reinterpret_cast<const EvalIn<NC>*>(ubase + (size_t)sb * buf_bytes + row_bytes);
        // 1. fp64 line set-up, one line per lane (a team: one component per warp)
        if (lane < nhf) {
            const double c1 = tb.c1mrho[lane], rho = tb.rho[lane], kdv = tb.kdv[lane];
            const float lw = tb.lw[lane];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (TEAM > 1 && c != warp) continue;
                LineCoef lc;
                float lo, hi;
                line_coef(c1, rho, kdv, lw, pr.v0, pr.inv_dv, in.v[c], in.inv_sigma[c], lc, lo, hi);
                w.lines[c][lane] = lc;
                w.wlo[c][lane] = lo;
                w.whi[c][lane] = hi;
            }
        }
        CompCoef cc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) cc[c] = in.cc[c];
        if (TEAM > 1)
            __syncthreads();
        else
            __syncwarp();
        // 2. which line groups reach chunk `lane` (and chunk `lane + 32`): bit 8c + g
        unsigned packed[2] = {0u, 0u};
        for (int h = 0; h < nhalf; ++h) {
            const int ch = lane + 32 * h;
            const float lo_edge = 32.f * ch, hi_edge = 32.f * ch + 31.f;
            unsigned pk = 0;
            if (ch < nchunks) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        const bool on = (w.whi[c][LG::start(g + 1) - 1] >= lo_edge) && (w.wlo[c][LG::start(g)] <= hi_edge);
                        if (on) pk |= 1u << (8 * c + g);
                    }
                }
            }
            packed[h] = pk;
        }

        // accumulators, packed along the column index: h2[u][k] = {sum J_u J_2k, sum J_u J_2k+1} for 2k+1 >= u
        // (odd rows carry one unused lane), b2[k] = {sum J_2k r, sum J_2k+1 r}
        constexpr int PH = P / 2;
        f32x2 h2[P][PH], b2[PH];
        float chi = 0.f;
#pragma unroll
        for (int u = 0; u < P; ++u)
#pragma unroll
            for (int kk = 0; kk < PH; ++kk) h2[u][kk] = 0ull;
#pragma unroll
        for (int kk = 0; kk < PH; ++kk) b2[kk] = 0ull;
        unsigned n_chunks = 0, n_lines = 0;   // roofline accounting: 32-channel chunks / hyperfine lines x chunks evaluated

        int ordinal = 0;   // running index of the active chunks of this pixel (a team deals them round-robin)
#pragma unroll 1
        for (int h = 0; h < nhalf; ++h) {
            const unsigned mypk = h ? packed[1] : packed[0];
            unsigned act = __ballot_sync(FULL, mypk != 0u);
            if (TEAM > 1) {
                unsigned mine = 0u;
                for (unsigned a = act; a; a &= a - 1u)
                    if ((ordinal++ & (TEAM - 1)) == warp) mine |= a & (0u - a);
                act = mine;
            }
            // The active chunks are taken TWO at a time: both channels of a lane run through the same line records
            // (one shared-memory load per line and pair), the same group tests (on the OR of the two chunks' masks;
            // a group evaluated for a chunk just outside its window adds terms below 2^-28 of the line's peak) and two
            // independent radiative-transfer / accumulation chains -- twice the instruction-level parallelism per
            // warp at 4 warps per scheduler, where the single-chunk loop waited on its own dependent chain (ncu r2c:
            // issue slots 57 % busy, top stalls fixed-latency "wait" 1.85 and short_scoreboard 0.73 cycles per issue).
            // A pixel's pairs depend on its own chunk list only.  Channel values and the pair's mask are fetched one
            // pair ahead (shared memory / redux.sync).
            int ja = 0, jb = 0;
            bool has_b = false;
            auto next_pair = [&]() {
                ja = __ffs(act) - 1;
                act &= act - 1;
                has_b = act != 0u;
                jb = has_b ? (__ffs(act) - 1) : ja;
                act &= act - 1;    // (0 & anything) == 0 when there was no second chunk
            };
            float yna = 0.f, ynb = 0.f;
            unsigned pkn = 0u;
            int jna = 0, jnb = 0;
            bool hnb = false;
            bool more = act != 0u;
            if (more) {
                next_pair();
                jna = ja, jnb = jb, hnb = has_b;
                yna = srow[((jna + 32 * h) << 5) + lane];
                ynb = hnb ? srow[((jnb + 32 * h) << 5) + lane] : __int_as_float(0x7fc00000);
                pkn = __reduce_or_sync(FULL, (lane == jna || lane == jnb) ? mypk : 0u);
            }
            while (more) {
                const int j0 = jna, j1 = jnb;
                const bool two = hnb;
                const float y[2] = {yna, ynb};
                const unsigned pk = pkn;        // warp-uniform (redux result)
                more = act != 0u;
                if (more) {
                    next_pair();
                    jna = ja, jnb = jb, hnb = has_b;
                    yna = srow[((jna + 32 * h) << 5) + lane];
                    ynb = hnb ? srow[((jnb + 32 * h) << 5) + lane] : __int_as_float(0x7fc00000);
                    pkn = __reduce_or_sync(FULL, (lane == jna || lane == jnb) ? mypk : 0u);
                }
                const float fi[2] = {(float)(((j0 + 32 * h) << 5) + lane), (float)(((j1 + 32 * h) << 5) + lane)};
                float Gs[2][NC], As[2][NC], Bs[2][NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float g0 = 0.f, a0 = 0.f, b0 = 0.f, g1 = 0.f, a1 = 0.f, b1 = 0.f;
#pragma unroll
                    for (int gi = 0; gi < NG; ++gi) {
                        if (pk & (1u << (8 * c + gi))) {
#pragma unroll
                            for (int kk = LG::start(gi); kk < LG::start(gi + 1); ++kk) {
                                const float4 q = *reinterpret_cast<const float4*>(&w.lines[c][kk]);
                                LineCoef lc;
                                lc.nf = q.x;
                                lc.b = q.y;
                                lc.a = q.z;
                                lc.lw = q.w;
                                gauss_accum(fi[0], lc, g0, a0, b0);
                                gauss_accum(fi[1], lc, g1, a1, b1);
                            }
                        }
                    }
                    Gs[0][c] = g0, As[0][c] = a0, Bs[0][c] = b0;
                    Gs[1][c] = g1, As[1][c] = a1, Bs[1][c] = b1;
                }
                n_chunks += two ? 2u : 1u;
                n_lines += (two ? 2u : 1u) * group_lines<MODEL, NC>(pk);
                float m[2], J[2][P];
                rt_channel<NC>(Gs[0], As[0], Bs[0], fi[0] - ic0, cc, m[0], J[0]);
                rt_channel<NC>(Gs[1], As[1], Bs[1], fi[1] - ic0, cc, m[1], J[1]);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (y[t] == y[t]) {   // NaN channels (the NaN padding, the filler of an odd last chunk): no weight
                        const float r = y[t] - m[t];
                        f32x2 Jp[PH];
#pragma unroll
                        for (int kk = 0; kk < PH; ++kk) Jp[kk] = pk2(J[t][2 * kk], J[t][2 * kk + 1]);
                        const f32x2 r2 = pk2(r, r);
#pragma unroll
                        for (int kk = 0; kk < PH; ++kk) b2[kk] = fma2(Jp[kk], r2, b2[kk]);
#pragma unroll
                        for (int u = 0; u < P; ++u) {
                            const f32x2 ju = pk2(J[t][u], J[t][u]);
#pragma unroll
                            for (int kk = u / 2; kk < PH; ++kk) h2[u][kk] = fma2(ju, Jp[kk], h2[u][kk]);
                        }
                        chi = fmaf(m[t], m[t] - 2.f * y[t], chi);
                    }
                }
            }
        }
        const double sumsq = in.sumsq;
        // canonical order of the reduction: packed upper triangle of H, b, chi2 term
        float acc[NACC];
        {
            int q = 0;
#pragma unroll
            for (int u = 0; u < P; ++u) {
#pragma unroll
                for (int v = u; v < P; ++v) {
                    float x0, x1;
                    upk2(h2[u][v / 2], x0, x1);
                    acc[q++] = (v & 1) ? x1 : x0;
                }
            }
#pragma unroll
            for (int u = 0; u < P; ++u) {
                float x0, x1;
                upk2(b2[u / 2], x0, x1);
                acc[NH + u] = (u & 1) ? x1 : x0;
            }
            acc[NACC - 1] = chi;
        }
        // 3. cross-lane reduction in fp64 through shared memory, RR accumulators per pass
#pragma unroll
        for (int base = 0; base < NACC; base += RR) {
#pragma unroll
            for (int e = 0; e < RR; ++e)
                if (base + e < NACC) wr.red[e][lane] = acc[base + e];
            __syncwarp();
            const int e = base + lane;
            if (lane < RR && e < NACC) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int m = 0; m < 32; m += 4) {
                    s0 += (double)wr.red[lane][m];
                    s1 += (double)wr.red[lane][m + 1];
                    s2 += (double)wr.red[lane][m + 2];
                    s3 += (double)wr.red[lane][m + 3];
                }
                const double sred = (s0 + s1) + (s2 + s3);
                if (TEAM > 1)
                    part[warp][e] = sred;
                else if (e < NH)
                    ps.st.Ht[e] = sred;
                else if (e < NH + P)
                    ps.st.bt[e - NH] = sred;
                else
                    ps.st.chi2t = sumsq + sred;
            }
            __syncwarp();
        }
        if (TEAM > 1) {
            // chunks (and the gaussians evaluated for them) were shared out over the warps of the team
            if (lane == 0) {
                part[warp][NACC] = (double)n_chunks;
                part[warp][NACC + 1] = (double)n_lines;
            }
            __syncthreads();
            if (warp == 0) {
                for (int e = lane; e < NACC; e += 32) {
                    double sred = part[0][e];
#pragma unroll
                    for (int t = 1; t < TEAM; ++t) sred += part[t][e];
                    if (e < NH)
                        ps.st.Ht[e] = sred;
                    else if (e < NH + P)
                        ps.st.bt[e - NH] = sred;
                    else
                        ps.st.chi2t = sumsq + sred;
                }
                if (lane == 0) {
                    double nc = 0.0, nl = 0.0;
#pragma unroll
                    for (int t = 0; t < TEAM; ++t) {
                        nc += part[t][NACC];
                        nl += part[t][NACC + 1];
                    }
                    atomicAdd(&ps.n_gauss, (unsigned)nl);      // (result unused: a fire-and-forget reduction)
                    atomicAdd(&ps.n_chunks, (unsigned)nc);
                }
            }
            __syncthreads();   // the stage and ``part`` are free for the next pixel
        } else {
            if (lane == 0) {
                atomicAdd(&ps.n_gauss, n_lines);
                atomicAdd(&ps.n_chunks, n_chunks);
            }
            __syncwarp();      // every lane is done with this stage: the next copy may overwrite it
        }
        px_cur = px_nxt;
        px_nxt = px_far;
    }
}

// ---- solver: P lanes per active pixel (b200fit_lanes.cuh) ------------------------------------------------------
// The state is stepped in registers: each lane loads its parameter's entries and its row of J^T J straight from
// the per-pixel state in global memory (L2-resident), and writes back only what changed.  History of this kernel
// (profiles/): thread per pixel on global state 30 ms per 2c fit, the same staged through shared memory 63 ms
// (copy loops serialise global round trips), both bound by the ~20k-instruction dependent chain of one serial step.
template <int NC>
__global__ void __launch_bounds__(SOLVE_THREADS, 8)   // 128 regs (a few spills) beat 168 / 202 and 96 / 80 / 64 (measured)
    fit_solve_kernel(const DeviceProblem pr, PixState<4 * NC>* __restrict__ states, const unsigned* __restrict__ list,
                     unsigned* __restrict__ next_list, FitQueues* __restrict__ fq, int cur,
                     double* __restrict__ params, double* __restrict__ perr, double* __restrict__ chi2,
                     double* __restrict__ err_used, int* __restrict__ status, unsigned* __restrict__ counts) {
    constexpr int P = 4 * NC;
    constexpr unsigned GROUPS = SOLVE_THREADS / P;   // pixels per block and pass
    __shared__ FitTables tb;
    // The grid is sized for a full list (it is part of a captured graph); in the late rounds most blocks have nothing
    // to do and leave before the table load and its barrier -- this kernel's latency is what the tail of a fit waits on.
    const unsigned count = fq->count[cur];
    if (blockIdx.x * GROUPS >= count) return;
    load_fit_tables(pr, tb);
    const LMConfig cf{tb.lo, tb.hi, pr.ftol, pr.max_iter};
    const LaneGroup<P> g;
    const unsigned gsub = threadIdx.x / P;   // group of this lane within the block
    // warp-uniform trip count: all groups of a warp loop together, the ones past the end carry ``valid = false``
    for (unsigned idx0 = blockIdx.x * GROUPS; idx0 < count; idx0 += gridDim.x * GROUPS) {
        const unsigned warp_first = idx0 + (gsub & ~((32u / P) - 1u));
        if (warp_first >= count) continue;   // the whole warp is past the end (uniform)
        const unsigned idx = idx0 + gsub;
        const bool valid = idx < count;
        const unsigned px = valid ? list[idx] : list[warp_first];
        PixState<P>& ps = states[px];
        LaneState<P> s;
        lane_load<P>(g, ps.st, cf, s);
        bool accepted = false;
        const int done = lane_step<P>(g, s, cf, valid, accepted);
        // coefficients of the next trial point: lanes 4c .. 4c + 3 of a group work on component c (the fp64 exp and
        // expm1 of the Planck terms are ~6 % of this kernel's instructions per component), lane 4c of a live group stores
        const int cmine = (NC > 1) ? (g.gl >> 2) : 0;
        double p4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) p4[k] = g.bcast(s.pt, 4 * cmine + k);
        if (!valid) continue;
        lane_store<P>(g, ps.st, s, accepted);   // results are written by fit_finalize_kernel
        if (!done) {
            CompCoef cc;
            comp_coefs(pr.T0c, pr.dT0, pr.Jbg0, pr.Jbg1, p4, cc);
            if ((g.gl & 3) == 0) {
                ps.in.cc[cmine] = cc;
                ps.in.v[cmine] = p4[0];
                ps.in.inv_sigma[cmine] = 1.0 / p4[1];
            }
            if (g.gl == 0) next_list[atomicAdd(&fq->count[cur ^ 1], 1u)] = px;
        }
    }
}

// ---- results: P lanes per pixel, once, after the last round -----------------------------------------------------
// Parameter errors sqrt(diag(err^2 (H_FF)^-1)) (0 for pegged parameters, mpfit calc_covar on the column-zeroed
// Jacobian) and the result rows.  Kept out of the solver kernel so that its code -- one more factorisation and
// eight more triangular solves -- does not sit in the instruction stream of every round.
template <int NC>
__global__ void __launch_bounds__(SOLVE_THREADS)
    fit_finalize_kernel(const DeviceProblem pr, const PixState<4 * NC>* __restrict__ states,
                        double* __restrict__ params, double* __restrict__ perr, double* __restrict__ chi2,
                        double* __restrict__ err_used, int* __restrict__ status, unsigned* __restrict__ counts) {
    constexpr int P = 4 * NC;
    constexpr unsigned GROUPS = SOLVE_THREADS / P;
    __shared__ FitTables tb;
    load_fit_tables(pr, tb);
    const LMConfig cf{tb.lo, tb.hi, pr.ftol, pr.max_iter};
    const LaneGroup<P> g;
    const long long npad = (pr.npix + (32 / P) - 1) / (32 / P) * (32 / P);   // whole warps loop together
    for (long long px_ = (long long)blockIdx.x * GROUPS + threadIdx.x / P; px_ < npad;
         px_ += (long long)gridDim.x * GROUPS) {
        const bool inside = px_ < pr.npix;
        const long long px = inside ? px_ : pr.npix - 1;
        const PixState<P>& ps = states[px];
        LaneState<P> s;
        lane_load<P>(g, ps.st, cf, s);
        const bool good = ps.fitted && s.status > 0;
        const double err = ps.in.err, err2 = err * err;
        const double pe_all = lane_error<P>(g, s, err2);   // warp-uniform collectives
        if (!ps.fitted || !inside) continue;   // zeros were written by the prologue
        const double pe = good ? pe_all : 0.0;
        params[(size_t)px * P + g.gl] = good ? s.p : 0.0;
        perr[(size_t)px * P + g.gl] = pe;
        if (g.gl == 0) {
            chi2[px] = good ? s.chi2 / err2 : 0.0;   // at the returned parameters (lm_reported_chi2)
            err_used[px] = err;
            status[px] = s.status;
            if (counts) {
                counts[4 * (size_t)px + 0] = (unsigned)s.nevals;
                counts[4 * (size_t)px + 1] = (unsigned)s.nrej;
                counts[4 * (size_t)px + 2] = ps.n_gauss;
                counts[4 * (size_t)px + 3] = ps.n_chunks;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// fp64 model in the reference's operation order (HyperfineModel.py:41-58, 91-107; BaseModel.py:191-192).
// Explicit _rn intrinsics keep nvcc from contracting mul+add into FMA, so the arithmetic is the numpy sequence,
// with two exceptions inside the hyperfine sum, which is ~80 % of this model's instructions: the gaussian argument
// is -(a*a) * (1 / (2 w^2)) with the reciprocal taken once per line (numpy divides: last-bit differences in a
// quarter of the terms) and its exponential is ``exp_neg`` below (CUDA's exp already differs from glibc's in the
// last bit).  Both are far inside the 1e-13 agreement the parity tests ask of the model, and the sample mask
// ``model > 0`` can only change for a channel whose optical depth sits within ~1e-12 (relative) of the 1.1e-16
// where 1 - exp(-tau) leaves zero -- rarer than the platform dependence of that channel documented in DESIGN.md.
// Hyperfine terms below exp(-50) = 2e-22 relative are skipped: where such a term is the largest one the optical
// depth is far below the 5.6e-17 at which 1 - exp(-tau) leaves zero (the model is exactly 0 there either way), and
// next to larger terms it is below their rounding error.
// Quantities that depend on the channel only (frequency, h nu / k, J(T_CMB)) come from a per-context table
// (chan_tables_kernel) instead of two to five divisions and an exponential per pixel and channel; same values.
constexpr int MASK_WORDS = B200FIT_MAX_NCHAN / 32;   // spectral masks are bit sets of at most 64 words
static_assert(MASK_WORDS == 64, "comp_setup_f64 fills two words per lane");

struct CompF64 {
    double tex, tau, nuoff[MAXHF_FULL], rcp2w2[MAXHF_FULL], tw[MAXHF_FULL];   // rcp2w2 = 1 / (2 nuwidth^2)
    double acut[MAXHF_FULL];   // |x + nuoff - line| beyond which the term is below exp(-50): skipped before the multiply
    unsigned char gbits[MASK_WORDS];        // per 32-channel word: bit g set when line group g (< 8) reaches into it
};

// Warp-cooperative set-up of one component: lane k prepares line k (the divisions and the square root are the
// expensive part), then lane g derives the channel range of line group g.  Call with all 32 lanes.
__device__ __forceinline__ void comp_setup_f64(const DeviceProblem& pr, const double* p4, CompF64& c, int lane) {
    const double vel = p4[0], width = p4[1], tau = p4[3];
    if (lane == 0) {
        c.tex = p4[2];
        c.tau = tau;
    }
    if (lane < pr.nfull) {
        const int k = lane;
        const double line = pr.lines_full[k];
        const double nuwidth = fabs(__dmul_rn(__ddiv_rn(width, CKMS), line));
        c.nuoff[k] = __dmul_rn(__ddiv_rn(vel, CKMS), line);
        const double d = __dmul_rn(2.0, __dmul_rn(nuwidth, nuwidth));   // 2.0 * nuwidth**2 (the divisor)
        c.rcp2w2[k] = __ddiv_rn(1.0, d);
        c.tw[k] = __dmul_rn(tau, pr.wts_full[k]);
        c.acut[k] = sqrt(50.0 * d) * (1.0 + 1e-9);   // generous: the exact arg > -50 test still follows
    }
    __syncwarp();
    // channel range [clo, chi] of every line group (x decreases with the channel index; +2 channels of margin):
    // channels outside it are beyond every acut of the group
    int clo = 1, chi = 0;
    if (lane < pr.ngroups_full) {
        const int g = lane;
        double xmin = INFINITY, xmax = -INFINITY;
        for (int k = pr.gstart_full[g]; k < pr.gstart_full[g + 1]; ++k) {
            const double ctr = pr.lines_full[k] - c.nuoff[k];
            xmin = fmin(xmin, ctr - c.acut[k]);
            xmax = fmax(xmax, ctr + c.acut[k]);
        }
        const double ilo = ((1.0 - xmax / pr.nu_ref_ghz) * CKMS - pr.v0) / pr.dv;
        const double ihi = ((1.0 - xmin / pr.nu_ref_ghz) * CKMS - pr.v0) / pr.dv;
        clo = (int)fmax(fmin(floor(ilo) - 2.0, 1e9), -1e9);
        chi = (int)fmax(fmin(ceil(ihi) + 2.0, 1e9), -1e9);
    }
    // which groups can contribute to each 32-channel word: the hyperfine sum then walks only those (a word outside
    // every window costs one byte load instead of a test per group, channel and component)
    unsigned m0 = 0u, m1 = 0u;   // words lane and lane + 32
    for (int g = 0; g < pr.ngroups_full; ++g) {
        const int lo = __shfl_sync(FULL, clo, g), hi = __shfl_sync(FULL, chi, g);
        if (!((lane << 5) + 31 < lo || (lane << 5) > hi)) m0 |= 1u << g;
        if (!(((lane + 32) << 5) + 31 < lo || ((lane + 32) << 5) > hi)) m1 |= 1u << g;
    }
    c.gbits[lane] = (unsigned char)m0;
    c.gbits[lane + 32] = (unsigned char)m1;
    __syncwarp();
}

__device__ __forceinline__ double xarr_ghz(const DeviceProblem& pr, int i) {
    const double v = __dadd_rn(pr.v0, __dmul_rn(pr.dv, (double)i));
    return __ddiv_rn(__dmul_rn(__dmul_rn(pr.nu_ref_ghz, 1e9), __dsub_rn(1.0, __ddiv_rn(v, CKMS))), 1e9);
}

// exp(x) for -50 <= x <= 0 (valid down to ~ -700), ~1 ulp: x = n ln2 + r, |r| <= ln2 / 2, degree-13 Taylor,
// exponent added in place.
// Half the instructions of exp(): no overflow / subnormal handling is needed on this range.
__device__ __forceinline__ double exp_neg(double x) {
    const double n = rint(x * 1.4426950408889634);
    double r = fma(n, -6.93147180369123816490e-01, x);
    r = fma(n, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;                 // 1 / 13!  (Horner; an Estrin tree, 5 deep instead of 13,
    p = fma(p, r, 2.08767569878681e-09);               //  measured the same: the term loop's branches, not this
    p = fma(p, r, 2.505210838544172e-08);              //  chain, are what the warps wait on)
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.388888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + ((int)n << 20), __double2loint(p));
}

// exp(x) as numpy evaluates it.  glibc's exp is within ~0.51 ulp, i.e. practically correctly rounded; CUDA's is
// within 1 ulp.  The difference matters in exactly one place: exp(-tau) for tau ~ 1e-16 decides whether
// 1 - exp(-tau) is 0 or 1.1e-16, i.e. whether a channel at the edge of a hyperfine window belongs to the AICc sample
// mask ``model > 0``.  For tiny |x| the correctly rounded result is fma(x, 1 + x/2, 1).
__device__ __forceinline__ double exp_np(double x) {
    return (fabs(x) < 9.5367431640625e-07) ? fma(x, fma(x, 0.5, 1.0), 1.0) : exp(x);
}

__device__ __forceinline__ double t_antenna(double T0, double T) {
    return __ddiv_rn(T0, __dsub_rn(exp(__ddiv_rn(T0, T)), 1.0));
}

// optical depth profile of one component at frequency x (HyperfineModel.py:99-103); terms below exp(-50) of their
// peak are skipped (see the note above CompF64)
__device__ __forceinline__ double tauprof_f64(const DeviceProblem& pr, const CompF64& cp, double x, int i) {
    double tauprof = 0.0;
    for (unsigned m = cp.gbits[i >> 5]; m; m &= m - 1u) {   // groups in reach of this word, ascending as in the sum
        const int g = __ffs(m) - 1;
        for (int k = pr.gstart_full[g]; k < pr.gstart_full[g + 1]; ++k) {
            const double a = __dsub_rn(__dadd_rn(x, cp.nuoff[k]), pr.lines_full[k]);
            if (fabs(a) > cp.acut[k]) continue;   // (a branch-free form that weights skipped terms by 0 was 65 % slower)
            const double arg = __dmul_rn(-__dmul_rn(a, a), cp.rcp2w2[k]);
            if (arg > -50.0) tauprof = __dadd_rn(tauprof, __dmul_rn(cp.tw[k], exp_neg(arg)));
        }
    }
    return tauprof;
}

// One radiative-transfer slab in the reference's operation order.  tau == 0 leaves the background untouched
// bit for bit (jt * (1 - 1) + bg * 1 == bg), so the Planck terms are not evaluated for it.
__device__ __forceinline__ double slab_f64(double bg, double T0, double tex, double tauprof) {
    if (tauprof == 0.0) return bg;
    const double e = exp_np(-tauprof);
    const double jt = t_antenna(T0, tex);
    return __dadd_rn(__dmul_rn(jt, __dsub_rn(1.0, e)), __dmul_rn(bg, e));
}

__device__ __forceinline__ double planck_T0(double x) {
    return __ddiv_rn(__dmul_rn(__dmul_rn(H_CGS, x), 1e9), KB_CGS);
}

// per-channel constants of the fp64 model: tab[i] = frequency (GHz), tab[MAX + i] = h nu / k, tab[2 MAX + i] =
// J(T_CMB, nu) -- the same operation sequences model_f64 used to run per pixel
__global__ void chan_tables_kernel(const DeviceProblem pr, double* __restrict__ tab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pr.nchan) return;
    const double x = xarr_ghz(pr, i);
    const double T0 = planck_T0(x);
    tab[i] = x;
    tab[B200FIT_MAX_NCHAN + i] = T0;
    tab[2 * B200FIT_MAX_NCHAN + i] = t_antenna(T0, TCMB);
}

// model of ``ncomp`` (<= 2) components at channel i; comps in shared/local memory.  A channel outside every
// hyperfine window is exactly 0 (bg0 - bg0) and costs only the window tests.
__device__ __forceinline__ double model_f64(const DeviceProblem& pr, const double* __restrict__ tab,
                                            const CompF64* comps, int ncomp, int i) {
    const double x = __ldg(tab + i);
    const double t0 = tauprof_f64(pr, comps[0], x, i);
    const double t1 = (ncomp > 1) ? tauprof_f64(pr, comps[1], x, i) : 0.0;
    if (t0 == 0.0 && t1 == 0.0) return 0.0;
    const double T0 = __ldg(tab + B200FIT_MAX_NCHAN + i);
    const double bg0 = __ldg(tab + 2 * B200FIT_MAX_NCHAN + i);
    double bg = slab_f64(bg0, T0, comps[0].tex, t0);
    if (ncomp > 1) bg = slab_f64(bg, T0, comps[1].tex, t1);
    return __dsub_rn(bg, bg0);
}

// model A (comps[0 .. nA)) and model B (comps[2 .. 2 + nB)) of one pixel at channel i, sharing the frequency axis and
// the Planck terms.  A = the 1-component fit, B = the 2-component fit in get_all_lnk_maps; A = a new fit, B = the
// old fit with the same number of components in replace_bad_pix (master_fitter.py:1449).
__device__ __forceinline__ void model_pair_f64(const DeviceProblem& pr, const double* __restrict__ tab,
                                               const CompF64* comps, bool vA, bool vB, int nA, int nB, int i,
                                               double& mA, double& mB) {
    const double x = __ldg(tab + i);
    const double a0 = vA ? tauprof_f64(pr, comps[0], x, i) : 0.0;
    const double a1 = (vA && nA > 1) ? tauprof_f64(pr, comps[1], x, i) : 0.0;
    const double b0 = vB ? tauprof_f64(pr, comps[2], x, i) : 0.0;
    const double b1 = (vB && nB > 1) ? tauprof_f64(pr, comps[3], x, i) : 0.0;
    mA = 0.0;
    mB = 0.0;
    if (a0 == 0.0 && a1 == 0.0 && b0 == 0.0 && b1 == 0.0) return;
    const double T0 = __ldg(tab + B200FIT_MAX_NCHAN + i);
    const double bg0 = __ldg(tab + 2 * B200FIT_MAX_NCHAN + i);
    if (a0 != 0.0 || a1 != 0.0)
        mA = __dsub_rn(slab_f64(slab_f64(bg0, T0, comps[0].tex, a0), T0, comps[1].tex, a1), bg0);
    if (b0 != 0.0 || b1 != 0.0)
        mB = __dsub_rn(slab_f64(slab_f64(bg0, T0, comps[2].tex, b0), T0, comps[3].tex, b1), bg0);
}

__device__ __forceinline__ bool params_valid(const double* p, int P) {
    bool any = false, fin = true;
    for (int j = 0; j < P; ++j) {
        any = any || (p[j] != 0.0);
        fin = fin && isfinite(p[j]);
    }
    return any && fin;
}

constexpr int AUX_WARPS = 4;

struct AuxShared {
    CompF64 comps[4];   // model A in [0, nA), model B in [2, 2 + nB)   (nA, nB in {1, 2})
};

// get_modelcube: one warp per pixel
__global__ void __launch_bounds__(AUX_WARPS * 32)
    model_kernel(const DeviceProblem pr, const double* __restrict__ tab, const double* __restrict__ params,
                 double* __restrict__ model, long long stride) {
    __shared__ AuxShared sh[AUX_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int P = 4 * pr.ncomp;
    for (long long pix = (long long)blockIdx.x * AUX_WARPS + warp; pix < pr.npix;
         pix += (long long)gridDim.x * AUX_WARPS) {
        const double* p = params + pix * P;
        const bool valid = params_valid(p, P);
        if (valid)
            for (int c = 0; c < pr.ncomp; ++c) comp_setup_f64(pr, p + 4 * c, sh[warp].comps[c], lane);
        for (int i = lane; i < (int)stride; i += 32) {
            double m = NAN;
            if (valid && i < pr.nchan) m = model_f64(pr, tab, sh[warp].comps, pr.ncomp, i);
            model[pix * stride + i] = m;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Spectral masks are bit sets: bits[w] bit b <-> channel 32*w + b, at most 64 words (B200FIT_MAX_NCHAN = 2048).

// Dilation along the spectral axis: True at k spreads to [k - e/2, k + (e-1)/2] (scipy binary_dilation with
// ones(e), origin 0 -- UltraCube.expand_mask :1658-1698; e = 20: [k-10, k+9]).  Word-parallel, e <= 64.
__device__ __forceinline__ unsigned dilate_word(const unsigned* bits, int nwords, int w, int expand) {
    const unsigned cur = bits[w];
    if (expand <= 1) return cur;
    const int up = (expand - 1) / 2, dn = expand / 2;
    const unsigned prev = (w > 0) ? bits[w - 1] : 0u, next = (w + 1 < nwords) ? bits[w + 1] : 0u;
    const unsigned long long lo = ((unsigned long long)cur << 32) | prev;    // left shifts pull bits out of prev
    const unsigned long long hi = ((unsigned long long)next << 32) | cur;    // right shifts pull bits out of next
    unsigned out = cur;
    for (int s = 1; s <= up; ++s) out |= (unsigned)((lo << s) >> 32);
    for (int s = 1; s <= dn; ++s) out |= (unsigned)(hi >> s);
    return out;
}

// In-place dilation of a warp's bit set (smem), clipped to [0, nchan).  Warp-collective.
__device__ __forceinline__ void dilate_bits(unsigned* bits, int nchan, int expand, int lane) {
    const int nwords = (nchan + 31) >> 5;
    unsigned d0 = 0, d1 = 0;
    if (lane < nwords) d0 = dilate_word(bits, nwords, lane, expand);
    if (lane + 32 < nwords) d1 = dilate_word(bits, nwords, lane + 32, expand);
    __syncwarp();
    const int tail = nchan & 31;
    if (lane < nwords) bits[lane] = (lane == nwords - 1 && tail) ? (d0 & ((1u << tail) - 1u)) : d0;
    if (lane + 32 < nwords) bits[lane + 32] = (lane + 32 == nwords - 1 && tail) ? (d1 & ((1u << tail) - 1u)) : d1;
    __syncwarp();
}

__device__ __forceinline__ void setup_pixel_models(const DeviceProblem& pr, AuxShared& sh, const double* pA,
                                                   const double* pB, bool vA, bool vB, int nA, int nB, int lane) {
    if (vA)
        for (int c = 0; c < nA; ++c) comp_setup_f64(pr, pA + 4 * c, sh.comps[c], lane);
    if (vB)
        for (int c = 0; c < nB; ++c) comp_setup_f64(pr, pB + 4 * c, sh.comps[2 + c], lane);
}

// Un-dilated union mask (model1 > 0) | (model2 > 0) of ONE pixel, written as a bit set: the spectral mask the
// include_nosamp rule copies into pixels without samples (UltraCube.py:1423-1438).
__global__ void __launch_bounds__(32)
    mask_bits_kernel(const DeviceProblem pr, const double* __restrict__ tab, int nA, int nB, const double* __restrict__ params1,
                     const double* __restrict__ params2, int model_f32, const long long* __restrict__ best /* {count, pix} or null */, long long pix_arg,
                     unsigned* __restrict__ bits_out) {
    __shared__ AuxShared sh;
    const int lane = threadIdx.x;
    long long pix = pix_arg;
    if (best) pix = (best[0] > 0) ? best[1] : -1;   // nothing has a sample anywhere: the fill mask is empty
    const int nwords = (pr.nchan + 31) >> 5;
    if (pix < 0 || pix >= pr.npix) {
        for (int wd = lane; wd < MASK_WORDS; wd += 32) bits_out[wd] = 0u;
        return;
    }
    const bool v1 = params1 && params_valid(params1 + pix * 4 * nA, 4 * nA);
    const bool v2 = params2 && params_valid(params2 + pix * 4 * nB, 4 * nB);
    setup_pixel_models(pr, sh, params1 + pix * 4 * nA, params2 + pix * 4 * nB, v1, v2, nA, nB, lane);
    for (int wd = 0; wd < MASK_WORDS; ++wd) {
        const int i = (wd << 5) + lane;
        bool on = false;
        if (wd < nwords && i < pr.nchan) {
            double m1, m2;
            model_pair_f64(pr, tab, sh.comps, v1, v2, nA, nB, i, m1, m2);
            on = model_f32 ? (((float)m1 > 0.f) || ((float)m2 > 0.f)) : ((m1 > 0.0) || (m2 > 0.0));
        }
        const unsigned b = __ballot_sync(FULL, on);
        if (lane == 0) bits_out[wd] = b;
    }
}

// First (lowest index) pixel with the largest mask count: key = count << 40 | (2^40 - 1 - index), atomicMax.
__global__ void __launch_bounds__(256)
    mask_argmax_kernel(const int* __restrict__ counts, long long npix, long long index_offset,
                       unsigned long long* __restrict__ key_out) {
    unsigned long long best = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k =
            ((unsigned long long)(unsigned)counts[i] << 40) | (0xFFFFFFFFFFull - (unsigned long long)(i + index_offset));
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(FULL, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(key_out, best);
}
__global__ void mask_argmax_finish_kernel(const unsigned long long* __restrict__ key, long long* __restrict__ best) {
    best[0] = (long long)(key[0] >> 40);
    best[1] = (long long)(0xFFFFFFFFFFull - (key[0] & 0xFFFFFFFFFFull));
}

// get_all_lnk_maps + the get_best_2c_parcube selection, one warp per pixel.
//   pass 1 (only_empty = 0): every pixel; pixels whose own mask is empty use fill_bits when given, else they get
//                            NaN statistics; mask_counts[pix] (if given) receives the un-dilated mask size.
//   pass 2 (only_empty = 1): only the pixels with mask_counts[pix] == 0, with fill_bits (the global include_nosamp
//                            mask, known only after pass 1 has seen every pixel of every rank).
__global__ void __launch_bounds__(AUX_WARPS * 32)
    aicc_kernel(const DeviceProblem pr, const double* __restrict__ tab, int nA, int nB,
                const float* __restrict__ spectra, long long stride,
                const long long* __restrict__ row_index, const double* __restrict__ params1,
                const double* __restrict__ params2, const unsigned* __restrict__ fill_bits, int only_empty, int expand,
                int model_f32, double thr10, double thr20, double thr21, double* __restrict__ rss_out,
                double* __restrict__ nsamp_out, double* __restrict__ lnk_out, signed char* __restrict__ ncomp_out,
                int* __restrict__ mask_counts, const unsigned* __restrict__ ext_bits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ AuxShared sh[AUX_WARPS];
    __shared__ unsigned bits[AUX_WARPS][MASK_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* m1s = reinterpret_cast<float*>(smem_raw) + (size_t)warp * 2 * stride;
    float* m2s = m1s + stride;
    double* m1d = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * stride;   // fp64 variant
    double* m2d = m1d + stride;
    const int nwords = (pr.nchan + 31) >> 5;
    for (long long pix = (long long)blockIdx.x * AUX_WARPS + warp; pix < pr.npix;
         pix += (long long)gridDim.x * AUX_WARPS) {
        if (only_empty && mask_counts[pix] != 0) continue;
        const bool v1 = params_valid(params1 + pix * 4 * nA, 4 * nA);
        const bool v2 = params_valid(params2 + pix * 4 * nB, 4 * nB);
        setup_pixel_models(pr, sh[warp], params1 + pix * 4 * nA, params2 + pix * 4 * nB, v1, v2, nA, nB, lane);
        // models (kept in smem) + union mask
        for (int wd = 0; wd < nwords; ++wd) {
            const int i = (wd << 5) + lane;
            bool on = false;
            if (i < pr.nchan) {
                double m1, m2;
                model_pair_f64(pr, tab, sh[warp].comps, v1, v2, nA, nB, i, m1, m2);
                if (model_f32) {
                    m1s[i] = (float)m1;
                    m2s[i] = (float)m2;
                    on = ((float)m1 > 0.f) || ((float)m2 > 0.f);
                } else {
                    m1d[i] = m1;
                    m2d[i] = m2;
                    on = (m1 > 0.0) || (m2 > 0.0);
                }
            }
            const unsigned b = __ballot_sync(FULL, on);
            if (lane == 0) bits[warp][wd] = b;
        }
        __syncwarp();
        if (ext_bits) {   // the caller's spectral mask instead of (model A > 0) | (model B > 0)  (get_rss(mask=...))
            for (int wd = lane; wd < nwords; wd += 32) bits[warp][wd] = ext_bits[pix * MASK_WORDS + wd];
            __syncwarp();
        }
        int cnt = 0;
        for (int wd = lane; wd < nwords; wd += 32) cnt += __popc(bits[warp][wd]);
        cnt = warp_sum_i(cnt);
        if (!only_empty && mask_counts && lane == 0) mask_counts[pix] = cnt;
        bool have_mask = cnt > 0;
        if (!have_mask && fill_bits) {   // include_nosamp: borrow the spectral mask of the max-N pixel
            for (int wd = lane; wd < nwords; wd += 32) bits[warp][wd] = fill_bits[wd];
            have_mask = true;
        }
        __syncwarp();
        if (have_mask) dilate_bits(bits[warp], pr.nchan, expand, lane);
        // rss over the dilated mask.  residual = data - model in the cube's dtype (fp32) when model_f32,
        // squared and summed in fp64 (numpy: (residual * mask.astype(float)) ** 2 -> float64).
        double r0 = 0.0, r1 = 0.0, r2 = 0.0;
        int ns = 0;
        const long long row = row_index ? row_index[pix] : pix;
        const float* y = spectra + row * stride;
        if (have_mask) {
            for (int wd = 0; wd < nwords; ++wd) {
                const unsigned bw = bits[warp][wd];
                if (!bw) continue;
                if (!((bw >> lane) & 1u)) continue;
                const int i = (wd << 5) + lane;
                ns += 1;
                const float d = y[i];
                if (!(d == d)) continue;   // nansum skips NaN data, but the channel still counts in N
                double e1, e2;
                const double e0 = (double)d;
                if (model_f32) {
                    e1 = (double)(d - m1s[i]);
                    e2 = (double)(d - m2s[i]);
                } else {
                    e1 = (double)d - m1d[i];
                    e2 = (double)d - m2d[i];
                }
                r0 += e0 * e0;
                r1 += e1 * e1;
                r2 += e2 * e2;
            }
        }
        r0 = warp_sum(r0);
        r1 = warp_sum(r1);
        r2 = warp_sum(r2);
        ns = warp_sum_i(ns);
        if (lane == 0) {
            // get_rss: rss == 0 -> NaN, N = NaN where rss NaN; method: N < 1 -> NaN, rss <= 0 -> NaN
            const double N = (double)ns;
            double rs[3] = {r0, r1, r2}, aic[3];
            for (int k = 0; k < 3; ++k) {
                double rr = rs[k];
                double nn = N;
                if (rr == 0.0) {
                    rr = NAN;
                    nn = NAN;
                }
                if (nn < 1.0) {
                    nn = NAN;
                    rr = NAN;
                }
                if (rr <= 0.0) rr = NAN;
                rs[k] = rr;
                const double p = (k == 0) ? 0.0 : ((k == 1) ? 4.0 * nA : 4.0 * nB);
                // aic.AIC / AICc (aic.py:136-181)
                aic[k] = nn * log(rr / nn) + 2.0 * p + (2.0 * p * (p + 1.0)) / (nn - p - 1.0);
                rss_out[pix * 3 + k] = rr;
            }
            nsamp_out[pix] = (N < 1.0) ? NAN : N;   // per-model NaNs follow from rss_out
            double l10 = -1.0 * (aic[1] - aic[0]) / 2.0;
            double l20 = -1.0 * (aic[2] - aic[0]) / 2.0;
            double l21 = -1.0 * (aic[2] - aic[1]) / 2.0;
            signed char nc = ((l21 > thr21) && (l20 > thr20)) ? (signed char)nB : (signed char)nA;
            if (!(l10 > thr10)) nc = 0;
            // lnk maps are NaN where the corresponding fit does not exist (UltraCube.py:1369-1377)
            if (!v1) l10 = NAN;
            if (!v2) {
                l20 = NAN;
                l21 = NAN;
            }
            lnk_out[pix * 3 + 0] = l10;
            lnk_out[pix * 3 + 1] = l20;
            lnk_out[pix * 3 + 2] = l21;
            ncomp_out[pix] = nc;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// What cubefit_simp does to the caller's guesses before fiteach (multi_v_fit.py:156-190), per pixel, on the device:
// a velocity guess of exactly 0 means "no guess" (-> NaN, :158-159); every parameter outside its limits is moved
// just inside (value < min -> min + eps, value > max -> max - eps, :178-190; NaN stays NaN).  Input field-major
// raw[P][npix] (the [P, ny, nx] maps as the caller holds them), output pixel-major out[npix][P] (what b200fit_fit
// reads).  A pixel with a non-finite guess is "not fitted" by the fit kernel itself (status 0, zeros) -- the
// reference's maskmap & all-finite rule -- so no compaction is needed.
__global__ void __launch_bounds__(256)
    prepare_guesses_kernel(long long npix, int P, const double* __restrict__ raw, DeviceLimits lim, double eps,
                           double* __restrict__ out) {
    for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < npix;
         px += (long long)gridDim.x * blockDim.x) {
        for (int j = 0; j < P; ++j) {
            double v = raw[(long long)j * npix + px];
            if ((j & 3) == 0 && v == 0.0) v = NAN;
            if (v < lim.lo[j]) v = lim.lo[j] + eps;
            if (v > lim.hi[j]) v = lim.hi[j] - eps;
            out[px * P + j] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// get_best_2c_parcube (UltraCube.py:1283-1379, include_1c, return_lnks) assembled on the device, field-major:
//   out[ 0.. 3] parcube1   out[ 4.. 7] errcube1   out[ 8..15] parcube2   out[16..23] errcube2
//   out[24..31] best parcube: the 2-component fit where ncomp == 2; the 1-component fit in planes 0-3 and NaN in
//               4-7 where ncomp == 1; NaN where ncomp == 0            out[32..39] best errcube, likewise
//   out[40..42] lnk10, lnk20, lnk21 (NaN where the fit does not exist)   out[43] ncomp
// One thread per pixel; every plane is written coalesced.  This is what a partitioned run gathers: [44][npix] rows.
constexpr int PACK_FIELDS = B200FIT_PACK_FIELDS;
__global__ void __launch_bounds__(256)
    pack_selection_kernel(long long npix, const double* __restrict__ params1, const double* __restrict__ perr1,
                          const double* __restrict__ params2, const double* __restrict__ perr2,
                          const double* __restrict__ lnk, const signed char* __restrict__ ncomp,
                          double* __restrict__ out, long long out_stride) {
    for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < npix;
         px += (long long)gridDim.x * blockDim.x) {
        const int nc = ncomp[px];
        double p1[4], e1[4], p2[8], e2[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            p1[j] = params1[px * 4 + j];
            e1[j] = perr1[px * 4 + j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            p2[j] = params2[px * 8 + j];
            e2[j] = perr2[px * 8 + j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            out[(0 + j) * out_stride + px] = p1[j];
            out[(4 + j) * out_stride + px] = e1[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            out[(8 + j) * out_stride + px] = p2[j];
            out[(16 + j) * out_stride + px] = e2[j];
            double bp = NAN, be = NAN;
            if (nc == 2) {
                bp = p2[j];
                be = e2[j];
            } else if (nc == 1 && j < 4) {
                bp = p1[j];
                be = e1[j];
            }
            out[(24 + j) * out_stride + px] = bp;
            out[(32 + j) * out_stride + px] = be;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) out[(40 + k) * out_stride + px] = lnk[px * 3 + k];
        out[43 * out_stride + px] = (double)nc;
    }
}

// ---------------------------------------------------------------------------------------------------------
// get_rms + reduced get_chisq of one fitted model (UltraCube.py:1701-1740, :1478-1565), one warp per pixel:
//   r = data - model;  rms = 1.4826 * nanmedian(|r - roll(r, 2)|) / sqrt(2);
//   chisq_red = nansum((r * mask)^2) / sum(mask) / rms^2,  mask = dilate(model > 0, expand)
// The median is taken with a warp-wide bitonic sort of |diff| in shared memory.
__global__ void __launch_bounds__(AUX_WARPS * 32)
    resid_stats_kernel(const DeviceProblem pr, const double* __restrict__ tab, const float* __restrict__ spectra,
                       long long stride,
                       const long long* __restrict__ row_index, const double* __restrict__ params, int expand,
                       int model_f32, int npow2, double* __restrict__ rms_out, double* __restrict__ chisq_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ AuxShared sh[AUX_WARPS];
    __shared__ unsigned bits[AUX_WARPS][MASK_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* res = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * npow2;   // residuals
    double* key = res + npow2;                                                      // |diff| to sort
    const int nchan = pr.nchan, nwords = (nchan + 31) >> 5;
    const int P = 4 * pr.ncomp;
    for (long long pix = (long long)blockIdx.x * AUX_WARPS + warp; pix < pr.npix;
         pix += (long long)gridDim.x * AUX_WARPS) {
        const double* p = params + pix * P;
        const bool valid = params_valid(p, P);
        if (!valid) {   // no model: residual NaN everywhere
            if (lane == 0) {
                rms_out[pix] = NAN;
                chisq_out[pix] = NAN;
            }
            continue;
        }
        for (int c = 0; c < pr.ncomp; ++c) comp_setup_f64(pr, p + 4 * c, sh[warp].comps[c], lane);
        const long long row = row_index ? row_index[pix] : pix;
        const float* y = spectra + row * stride;
        for (int wd = 0; wd < nwords; ++wd) {
            const int i = (wd << 5) + lane;
            bool on = false;
            if (i < nchan) {
                const double m = model_f64(pr, tab, sh[warp].comps, pr.ncomp, i);
                const float d = y[i];
                if (model_f32) {
                    const float mf = (float)m;
                    on = mf > 0.f;
                    res[i] = (double)(d - mf);
                } else {
                    on = m > 0.0;
                    res[i] = (double)d - m;
                }
            }
            const unsigned b = __ballot_sync(FULL, on);
            if (lane == 0) bits[warp][wd] = b;
        }
        __syncwarp();
        dilate_bits(bits[warp], nchan, expand, lane);
        double c2 = 0.0;
        int ns = 0, nval = 0;
        for (int i = lane; i < npow2; i += 32) {
            double k = INFINITY;
            if (i < nchan) {
                const double r = res[i];
                const double rb = res[(i >= 2) ? i - 2 : i - 2 + nchan];
                double df = model_f32 ? (double)((float)r - (float)rb) : (r - rb);
                df = fabs(df);
                if (df == df) {
                    k = df;
                    ++nval;
                }
                if ((bits[warp][i >> 5] >> (i & 31)) & 1u) {
                    ++ns;
                    if (r == r) c2 += r * r;
                }
            }
            key[i] = k;
        }
        c2 = warp_sum(c2);
        ns = warp_sum_i(ns);
        nval = warp_sum_i(nval);
        __syncwarp();
        for (int k = 2; k <= npow2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = lane; t < (npow2 >> 1); t += 32) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
                    const int l = i | j;
                    const bool up = ((i & k) == 0);
                    const double a = key[i], b = key[l];
                    if ((a > b) == up) {
                        key[i] = b;
                        key[l] = a;
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0) {
            double med = NAN;
            if (nval > 0) {
                if (nval & 1)
                    med = key[nval >> 1];
                else
                    med = model_f32 ? (double)(((float)key[(nval >> 1) - 1] + (float)key[nval >> 1]) * 0.5f)
                                    : 0.5 * (key[(nval >> 1) - 1] + key[nval >> 1]);
            }
            double rms, rms2;
            if (model_f32) {   // numpy keeps float32: 1.4826 * med / 2**0.5, then rms**2, all in float32
                const float rf = (1.4826f * (float)med) / 1.4142135623730951f;
                rms = (double)rf;
                rms2 = (double)(rf * rf);
            } else {
                rms = 1.4826 * med / 1.4142135623730951;
                rms2 = rms * rms;
            }
            rms_out[pix] = rms;
            const double red = (ns == 0) ? NAN : (double)ns;
            chisq_out[pix] = (c2 / red) / rms2;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Robust noise + peak per pixel: signals.get_rms_robust (mufasa/signals.py:48-121) and the per-pixel half of get_snr
// (:24-45) / multi_v_fit.snr_estimate (:677-702), one warp per pixel, fp64 like numpy on a float64 copy:
//   rms0 = 1.4826... * nanmedian(|y - nanmedian(y)|)            over the included (finite, planemask) channels
//   sig  = dilate(y > sigma_cut * rms0, expand along the spectral axis)                       (refine_rms :94-121)
//   rms  = the same MAD over included & ~sig channels;   peak = max over the included channels
// Medians by bitonic sort in shared memory (npow2 >= nchan keys, +inf padding).
constexpr double MAD_TO_STD = 1.482602218505602;

__device__ __forceinline__ void warp_bitonic_sort(double* key, int npow2, int lane) {
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (npow2 >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j clear
                const int l = i | j;
                const bool up = ((i & k) == 0);
                const double a = key[i], b = key[l];
                if ((a > b) == up) {
                    key[i] = b;
                    key[l] = a;
                }
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ double sorted_median(const double* key, int n) {   // numpy: mean of the two middle values
    if (n <= 0) return NAN;
    return (n & 1) ? key[n >> 1] : 0.5 * (key[(n >> 1) - 1] + key[n >> 1]);
}

// MAD-based sigma of the channels flagged in ``inc`` (bit set); warp-collective, result on every lane
__device__ __forceinline__ double warp_mad_std(const float* __restrict__ y, const unsigned* inc, int nchan, int npow2,
                                               double* key, int lane) {
    int n = 0;
    for (int i = lane; i < npow2; i += 32) {
        const bool on = (i < nchan) && ((inc[i >> 5] >> (i & 31)) & 1u);
        key[i] = on ? (double)y[i] : INFINITY;
        n += on ? 1 : 0;
    }
    n = warp_sum_i(n);
    __syncwarp();
    warp_bitonic_sort(key, npow2, lane);
    const double med = sorted_median(key, n);
    __syncwarp();
    for (int i = lane; i < npow2; i += 32) {
        const bool on = (i < nchan) && ((inc[i >> 5] >> (i & 31)) & 1u);
        key[i] = on ? fabs((double)y[i] - med) : INFINITY;
    }
    __syncwarp();
    warp_bitonic_sort(key, npow2, lane);
    const double mad = sorted_median(key, n);
    __syncwarp();
    return MAD_TO_STD * mad;
}

__global__ void __launch_bounds__(AUX_WARPS * 32)
    rms_robust_kernel(int nchan, long long npix, const float* __restrict__ spectra, long long stride,
                      const long long* __restrict__ row_index, const unsigned char* __restrict__ planemask,
                      double sigma_cut, int expand, int npow2, double* __restrict__ rms0_out,
                      double* __restrict__ rms_out, double* __restrict__ peak_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned inc[AUX_WARPS][MASK_WORDS];
    __shared__ unsigned sig[AUX_WARPS][MASK_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* key = reinterpret_cast<double*>(smem_raw) + (size_t)warp * npow2;
    const int nwords = (nchan + 31) >> 5;
    for (long long pix = (long long)blockIdx.x * AUX_WARPS + warp; pix < npix; pix += (long long)gridDim.x * AUX_WARPS) {
        const long long row = row_index ? row_index[pix] : pix;
        const float* y = spectra + row * stride;
        const bool plane = planemask ? (planemask[pix] != 0) : true;
        double peak = -INFINITY;
        for (int wd = 0; wd < nwords; ++wd) {
            const int i = (wd << 5) + lane;
            const float v = (i < nchan) ? y[i] : NAN;
            const bool on = plane && (v - v == 0.f);                 // finite
            if (on) peak = fmax(peak, (double)v);
            const unsigned b = __ballot_sync(FULL, on);
            if (lane == 0) inc[warp][wd] = b;
        }
        peak = warp_max(peak);
        __syncwarp();
        const double rms0 = warp_mad_std(y, inc[warp], nchan, npow2, key, lane);
        // signal mask on the RAW data (LazyComparisonMask cube > rms * sigma_cut), dilated along the spectral axis
        const double thr = sigma_cut * rms0;
        for (int wd = 0; wd < nwords; ++wd) {
            const int i = (wd << 5) + lane;
            const bool on = (i < nchan) && ((double)y[i] > thr);    // NaN compares false
            const unsigned b = __ballot_sync(FULL, on);
            if (lane == 0) sig[warp][wd] = b;
        }
        __syncwarp();
        dilate_bits(sig[warp], nchan, expand, lane);
        for (int wd = lane; wd < nwords; wd += 32) sig[warp][wd] = inc[warp][wd] & ~sig[warp][wd];
        __syncwarp();
        const double rms = warp_mad_std(y, sig[warp], nchan, npow2, key, lane);
        if (lane == 0) {
            rms0_out[pix] = rms0;
            rms_out[pix] = rms;
            peak_out[pix] = (peak == -INFINITY) ? NAN : peak;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Cube layout: [nchan][nplane] (FITS) -> rows [npix][stride], compacted by pix_index, NaN padded, optional flip.
__global__ void __launch_bounds__(256)
    gather_kernel(const float* __restrict__ cube, int nchan, long long nplane, const long long* __restrict__ pix_index,
                  long long npix, int flip, float* __restrict__ out, long long stride) {
    __shared__ float tile[32][33];
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    // read: rows = channels, columns = pixels (coalesced when the selected pixels are contiguous)
    long long src_pix = -1;
    if (p0 + tx < npix) src_pix = pix_index ? pix_index[p0 + tx] : (p0 + tx);
    for (int r = ty; r < 32; r += 8) {
        const int ch = c0 + r;
        float v = NAN;
        if (ch < nchan && src_pix >= 0) {
            const int sch = flip ? (nchan - 1 - ch) : ch;
            v = __ldcs(cube + (long long)sch * nplane + src_pix);
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    // write: rows = pixels, columns = channels (coalesced)
    for (int r = ty; r < 32; r += 8) {
        const long long p = p0 + r;
        const int ch = c0 + tx;
        if (p < npix && ch < stride) out[p * stride + ch] = tile[tx][r];
    }
}

}  // namespace b200fit

// =============================================================================================================
// C ABI
using namespace b200fit;

#include "b200fit_ctx.h"

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <int NC>
static size_t fit_workspace_bytes(long long npix) {
    return sizeof(FitQueues) + 2 * align_up((size_t)npix * sizeof(unsigned), 256) +
           (size_t)npix * sizeof(PixState<4 * NC>) + 256;
}

// Round loop.  Every pixel finishes within 2*max_iter + 3 rounds (lm_update caps evaluations); the host checks
// the active count every CHECK rounds through a pinned counter, two checks behind the launches, so the GPU never
// idles waiting for the host and at most 2*CHECK empty rounds are launched after the last pixel converged.
// A FitJob is one fit in flight; b200fit_fit_batch interleaves the rounds of several jobs on separate streams so
// that the latency-bound late rounds of one fit (few slowly converging pixels left) overlap another fit's work.
struct FitJob {
    int nc = 0, model_id = 0;
    DeviceProblem pr;
    const float* d_spectra = nullptr;
    long long stride = 0;
    const long long* d_row_index = nullptr;
    const double* d_guesses = nullptr;
    const double* d_err = nullptr;
    double *d_params = nullptr, *d_perr = nullptr, *d_chi2 = nullptr, *d_err_used = nullptr;
    int* d_status = nullptr;
    unsigned* d_counts = nullptr;
    FitQueues* fq = nullptr;
    unsigned* lists[2] = {nullptr, nullptr};
    void* states = nullptr;
    cudaStream_t stream = nullptr;
    unsigned egrid = 0, egrid_team = 0, sgrid = 0, sgrid_small = 0;
    unsigned seen = 0xffffffffu, small_limit = 0;   // last active count the host has seen; <= small_limit: small solver grid
    size_t esmem = 0, esmem_team = 0;   // dynamic shared memory of the evaluation kernels (staging buffers)
    bool use_graph = false;             // rounds are launched as captured groups (job_rounds)
    cudaGraphExec_t gexec[4] = {nullptr, nullptr, nullptr, nullptr};   // ... one executable graph per evaluation variant
                                                                        //     and solver grid (index team + 2 small)
    void* wait_event = nullptr;
    int r = 0, max_rounds = 0, nchecks = 0, slot0 = 0;
    bool finished = false, prof = false;
};

constexpr int FIT_CHECK = 8, FIT_RING = 4;

template <int NC>
static int job_begin(b200fit_ctx* ctx, FitJob& jb, void* d_workspace) {
    constexpr int P = 4 * NC;
    const DeviceProblem& pr = jb.pr;
    unsigned char* ws = (unsigned char*)d_workspace;
    jb.fq = (FitQueues*)ws;
    const size_t list_bytes = align_up((size_t)pr.npix * sizeof(unsigned), 256);
    jb.lists[0] = (unsigned*)(ws + sizeof(FitQueues));
    jb.lists[1] = (unsigned*)(ws + sizeof(FitQueues) + list_bytes);
    jb.states = ws + sizeof(FitQueues) + 2 * list_bytes;
    if (pr.npix >= (1LL << 32)) return fail(ctx, B200FIT_E_ARG, "fit: npix must be < 2^32 per call");
    cudaStream_t stream = jb.stream;
    CK(cudaMemsetAsync(jb.fq, 0, sizeof(FitQueues), stream));
    if (jb.prof) {
        ctx->prof_used = 0;
        ctx->prof_rounds = 0;
        ctx->next_event(stream);   // [0] before the prologue
    }
    const long long pwarps = (pr.npix + EVAL_WARPS - 1) / EVAL_WARPS;
    const long long cap = (long long)ctx->num_sms * 16;
    const unsigned pgrid = (unsigned)(pwarps < cap ? pwarps : cap);
    fit_prologue_kernel<NC><<<pgrid, EVAL_WARPS * 32, 0, stream>>>(
        pr, jb.d_spectra, jb.stride, jb.d_row_index, jb.d_guesses, jb.d_err, (PixState<P>*)jb.states, jb.lists[0], jb.fq,
        jb.d_params, jb.d_perr, jb.d_chi2, jb.d_err_used, jb.d_status, jb.d_counts);
    CK(cudaGetLastError());
    ctx->launches += 1;
    if (jb.prof) ctx->next_event(stream);   // [1] after the prologue

    // the evaluation kernel's hyperfine groups are compile-time: they must be the ones the host derived
    {
        bool same = true;
        if (jb.model_id == B200FIT_MODEL_NH3_11) {
            using LG = LineGroups<B200FIT_MODEL_NH3_11>;
            same = pr.ngroups == LG::NG;
            for (int g = 0; same && g <= LG::NG; ++g) same = pr.gstart[g] == LG::start(g);
        } else {
            using LG = LineGroups<B200FIT_MODEL_N2HP_10>;
            same = pr.ngroups == LG::NG;
            for (int g = 0; same && g <= LG::NG; ++g) same = pr.gstart[g] == LG::start(g);
        }
        if (!same) return fail(ctx, B200FIT_E_MODEL, "fit: hyperfine group table mismatch");
    }
    // the evaluation kernels stage spectrum rows in dynamic shared memory: two stages per pixel in flight
    jb.esmem = eval_dyn_smem<NC, 1>(jb.stride);
    jb.esmem_team = eval_dyn_smem<NC, EVAL_WARPS>(jb.stride);
    int eval_per_sm = 0, team_per_sm = 0;
    if (jb.model_id == B200FIT_MODEL_NH3_11) {
        CK(cudaFuncSetAttribute(fit_eval_kernel<NC, B200FIT_MODEL_NH3_11, 1>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jb.esmem));
        CK(cudaFuncSetAttribute(fit_eval_kernel<NC, B200FIT_MODEL_NH3_11, EVAL_WARPS>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jb.esmem_team));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&eval_per_sm, fit_eval_kernel<NC, B200FIT_MODEL_NH3_11, 1>,
                                                         EVAL_WARPS * 32, jb.esmem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &team_per_sm, fit_eval_kernel<NC, B200FIT_MODEL_NH3_11, EVAL_WARPS>, EVAL_WARPS * 32, jb.esmem_team));
    } else {
        CK(cudaFuncSetAttribute(fit_eval_kernel<NC, B200FIT_MODEL_N2HP_10, 1>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jb.esmem));
        CK(cudaFuncSetAttribute(fit_eval_kernel<NC, B200FIT_MODEL_N2HP_10, EVAL_WARPS>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)jb.esmem_team));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&eval_per_sm, fit_eval_kernel<NC, B200FIT_MODEL_N2HP_10, 1>,
                                                         EVAL_WARPS * 32, jb.esmem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &team_per_sm, fit_eval_kernel<NC, B200FIT_MODEL_N2HP_10, EVAL_WARPS>, EVAL_WARPS * 32, jb.esmem_team));
    }
    if (eval_per_sm < 1 || team_per_sm < 1) return fail(ctx, B200FIT_E_ARG, "fit_eval_kernel does not fit on an SM");
    long long egrid = (long long)ctx->num_sms * eval_per_sm;        // a whole number of CTAs per SM
    long long egrid_team = (long long)ctx->num_sms * team_per_sm;   // block-per-pixel rounds
    if (egrid_team > pr.npix) egrid_team = pr.npix;
    if (egrid > pwarps) egrid = pwarps;
    const long long sgroups = SOLVE_THREADS / P;                     // pixels per solver block and pass
    const long long sblocks = (pr.npix + sgroups - 1) / sgroups;
    long long sgrid = (long long)ctx->num_sms * 32;
    if (sgrid > sblocks) sgrid = sblocks;
    jb.egrid = (unsigned)egrid;
    jb.egrid_team = (unsigned)egrid_team;
    jb.sgrid = (unsigned)sgrid;
    // Late rounds: the solver's latency is what the tail of a fit waits on, and scheduling num_sms x 32 blocks of which
    // nearly all leave at once is a good part of it.  Once the host has SEEN (with its usual lag; counts only fall) an
    // active count that two blocks per SM cover in one pass, the groups are launched with that grid.  Per-pixel
    // arithmetic does not depend on the grid.
    jb.sgrid_small = (unsigned)std::min<long long>(sgrid, (long long)ctx->num_sms * 2);
    jb.small_limit = (unsigned)(jb.sgrid_small * sgroups);
    jb.seen = 0xffffffffu;
    jb.max_rounds = (2 * pr.max_iter + 3 + FIT_CHECK - 1) / FIT_CHECK * FIT_CHECK;   // whole groups of rounds
    jb.r = 0;
    jb.nchecks = 0;
    jb.finished = false;
    return B200FIT_OK;
}

// The two kernels of one round on the job's stream (directly, or recorded into a graph being captured there).
template <int NC>
static void launch_round(FitJob& jb, int cur, bool team, bool small = false) {
    constexpr int P = 4 * NC;
    cudaStream_t stream = jb.stream;
    auto* ek = (jb.model_id == B200FIT_MODEL_NH3_11)
                   ? (team ? fit_eval_kernel<NC, B200FIT_MODEL_NH3_11, EVAL_WARPS>
                           : fit_eval_kernel<NC, B200FIT_MODEL_NH3_11, 1>)
                   : (team ? fit_eval_kernel<NC, B200FIT_MODEL_N2HP_10, EVAL_WARPS>
                           : fit_eval_kernel<NC, B200FIT_MODEL_N2HP_10, 1>);
    ek<<<team ? jb.egrid_team : jb.egrid, EVAL_WARPS * 32, team ? jb.esmem_team : jb.esmem, stream>>>(
        jb.pr, jb.d_spectra, jb.stride, jb.d_row_index, (PixState<P>*)jb.states, jb.lists[cur], jb.fq, cur);
    fit_solve_kernel<NC><<<small ? jb.sgrid_small : jb.sgrid, SOLVE_THREADS, 0, stream>>>(
        jb.pr, (PixState<P>*)jb.states, jb.lists[cur], jb.lists[cur ^ 1], jb.fq, cur, jb.d_params, jb.d_perr, jb.d_chi2,
        jb.d_err_used, jb.d_status, jb.d_counts);
}

// One GROUP of FIT_CHECK rounds, then the (asynchronous) look at the active count.  The group is a CUDA graph,
// captured once per job and evaluation variant and replayed: the late rounds of a fit are a chain of small dependent
// kernels, and a graph launch replaces 2 * FIT_CHECK stream launches (host time, and the gap between dependent
// kernels on the device).  Rounds after the last pixel finished find an empty list and return at once.
template <int NC>
static int job_rounds(b200fit_ctx* ctx, FitJob& jb) {
    cudaStream_t stream = jb.stream;
    const int r0 = jb.r;
    // late rounds: a block per pixel (see fit_eval_kernel); a function of the round only
    static_assert(fit_team_round<NC>() % FIT_CHECK == 0, "the evaluation variant changes between groups of rounds");
    const bool team = r0 >= fit_team_round<NC>();
    if (jb.use_graph) {
        const bool small = jb.seen <= jb.small_limit;
        cudaGraphExec_t& ge = jb.gexec[(team ? 1 : 0) + (small ? 2 : 0)];
        if (!ge) {
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            for (int r = 0; r < FIT_CHECK; ++r) launch_round<NC>(jb, r & 1, team, small);   // r0 is even: cur = r & 1
            const cudaError_t ce = cudaStreamEndCapture(stream, &graph);
            if (ce != cudaSuccess || !graph) return fail(ctx, B200FIT_E_CUDA, "fit: stream capture of a round group", ce);
            const cudaError_t ci = cudaGraphInstantiate(&ge, graph, 0);
            cudaGraphDestroy(graph);
            if (ci != cudaSuccess) return fail(ctx, B200FIT_E_CUDA, "fit: cudaGraphInstantiate", ci);
        }
        CK(cudaGraphLaunch(ge, stream));
        ctx->launches += 2 * FIT_CHECK;
        ctx->graph_kernels += 2 * FIT_CHECK;
        ctx->graph_launches += 1;
        jb.r = r0 + FIT_CHECK;
    } else {
        for (int r = r0; r < r0 + FIT_CHECK; ++r) {
            constexpr int P = 4 * NC;
            const int cur = r & 1;
            auto* ek = (jb.model_id == B200FIT_MODEL_NH3_11)
                           ? (team ? fit_eval_kernel<NC, B200FIT_MODEL_NH3_11, EVAL_WARPS>
                                   : fit_eval_kernel<NC, B200FIT_MODEL_NH3_11, 1>)
                           : (team ? fit_eval_kernel<NC, B200FIT_MODEL_N2HP_10, EVAL_WARPS>
                                   : fit_eval_kernel<NC, B200FIT_MODEL_N2HP_10, 1>);
            ek<<<team ? jb.egrid_team : jb.egrid, EVAL_WARPS * 32, team ? jb.esmem_team : jb.esmem, stream>>>(
                jb.pr, jb.d_spectra, jb.stride, jb.d_row_index, (PixState<P>*)jb.states, jb.lists[cur], jb.fq, cur);
            if (jb.prof) ctx->next_event(stream);   // [2 + 2r] after the evaluation of round r
            fit_solve_kernel<NC><<<(jb.seen <= jb.small_limit) ? jb.sgrid_small : jb.sgrid, SOLVE_THREADS, 0, stream>>>(
                jb.pr, (PixState<P>*)jb.states, jb.lists[cur], jb.lists[cur ^ 1], jb.fq, cur, jb.d_params, jb.d_perr,
                jb.d_chi2, jb.d_err_used, jb.d_status, jb.d_counts);
            ctx->launches += 2;
            if (jb.prof) {
                ctx->next_event(stream);            // [3 + 2r] after the solver step of round r
                ctx->prof_rounds = r + 1;
            }
        }
        jb.r = r0 + FIT_CHECK;
    }
    {
        const int cur = (jb.r - 1) & 1;             // the last round of the group appended to list cur ^ 1
        const int slot = jb.slot0 + jb.nchecks % FIT_RING;
        if (jb.nchecks >= FIT_RING - 2) {   // wait for the check issued (FIT_RING - 2) checks ago
            const int old = jb.slot0 + (jb.nchecks - (FIT_RING - 2)) % FIT_RING;
            CK(cudaEventSynchronize(ctx->check_ev[old]));
            jb.seen = ctx->h_count[old];
            if (jb.seen == 0) jb.finished = true;
        }
        if (!jb.finished) {
            CK(cudaMemcpyAsync((void*)&ctx->h_count[slot], &jb.fq->count[cur ^ 1], sizeof(unsigned),
                               cudaMemcpyDeviceToHost, stream));
            CK(cudaEventRecord(ctx->check_ev[slot], stream));
            ++jb.nchecks;
        }
    }
    if (jb.r >= jb.max_rounds) jb.finished = true;
    return B200FIT_OK;
}

// Executable graphs of finished calls are destroyed once the work that used them has completed.
static void retire_graphs(b200fit_ctx* ctx, bool wait) {
    if (ctx->retired.empty()) return;
    if (wait)
        cudaEventSynchronize(ctx->retired_ev);
    else if (cudaEventQuery(ctx->retired_ev) != cudaSuccess)
        return;
    for (cudaGraphExec_t g : ctx->retired) cudaGraphExecDestroy(g);
    ctx->retired.clear();
}

// (Re)build the per-channel table of the fp64 model when the spectral axis differs from the cached one.
static int ensure_chan_tables(b200fit_ctx* ctx, const DeviceProblem& pr, cudaStream_t stream) {
    if (pr.ngroups_full > 8) return fail(ctx, B200FIT_E_MODEL, "fp64 model: more than 8 hyperfine line groups");
    if (!ctx->d_chan) CK(cudaMalloc((void**)&ctx->d_chan, 3 * B200FIT_MAX_NCHAN * sizeof(double)));
    if (ctx->chan_n == pr.nchan && ctx->chan_key[0] == pr.v0 && ctx->chan_key[1] == pr.dv &&
        ctx->chan_key[2] == pr.nu_ref_ghz)
        return B200FIT_OK;
    chan_tables_kernel<<<(pr.nchan + 127) / 128, 128, 0, stream>>>(pr, ctx->d_chan);
    CK(cudaGetLastError());
    ctx->launches += 1;
    ctx->chan_n = pr.nchan;
    ctx->chan_key[0] = pr.v0;
    ctx->chan_key[1] = pr.dv;
    ctx->chan_key[2] = pr.nu_ref_ghz;
    return B200FIT_OK;
}

extern "C" {

int b200fit_abi_version(void) { return B200FIT_ABI_VERSION; }

int64_t b200fit_nchan_stride(int32_t nchan) { return ((int64_t)nchan + 31) / 32 * 32; }

int b200fit_create(b200fit_ctx** out, int device) {
    if (!out) return B200FIT_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return B200FIT_E_NOGPU;
    if (device < 0 || device >= ndev) return B200FIT_E_ARG;
    b200fit_ctx* ctx = new b200fit_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return B200FIT_E_CUDA;
    }
    ctx->num_sms = prop.multiProcessorCount;
    ctx->h_count = nullptr;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_count, B200FIT_MAX_BATCH * 4 * sizeof(unsigned)) != cudaSuccess) {
        delete ctx;
        return B200FIT_E_CUDA;
    }
    bool okc = cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming) == cudaSuccess;
    okc = okc && cudaEventCreateWithFlags(&ctx->retired_ev, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < B200FIT_MAX_BATCH * 4; ++i)
        okc = okc && cudaEventCreateWithFlags(&ctx->check_ev[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < B200FIT_MAX_BATCH; ++i) {
        okc = okc && cudaEventCreateWithFlags(&ctx->join_ev[i], cudaEventDisableTiming) == cudaSuccess;
        okc = okc && cudaStreamCreateWithFlags(&ctx->job_stream[i], cudaStreamNonBlocking) == cudaSuccess;
    }
    if (!okc) {
        delete ctx;
        return B200FIT_E_CUDA;
    }
    *out = ctx;
    return B200FIT_OK;
}

void b200fit_destroy(b200fit_ctx* ctx) {
    if (!ctx) return;
    retire_graphs(ctx, true);
    cudaEventDestroy(ctx->retired_ev);
    for (int i = 0; i < B200FIT_MAX_BATCH * 4; ++i) cudaEventDestroy(ctx->check_ev[i]);
    for (int i = 0; i < B200FIT_MAX_BATCH; ++i) {
        cudaEventDestroy(ctx->join_ev[i]);
        cudaStreamDestroy(ctx->job_stream[i]);
    }
    cudaEventDestroy(ctx->fork_ev);
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    if (ctx->h_count) cudaFreeHost((void*)ctx->h_count);
    if (ctx->d_chan) cudaFree(ctx->d_chan);
    if (ctx->d_redo) cudaFree(ctx->d_redo);
    delete ctx;
}

const char* b200fit_last_error(const b200fit_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int b200fit_num_sms(const b200fit_ctx* ctx) { return ctx ? ctx->num_sms : 0; }

uint64_t b200fit_launch_count(const b200fit_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t b200fit_host_launch_count(const b200fit_ctx* ctx) {
    return ctx ? ctx->launches - ctx->graph_kernels + ctx->graph_launches : 0;
}

}  // extern "C"

namespace b200fit {
// XU (special-function unit) micro-benchmark: 8 independent ex2.approx chains per thread, nothing else in the loop
// but the loop counter -- the measured peak bench.py holds the evaluation kernel's executed ex2 rate against
// ("peak measured by an in-repo micro-benchmark on the same box").
__global__ void __launch_bounds__(256) xu_peak_kernel(float* __restrict__ sink, int iters) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = -1e-3f * (float)(threadIdx.x + k + 1);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = b2_ex2(x[k]) - 1.0009765625f;   // stays in (-1, 0): never saturates
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += x[k];
    if (sum == 12345.678f) sink[0] = sum;   // keeps the chains alive
}
}  // namespace b200fit

extern "C" {

int b200fit_measure_xu_peak(b200fit_ctx* ctx, double* ex2_per_second, void* stream_) {
    if (!ctx || !ex2_per_second) return B200FIT_E_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    float* d_sink = nullptr;
    CK(cudaMalloc((void**)&d_sink, sizeof(float)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int blocks = ctx->num_sms * 8, iters = 4096;
    xu_peak_kernel<<<blocks, 256, 0, stream>>>(d_sink, 64);          // warm-up
    CK(cudaEventRecord(e0, stream));
    xu_peak_kernel<<<blocks, 256, 0, stream>>>(d_sink, iters);
    CK(cudaEventRecord(e1, stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_sink);
    ctx->launches += 2;
    *ex2_per_second = (double)blocks * 256.0 * 32.0 * (double)iters / ((double)ms * 1e-3);
    return B200FIT_OK;
}

int b200fit_set_profiling(b200fit_ctx* ctx, int32_t on) {
    if (!ctx) return B200FIT_E_ARG;
    ctx->profiling = on ? 1 : 0;
    ctx->prof_used = 0;
    ctx->prof_rounds = 0;
    return B200FIT_OK;
}

int b200fit_get_profile(b200fit_ctx* ctx, double* out) {
    if (!ctx || !out) return B200FIT_E_ARG;
    for (int i = 0; i < 4; ++i) out[i] = 0.0;
    const int rounds = ctx->prof_rounds;
    if (ctx->prof_used < (size_t)(2 + 2 * rounds) || ctx->prof_used < 2) return B200FIT_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->prof_ev[1 + 2 * rounds]));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->prof_ev[0], ctx->prof_ev[1]));
    out[0] = ms;
    for (int r = 0; r < rounds; ++r) {
        CK(cudaEventElapsedTime(&ms, ctx->prof_ev[1 + 2 * r], ctx->prof_ev[2 + 2 * r]));
        out[1] += ms;
        CK(cudaEventElapsedTime(&ms, ctx->prof_ev[2 + 2 * r], ctx->prof_ev[3 + 2 * r]));
        out[2] += ms;
    }
    out[3] = rounds;
    return B200FIT_OK;
}

size_t b200fit_workspace_bytes(const b200fit_problem* prob) {
    // queues + two active lists + the per-pixel LM state that crosses the round loop's kernels
    if (!prob || prob->npix < 0) return 0;
    const size_t a = fit_workspace_bytes<1>(prob->npix), b = fit_workspace_bytes<2>(prob->npix);
    const size_t fit = (prob->ncomp == 1) ? a : b;
    return fit < 256 ? 256 : fit;
}

int b200fit_gather_spectra(b200fit_ctx* ctx, const float* d_cube, int32_t nchan, int64_t nplane,
                           const int64_t* d_pix_index, int64_t npix, int32_t flip, float* d_spectra,
                           int64_t nchan_stride, void* stream) {
    if (!ctx || !d_cube || !d_spectra || nchan <= 0 || nchan_stride < nchan || (nchan_stride & 31)) return fail(ctx, B200FIT_E_ARG, "gather_spectra: bad argument");
    if (npix == 0) return B200FIT_OK;
    CK(cudaSetDevice(ctx->device));
    dim3 grid((unsigned)((npix + 31) / 32), (unsigned)(nchan_stride / 32));
    gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_cube, nchan, nplane, (const long long*)d_pix_index, npix,
                                                          flip, d_spectra, nchan_stride);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_upload_slab(b200fit_ctx* ctx, const float* h_cube, int32_t nchan, int64_t plane, int64_t first, int64_t count,
                        float* d_dst, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (nchan == 0 || count == 0) return B200FIT_OK;
    if (!h_cube || !d_dst || nchan < 0 || plane <= 0 || first < 0 || count < 0 || first + count > plane)
        return fail(ctx, B200FIT_E_ARG, "upload_slab: bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(d_dst, (size_t)count * sizeof(float), h_cube + first, (size_t)plane * sizeof(float),
                         (size_t)count * sizeof(float), (size_t)nchan, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return B200FIT_OK;
}

int b200fit_fit_batch(b200fit_ctx* ctx, const b200fit_fit_args* args, int32_t njobs, void* stream_) {
    if (!ctx || !args) return fail(ctx, B200FIT_E_ARG, "fit: null context/arguments");
    if (njobs < 0 || njobs > B200FIT_MAX_BATCH) return fail(ctx, B200FIT_E_ARG, "fit: njobs must be in [0, B200FIT_MAX_BATCH]");
    cudaStream_t stream = (cudaStream_t)stream_;
    FitJob jobs[B200FIT_MAX_BATCH];
    int live = 0;
    for (int j = 0; j < njobs; ++j) {
        const b200fit_fit_args& a = args[j];
        if (!a.prob) return fail(ctx, B200FIT_E_ARG, "fit: null problem");
        FitJob& jb = jobs[live];
        const int rc = make_device_problem(*a.prob, jb.pr);
        if (rc) return fail(ctx, rc, "fit: invalid problem description");
        if (jb.pr.npix == 0) continue;
        if (!a.d_spectra || !a.d_guesses || !a.d_params || !a.d_perr || !a.d_chi2 || !a.d_err_used || !a.d_status ||
            !a.d_workspace)
            return fail(ctx, B200FIT_E_ARG, "fit: null device pointer");
        if (a.nchan_stride < jb.pr.nchan || (a.nchan_stride & 31)) return fail(ctx, B200FIT_E_ARG, "fit: bad nchan_stride");
        if (jb.pr.weight_mode == 1 && !a.d_err) return fail(ctx, B200FIT_E_ARG, "fit: weight_mode 1 needs d_err");
        jb.nc = jb.pr.ncomp;
        jb.model_id = a.prob->model_id;
        jb.d_spectra = a.d_spectra;
        jb.stride = a.nchan_stride;
        jb.d_row_index = (const long long*)a.d_row_index;
        jb.d_guesses = a.d_guesses;
        jb.d_err = a.d_err;
        jb.d_params = a.d_params;
        jb.d_perr = a.d_perr;
        jb.d_chi2 = a.d_chi2;
        jb.d_err_used = a.d_err_used;
        jb.d_status = a.d_status;
        jb.d_counts = a.d_counts;
        jb.wait_event = a.wait_event;
        jb.slot0 = live * FIT_RING;
        jb.states = a.d_workspace;   // resolved in job_begin
        ++live;
    }
    if (live == 0) return B200FIT_OK;
    CK(cudaSetDevice(ctx->device));
    retire_graphs(ctx, false);
    // Round groups are replayed as CUDA graphs, captured on the library's own streams (the caller's may be the legacy
    // default stream, which cannot be captured): every job is forked off the caller's stream, also a single one.
    // A profiled fit (events between the kernels of every round) is launched kernel by kernel on the caller's stream.
    const bool profiled = ctx->profiling != 0 && live == 1;
    const bool forked = !profiled;
    if (forked) CK(cudaEventRecord(ctx->fork_ev, stream));
    for (int j = 0; j < live; ++j) {
        FitJob& jb = jobs[j];
        jb.stream = forked ? ctx->job_stream[j] : stream;
        jb.prof = profiled;
        jb.use_graph = forked;
        if (forked) CK(cudaStreamWaitEvent(jb.stream, ctx->fork_ev, 0));
        if (jb.wait_event) CK(cudaStreamWaitEvent(jb.stream, (cudaEvent_t)jb.wait_event, 0));
        void* ws = jb.states;
        const int rc = (jb.nc == 1) ? job_begin<1>(ctx, jb, ws) : job_begin<2>(ctx, jb, ws);
        if (rc) return rc;
    }
    for (bool any = true; any;) {
        any = false;
        for (int j = 0; j < live; ++j) {
            FitJob& jb = jobs[j];
            if (jb.finished) continue;
            const int rc = (jb.nc == 1) ? job_rounds<1>(ctx, jb) : job_rounds<2>(ctx, jb);
            if (rc) return rc;
            any = true;
        }
    }
    for (int j = 0; j < live; ++j) {   // parameter errors + result rows, once per fit
        FitJob& jb = jobs[j];
        const long long groups = SOLVE_THREADS / (4 * jb.nc);
        long long fgrid = (jb.pr.npix + groups - 1) / groups;
        if (fgrid > (long long)ctx->num_sms * 32) fgrid = (long long)ctx->num_sms * 32;
        if (jb.nc == 1)
            fit_finalize_kernel<1><<<(unsigned)fgrid, SOLVE_THREADS, 0, jb.stream>>>(
                jb.pr, (const PixState<4>*)jb.states, jb.d_params, jb.d_perr, jb.d_chi2, jb.d_err_used, jb.d_status,
                jb.d_counts);
        else
            fit_finalize_kernel<2><<<(unsigned)fgrid, SOLVE_THREADS, 0, jb.stream>>>(
                jb.pr, (const PixState<8>*)jb.states, jb.d_params, jb.d_perr, jb.d_chi2, jb.d_err_used, jb.d_status,
                jb.d_counts);
        CK(cudaGetLastError());
        ctx->launches += 1;
    }
    if (forked) {
        for (int j = 0; j < live; ++j) {
            CK(cudaEventRecord(ctx->join_ev[j], jobs[j].stream));
            CK(cudaStreamWaitEvent(stream, ctx->join_ev[j], 0));
        }
    }
    for (int j = 0; j < live; ++j)
        for (cudaGraphExec_t g : jobs[j].gexec)
            if (g) ctx->retired.push_back(g);
    if (!ctx->retired.empty()) CK(cudaEventRecord(ctx->retired_ev, stream));
    CK(cudaGetLastError());
    return B200FIT_OK;
}

int b200fit_fit(b200fit_ctx* ctx, const b200fit_problem* prob, const float* d_spectra, int64_t nchan_stride,
                const int64_t* d_row_index, const double* d_guesses, const double* d_err, double* d_params, double* d_perr, double* d_chi2,
                double* d_err_used, int32_t* d_status, uint32_t* d_counts, void* d_workspace, void* stream) {
    b200fit_fit_args a;
    a.prob = prob;
    a.d_spectra = d_spectra;
    a.nchan_stride = nchan_stride;
    a.d_row_index = d_row_index;
    a.d_guesses = d_guesses;
    a.d_err = d_err;
    a.d_params = d_params;
    a.d_perr = d_perr;
    a.d_chi2 = d_chi2;
    a.d_err_used = d_err_used;
    a.d_status = d_status;
    a.d_counts = d_counts;
    a.d_workspace = d_workspace;
    a.wait_event = nullptr;
    return b200fit_fit_batch(ctx, &a, 1, stream);
}

// Grid of the warp-per-pixel statistics kernels.  Pixels differ a lot in cost (no fit / 1 / 2 components, line width),
// so the work is cut into many more blocks than fit on the device -- 4 pixels per warp -- and the hardware scheduler
// balances them; a grid of exactly 8 blocks per SM left a second, partly filled wave (5-6 blocks are resident).
static long long aux_blocks(const b200fit_ctx* ctx, long long npix) {
    const long long need = (npix + AUX_WARPS - 1) / AUX_WARPS;
    long long blocks = (need + 3) / 4;
    const long long floor_blocks = (long long)ctx->num_sms * 8;
    if (blocks < floor_blocks) blocks = floor_blocks;
    return blocks < need ? blocks : (need > 0 ? need : 1);
}

int b200fit_model(b200fit_ctx* ctx, const b200fit_problem* prob, const double* d_params, double* d_model,
                  int64_t nchan_stride, void* stream) {
    if (!ctx || !prob) return fail(ctx, B200FIT_E_ARG, "model: null context/problem");
    DeviceProblem pr;
    const int rc = make_device_problem(*prob, pr);
    if (rc) return fail(ctx, rc, "model: invalid problem description");
    if (pr.npix == 0) return B200FIT_OK;
    if (!d_params || !d_model || nchan_stride < pr.nchan) return fail(ctx, B200FIT_E_ARG, "model: bad argument");
    CK(cudaSetDevice(ctx->device));
    if (int rc = ensure_chan_tables(ctx, pr, (cudaStream_t)stream)) return rc;
    model_kernel<<<(unsigned)aux_blocks(ctx, pr.npix), AUX_WARPS * 32, 0, (cudaStream_t)stream>>>(
        pr, ctx->d_chan, d_params, d_model, nchan_stride);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_mask_argmax(b200fit_ctx* ctx, const int32_t* d_mask_counts, int64_t npix, int64_t index_offset,
                        int64_t* d_best, void* d_workspace, void* stream) {
    if (!ctx || !d_mask_counts || !d_best || !d_workspace || npix < 0)
        return fail(ctx, B200FIT_E_ARG, "mask_argmax: bad argument");
    CK(cudaSetDevice(ctx->device));
    unsigned long long* key = (unsigned long long*)d_workspace + 8;   // slot 0 is the fit work queue
    CK(cudaMemsetAsync(key, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    if (npix > 0) {
        long long blocks = (npix + 255) / 256;
        if (blocks > (long long)ctx->num_sms * 8) blocks = (long long)ctx->num_sms * 8;
        mask_argmax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_mask_counts, npix, index_offset, key);
        CK(cudaGetLastError());
        ctx->launches += 1;
    }
    mask_argmax_finish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(key, (long long*)d_best);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_mask_bits_pair(b200fit_ctx* ctx, const b200fit_problem* prob, int32_t ncomp_a, int32_t ncomp_b,
                           const double* d_params1, const double* d_params2, int32_t model_f32, const int64_t* d_best,
                           int64_t pix, uint32_t* d_bits, void* stream) {
    if (!ctx || !prob) return fail(ctx, B200FIT_E_ARG, "mask_bits: null context/problem");
    if (ncomp_a < 1 || ncomp_a > 2 || ncomp_b < 1 || ncomp_b > 2) return fail(ctx, B200FIT_E_ARG, "mask_bits: ncomp must be 1 or 2");
    DeviceProblem pr;
    const int rc = make_device_problem(*prob, pr);
    if (rc) return fail(ctx, rc, "mask_bits: invalid problem description");
    if (!d_bits || (!d_params1 && !d_params2)) return fail(ctx, B200FIT_E_ARG, "mask_bits: bad argument");
    CK(cudaSetDevice(ctx->device));
    if (int rc = ensure_chan_tables(ctx, pr, (cudaStream_t)stream)) return rc;
    mask_bits_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pr, ctx->d_chan, ncomp_a, ncomp_b, d_params1, d_params2,
                                                        model_f32, (const long long*)d_best, pix, d_bits);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_mask_bits(b200fit_ctx* ctx, const b200fit_problem* prob, const double* d_params1,
                      const double* d_params2, int32_t model_f32, const int64_t* d_best, int64_t pix,
                      uint32_t* d_bits, void* stream) {
    return b200fit_mask_bits_pair(ctx, prob, 1, 2, d_params1, d_params2, model_f32, d_best, pix, d_bits, stream);
}

int b200fit_aicc_pair(b200fit_ctx* ctx, const b200fit_problem* prob, int32_t ncomp_a, int32_t ncomp_b,
                      const float* d_spectra, int64_t nchan_stride, const int64_t* d_row_index,
                      const double* d_params1, const double* d_params2, const uint32_t* d_fill_bits,
                      int32_t only_empty, int32_t expand, int32_t model_f32, const double* thr, double* d_rss,
                      double* d_nsamp, double* d_lnk, signed char* d_ncomp, int32_t* d_mask_counts, void* stream) {
    return b200fit_aicc_pair_masked(ctx, prob, ncomp_a, ncomp_b, d_spectra, nchan_stride, d_row_index, d_params1,
                                    d_params2, nullptr, d_fill_bits, only_empty, expand, model_f32, thr, d_rss, d_nsamp,
                                    d_lnk, d_ncomp, d_mask_counts, stream);
}

int b200fit_aicc_pair_masked(b200fit_ctx* ctx, const b200fit_problem* prob, int32_t ncomp_a, int32_t ncomp_b,
                             const float* d_spectra, int64_t nchan_stride, const int64_t* d_row_index,
                             const double* d_params1, const double* d_params2, const uint32_t* d_mask_bits,
                             const uint32_t* d_fill_bits, int32_t only_empty, int32_t expand, int32_t model_f32,
                             const double* thr, double* d_rss, double* d_nsamp, double* d_lnk, signed char* d_ncomp,
                             int32_t* d_mask_counts, void* stream) {
    if (!ctx || !prob) return fail(ctx, B200FIT_E_ARG, "aicc: null context/problem");
    if (ncomp_a < 1 || ncomp_a > 2 || ncomp_b < 1 || ncomp_b > 2) return fail(ctx, B200FIT_E_ARG, "aicc: ncomp must be 1 or 2");
    DeviceProblem pr;
    const int rc = make_device_problem(*prob, pr);
    if (rc) return fail(ctx, rc, "aicc: invalid problem description");
    if (pr.npix == 0) return B200FIT_OK;
    if (!d_spectra || !d_params1 || !d_params2 || !d_rss || !d_nsamp || !d_lnk || !d_ncomp || nchan_stride < pr.nchan)
        return fail(ctx, B200FIT_E_ARG, "aicc: bad argument");
    if (expand < 0 || expand > 64) return fail(ctx, B200FIT_E_ARG, "aicc: expand must be in [0, 64]");
    if (only_empty && (!d_mask_counts || !d_fill_bits))
        return fail(ctx, B200FIT_E_ARG, "aicc: pass 2 needs the mask counts of pass 1 and the fill mask");
    CK(cudaSetDevice(ctx->device));
    const double t10 = thr ? thr[0] : 5.0, t20 = thr ? thr[1] : 5.0, t21 = thr ? thr[2] : 5.0;
    const size_t smem = (size_t)AUX_WARPS * 2 * nchan_stride * (model_f32 ? sizeof(float) : sizeof(double));
    CK(cudaFuncSetAttribute(aicc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (int rc = ensure_chan_tables(ctx, pr, (cudaStream_t)stream)) return rc;
    aicc_kernel<<<(unsigned)aux_blocks(ctx, pr.npix), AUX_WARPS * 32, smem, (cudaStream_t)stream>>>(
        pr, ctx->d_chan, ncomp_a, ncomp_b, d_spectra, nchan_stride, (const long long*)d_row_index, d_params1, d_params2, d_fill_bits,
        only_empty, expand, model_f32, t10, t20, t21, d_rss, d_nsamp, d_lnk, d_ncomp, d_mask_counts, d_mask_bits);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_aicc(b200fit_ctx* ctx, const b200fit_problem* prob, const float* d_spectra, int64_t nchan_stride,
                 const int64_t* d_row_index, const double* d_params1, const double* d_params2,
                 const uint32_t* d_fill_bits, int32_t only_empty, int32_t expand, int32_t model_f32, const double* thr,
                 double* d_rss, double* d_nsamp, double* d_lnk, signed char* d_ncomp, int32_t* d_mask_counts,
                 void* stream) {
    return b200fit_aicc_pair(ctx, prob, 1, 2, d_spectra, nchan_stride, d_row_index, d_params1, d_params2, d_fill_bits,
                             only_empty, expand, model_f32, thr, d_rss, d_nsamp, d_lnk, d_ncomp, d_mask_counts, stream);
}

int b200fit_prepare_guesses(b200fit_ctx* ctx, int64_t npix, int32_t npar, const double* d_raw, const double* lo,
                            const double* hi, double eps, double* d_guesses, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (npix == 0) return B200FIT_OK;
    if (!d_raw || !d_guesses || !lo || !hi || npix < 0 || (npar != 4 && npar != 8))
        return fail(ctx, B200FIT_E_ARG, "prepare_guesses: bad argument");
    CK(cudaSetDevice(ctx->device));
    DeviceLimits lim;
    for (int j = 0; j < 8; ++j) {
        lim.lo[j] = (j < npar) ? lo[j] : 0.0;
        lim.hi[j] = (j < npar) ? hi[j] : 0.0;
    }
    long long blocks = (npix + 255) / 256;
    if (blocks > (long long)ctx->num_sms * 16) blocks = (long long)ctx->num_sms * 16;
    prepare_guesses_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(npix, npar, d_raw, lim, eps, d_guesses);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_pack_selection(b200fit_ctx* ctx, int64_t npix, const double* d_params1, const double* d_perr1,
                           const double* d_params2, const double* d_perr2, const double* d_lnk,
                           const signed char* d_ncomp, double* d_out, int64_t out_stride, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (npix == 0) return B200FIT_OK;
    if (!d_params1 || !d_perr1 || !d_params2 || !d_perr2 || !d_lnk || !d_ncomp || !d_out || npix < 0 || out_stride < npix)
        return fail(ctx, B200FIT_E_ARG, "pack_selection: bad argument");
    CK(cudaSetDevice(ctx->device));
    long long blocks = (npix + 255) / 256;
    if (blocks > (long long)ctx->num_sms * 16) blocks = (long long)ctx->num_sms * 16;
    pack_selection_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(npix, d_params1, d_perr1, d_params2, d_perr2,
                                                                               d_lnk, d_ncomp, d_out, out_stride);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_rms_snr(b200fit_ctx* ctx, int32_t nchan, int64_t npix, const float* d_spectra, int64_t nchan_stride,
                    const int64_t* d_row_index, const unsigned char* d_planemask, double sigma_cut, int32_t expand,
                    double* d_rms0, double* d_rms, double* d_peak, void* stream) {
    if (!ctx) return B200FIT_E_ARG;
    if (npix == 0) return B200FIT_OK;
    if (!d_spectra || !d_rms0 || !d_rms || !d_peak || nchan < 1 || nchan > B200FIT_MAX_NCHAN || nchan_stride < nchan ||
        npix < 0)
        return fail(ctx, B200FIT_E_ARG, "rms_snr: bad argument");
    if (expand < 0 || expand > 64) return fail(ctx, B200FIT_E_ARG, "rms_snr: expand must be in [0, 64]");
    CK(cudaSetDevice(ctx->device));
    int npow2 = 64;
    while (npow2 < nchan) npow2 <<= 1;
    const size_t smem = (size_t)AUX_WARPS * npow2 * sizeof(double);
    CK(cudaFuncSetAttribute(rms_robust_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rms_robust_kernel<<<(unsigned)aux_blocks(ctx, npix), AUX_WARPS * 32, smem, (cudaStream_t)stream>>>(
        nchan, npix, d_spectra, nchan_stride, (const long long*)d_row_index, d_planemask, sigma_cut, expand, npow2,
        d_rms0, d_rms, d_peak);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

int b200fit_resid_stats(b200fit_ctx* ctx, const b200fit_problem* prob, const float* d_spectra, int64_t nchan_stride,
                        const int64_t* d_row_index, const double* d_params, int32_t expand, int32_t model_f32,
                        double* d_rms, double* d_chisq_red, void* stream) {
    if (!ctx || !prob) return fail(ctx, B200FIT_E_ARG, "resid_stats: null context/problem");
    DeviceProblem pr;
    const int rc = make_device_problem(*prob, pr);
    if (rc) return fail(ctx, rc, "resid_stats: invalid problem description");
    if (pr.npix == 0) return B200FIT_OK;
    if (!d_spectra || !d_params || !d_rms || !d_chisq_red || nchan_stride < pr.nchan)
        return fail(ctx, B200FIT_E_ARG, "resid_stats: bad argument");
    if (expand < 0 || expand > 64) return fail(ctx, B200FIT_E_ARG, "resid_stats: expand must be in [0, 64]");
    CK(cudaSetDevice(ctx->device));
    int npow2 = 64;
    while (npow2 < pr.nchan) npow2 <<= 1;
    const size_t smem = (size_t)AUX_WARPS * 2 * npow2 * sizeof(double);
    CK(cudaFuncSetAttribute(resid_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (int rc = ensure_chan_tables(ctx, pr, (cudaStream_t)stream)) return rc;
    resid_stats_kernel<<<(unsigned)aux_blocks(ctx, pr.npix), AUX_WARPS * 32, smem, (cudaStream_t)stream>>>(
        pr, ctx->d_chan, d_spectra, nchan_stride, (const long long*)d_row_index, d_params, expand, model_f32, npow2, d_rms,
        d_chisq_red);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return B200FIT_OK;
}

}  // extern "C"
