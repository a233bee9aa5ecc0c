// b200fit_ctx.h -- the per-device context behind the opaque b200fit_ctx handle, shared by the library's .cu files.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/b200fit.h"

struct b200fit_ctx {
    int device;
    int num_sms;
    std::string err;
    cudaEvent_t check_ev[B200FIT_MAX_BATCH * 4];
    volatile unsigned* h_count;   // pinned, 4 slots per job: active-pixel counts read back by the fit's round loop
    cudaStream_t job_stream[B200FIT_MAX_BATCH];
    cudaEvent_t fork_ev, join_ev[B200FIT_MAX_BATCH];
    std::vector<cudaGraphExec_t> retired;     // executable graphs of earlier fit calls, destroyed once retired_ev fired
    cudaEvent_t retired_ev = nullptr;
    unsigned long long launches = 0;          // kernels launched by this context (all entry points)
    unsigned long long graph_kernels = 0;     // ... of which inside replayed CUDA graphs
    unsigned long long graph_launches = 0;    // graph launches that carried them
    double* d_chan = nullptr;                 // per-channel constants of the fp64 model (3 x B200FIT_MAX_NCHAN)
    int chan_n = 0;                           // ... valid for this spectral axis: nchan, {v0, dv, nu_ref}
    double chan_key[3] = {0.0, 0.0, 0.0};
    int profiling = 0;                        // record CUDA events around the fit's kernels
    std::vector<cudaEvent_t> prof_ev;         // event pool (grown on demand)
    size_t prof_used = 0;
    int prof_rounds = 0;
    unsigned* d_redo = nullptr;               // convolution: tiles the fast separable kernel hands to the generic one
    size_t redo_cap = 0;

    cudaEvent_t next_event(cudaStream_t stream) {
        if (prof_used == prof_ev.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            prof_ev.push_back(e);
        }
        cudaEvent_t e = prof_ev[prof_used++];
        cudaEventRecord(e, stream);
        return e;
    }
};

static inline int fail(b200fit_ctx* ctx, int code, const char* what, cudaError_t ce = cudaSuccess) {
    if (ctx) {
        ctx->err = what;
        if (ce != cudaSuccess) {
            ctx->err += ": ";
            ctx->err += cudaGetErrorString(ce);
        }
    }
    return code;
}

#define CK(call)                                                    \
    do {                                                            \
        cudaError_t ce_ = (call);                                   \
        if (ce_ != cudaSuccess) return fail(ctx, B200FIT_E_CUDA, #call, ce_); \
    } while (0)

