// fd_host.h — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/flashb200.h"

namespace fd {

void set_error(const char* fmt, ...);

#define FD_CHECK_ARG(cond, ...)        \
    do {                               \
        if (!(cond)) {                 \
            fd::set_error(__VA_ARGS__); \
            return -1;                 \
        }                              \
    } while (0)

#define FD_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            fd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                     \
            return -2;                                                                   \
        }                                                                                \
    } while (0)

#define FD_CHECK_LAUNCH()                                                                \
    do {                                                                                 \
        fd::count_launch();                                                              \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            fd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                          __FILE__, __LINE__);                                           \
            return -3;                                                                   \
        }                                                                                \
    } while (0)

// Encode a bf16 tiled tensor map with 128-byte swizzle.  dims/strides inner->outer; strides in
// bytes for dims 1..rank-1.  Returns 0 on success.
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box);

// same with an explicit swizzle span in bytes (32 / 64 / 128; 0 = none) — the TMA-store epilogue stages 32- and
// 64-byte rows
int encode_tmap_bf16_sw(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

int num_sms();

// launch accounting (fd_launch_count) and optional per-launch CUDA-event profiling (fd_profile_*)
void count_launch();
bool profiling_on();
enum ProfCat { PROF_GEMM = 0, PROF_CONV = 1, PROF_ATTN_FWD = 2, PROF_ATTN_BWD = 3, PROF_NCAT = 4 };
struct ProfScope {
    cudaStream_t stream;
    int slot;
    ProfScope(cudaStream_t s, int cat, double work, int M = 0, int N = 0, int K = 0);   // work = algorithmic FLOPs
    ~ProfScope();
};

}  // namespace fd
