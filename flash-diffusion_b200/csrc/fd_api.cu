// fd_api.cu — error handling, device queries and TMA descriptor encoding for libflashb200.
#include <atomic>
#include <mutex>
#include <string.h>
#include <vector>

#include "fd_host.h"

namespace fd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// The driver entry point is resolved at run time so that the library links (and loads on a
// CPU-only box) without libcuda.
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e =
            cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    });
    return fn;
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
    return encode_tmap_bf16_sw(map, base, rank, dims, strides_bytes, box, 128);
}

int encode_tmap_bf16_sw(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    // cuTensorMapEncodeTiled is a DRIVER call: it needs the primary context current on this thread.
    // A thread that has not yet issued a runtime call that binds it (e.g. a PyTorch autograd worker
    // entering our backward first) would get CUDA_ERROR_INVALID_CONTEXT, so bind it once per thread.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        cudaFree(0);
        ctx_bound = true;
    }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return -4;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
    }
    if (((uintptr_t)base & 15) != 0) {
        set_error("TMA base address %p not 16-byte aligned", base);
        return -5;
    }
    for (int i = 0; i + 1 < rank; ++i) {
        if (gstr[i] % 16 != 0) {
            set_error("TMA global stride %llu (dim %d) not a multiple of 16 bytes",
                      (unsigned long long)gstr[i], i + 1);
            return -5;
        }
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                    gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                    : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                    : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
                  (int)r, rank, (unsigned long long)dims[0],
                  (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
        return -6;
    }
    return 0;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec {
    cudaEvent_t e0, e1;
    int cat;
    double work;
    int M, N, K;
};
static bool g_prof = false;
bool profiling_on() { return g_prof; }
static std::vector<ProfRec> g_recs;
static std::mutex g_prof_mu;

ProfScope::ProfScope(cudaStream_t s, int cat, double work, int M, int N, int K) : stream(s), slot(-1) {
    if (!g_prof) return;
    ProfRec r;
    r.cat = cat;
    r.work = work;
    r.M = M;
    r.N = N;
    r.K = K;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_recs.push_back(r);
    slot = (int)g_recs.size() - 1;
}
ProfScope::~ProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_recs[slot].e1, stream);
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace fd

extern "C" {

const char* fd_last_error(void) { return fd::g_err; }

int fd_version(void) { return 100; }

long long fd_launch_count(void) { return fd::g_launches.load(); }

void fd_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(fd::g_prof_mu);
    fd::g_prof = on != 0;
}

int fd_profile_dump(const char* path) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(fd::g_prof_mu);
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "cat,M,N,K,ms,flops\n");
    for (auto& r : fd::g_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess)
            fprintf(f, "%d,%d,%d,%d,%.6f,%.0f\n", r.cat, r.M, r.N, r.K, t, r.work);
    }
    fclose(f);
    return 0;
}

int fd_profile_summary(double* ms, double* work, long long* counts, int ncat) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(fd::g_prof_mu);
    for (int i = 0; i < ncat; ++i) {
        ms[i] = 0;
        work[i] = 0;
        counts[i] = 0;
    }
    for (auto& r : fd::g_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess && r.cat < ncat) {
            ms[r.cat] += t;
            work[r.cat] += r.work;
            counts[r.cat] += 1;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    fd::g_recs.clear();
    return 0;
}

int fd_sm_arch(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
        return -1;
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return major * 10 + minor;
}

}  // extern "C"
