// fd_attn.cu — fused FlashAttention-style forward for sm_100a (head dim 64, bf16, fp32 softmax).
//
//   O = softmax(Q K^T * scale) V       per (batch, head), Q/K/V read in place from [B, N, H*64]
//
// One CTA per (128-query tile, head, batch); 6 warps:
//   warp 0    TMA producer  : Q once, K/V tiles (128 keys) through a 2-stage ring
//   warp 1    MMA issuer    : S = Q K^T (tcgen05, 128x128x64) and O += P V (128x64x128), TMEM accum
//   warps 2-5 softmax       : tcgen05.ld S, online softmax (exp2, warp-free: one row per thread),
//                             P -> shared memory (bf16, 128B-swizzled K-major), O rescale in TMEM
// TMEM: S in columns [0,128), O in [128,192).  V is consumed as an MN-major B operand straight
// from its row-major [key, d] tile, so no transpose is materialised.
// Two CTAs fit per SM (112 KB smem, 256 TMEM columns each) and overlap each other's
// softmax / MMA phases.
//
// UPSTREAM math: diffusers Attention + AttnProcessor2_0 -> F.scaled_dot_product_attention
// (SURVEY.md §2.2); reference call path src/flash/models/unets/unet.py:108-119.
#include "fd_common.cuh"
#include "fd_host.h"

namespace fd {

constexpr int ATT_BM = 128;   // queries per CTA
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB
// No alignment slack: two CTAs must fit in 228 KB; the dynamic smem window is 1024-aligned
// (checked at kernel entry).
constexpr int ATT_SMEM = ATT_TILE_BYTES /*Q*/ + 4 * ATT_TILE_BYTES /*K,V x2*/ +
                         2 * ATT_TILE_BYTES /*P*/ + 256;

struct AttnKParams {
    int Nq, Nkv;
    float scale_log2;
    bf16* o;
    long long ldo, o_batch_stride;
    float* lse;  // [B,H,Nq] or null
    int H;
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
          "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
          "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
          "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
          "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnKParams p) {
    extern __shared__ uint8_t smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0) {
        if (threadIdx.x == 0) printf("fd_attn: dynamic smem base not 1024-aligned\n");
        __trap();
    }
    uint8_t* sQ = smem_raw;
    uint8_t* sK = sQ + ATT_TILE_BYTES;          // 2 stages
    uint8_t* sV = sK + 2 * ATT_TILE_BYTES;      // 2 stages
    uint8_t* sP = sV + 2 * ATT_TILE_BYTES;      // 2 sub-tiles of [128][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;   // [2]
    uint64_t* kv_empty = bars + 3;  // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_ready = bars + 6;
    uint64_t* o_done = bars + 7;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int n_kv_tiles = (p.Nkv + ATT_BN - 1) / ATT_BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_ready, 128);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_holder, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tmem_S = tmem_base;
    const uint32_t tmem_O = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(&tmQ, q_full, sQ, head * ATT_D, q_tile * ATT_BM, batch);
            for (int j = 0; j < n_kv_tiles; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&kv_empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
                tma_load_3d(&tmK, &kv_full[st], sK + st * ATT_TILE_BYTES, head * ATT_D, j * ATT_BN, batch);
                tma_load_3d(&tmV, &kv_full[st], sV + st * ATT_TILE_BYTES, head * ATT_D, j * ATT_BN, batch);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_BN, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, 0, 1);  // B (=V) MN-major
            mbar_wait(q_full, 0);
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t p_addr = smem_u32(sP);
            for (int j = 0; j < n_kv_tiles; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&kv_full[st], ph);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sK + st * ATT_TILE_BYTES);
                const uint32_t v_addr = smem_u32(sV + st * ATT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k)
                    tc_mma_bf16(tmem_S, make_desc_k_sw128(q_addr + k * 32), make_desc_k_sw128(k_addr + k * 32),
                                idesc_qk, k != 0 ? 1u : 0u);
                tc_commit(s_full);
                mbar_wait(p_ready, j & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < ATT_BN / 16; ++k)
                    tc_mma_bf16(tmem_O, make_desc_k_sw128(p_addr + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32),
                                make_desc_mn_sw128(v_addr + k * 2048, 0, 1024), idesc_pv,
                                (j | k) != 0 ? 1u : 0u);
                tc_commit(&kv_empty[st]);
                tc_commit(o_done);
            }
        }
    } else {
        // softmax warps 2..5: TMEM lane quarter = warp % 4
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;              // row within the query tile
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_kv_tiles; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int kv_valid = min(ATT_BN, p.Nkv - j * ATT_BN);
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < ATT_BN / 32; ++c) {
                if (c * 32 >= kv_valid) break;
                uint32_t r[32];
                tmem_ld_32x32(tmem_S + lane_base + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float alpha = exp2f(m_run - m_new);
            // previous P V must be complete before O / P are touched again
            if (j > 0) {
                mbar_wait(o_done, (j - 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < ATT_D / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_O + lane_base + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                    tmem_st_32x32(tmem_O + lane_base + c * 32, r);
                }
                tmem_st_wait();
            }
            // pass 2: exponentials -> P (bf16) in swizzled K-major shared memory
            float psum = 0.f;
#pragma unroll 1
            for (int c = 0; c < ATT_BN / 32; ++c) {
                uint32_t r[32];
                uint32_t pk[16];
                if (c * 32 < kv_valid) {
                    tmem_ld_32x32(tmem_S + lane_base + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float p0 = (c * 32 + i < kv_valid) ? exp2f(__uint_as_float(r[i]) * p.scale_log2 - m_new) : 0.f;
                        float p1 = (c * 32 + i + 1 < kv_valid) ? exp2f(__uint_as_float(r[i + 1]) * p.scale_log2 - m_new) : 0.f;
                        psum += p0 + p1;
                        pk[i >> 1] = pack_bf16x2(p0, p1);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = 0u;
                }
                uint8_t* sub = sP + (c >> 1) * ATT_TILE_BYTES + row * 128;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int chunk = ((c & 1) * 4 + q4) ^ (row & 7);
                    *reinterpret_cast<uint4*>(sub + chunk * 16) =
                        make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
                }
            }
            l_run = l_run * alpha + psum;
            m_run = m_new;
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(p_ready);
        }
        // epilogue: O / l -> bf16
        mbar_wait(o_done, (n_kv_tiles - 1) & 1);
        tc_fence_after();
        const int q_row = q_tile * ATT_BM + row;
        const float inv_l = 1.f / l_run;
        bf16* orow = p.o + (long long)batch * p.o_batch_stride + (long long)q_row * p.ldo + head * ATT_D;
#pragma unroll 1
        for (int c = 0; c < ATT_D / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_O + lane_base + c * 32, r);
            tmem_ld_wait();
            if (q_row < p.Nq) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(r[8 * q4 + 0]) * inv_l, __uint_as_float(r[8 * q4 + 1]) * inv_l);
                    u.y = pack_bf16x2(__uint_as_float(r[8 * q4 + 2]) * inv_l, __uint_as_float(r[8 * q4 + 3]) * inv_l);
                    u.z = pack_bf16x2(__uint_as_float(r[8 * q4 + 4]) * inv_l, __uint_as_float(r[8 * q4 + 5]) * inv_l);
                    u.w = pack_bf16x2(__uint_as_float(r[8 * q4 + 6]) * inv_l, __uint_as_float(r[8 * q4 + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + c * 32 + q4 * 8) = u;
                }
            }
        }
        if (p.lse != nullptr && q_row < p.Nq)
            p.lse[((long long)batch * p.H + head) * p.Nq + q_row] = (m_run + log2f(l_run)) * 0.69314718055994531f;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

static int make_qkv_tmap(CUtensorMap* m, const void* base, int H, int N, int B, int64_t ld,
                         int64_t batch_stride) {
    const uint64_t dims[3] = {(uint64_t)H * ATT_D, (uint64_t)N, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)batch_stride * 2};
    const uint32_t box[3] = {(uint32_t)ATT_D, 128u, 1u};
    return encode_tmap_bf16(m, base, 3, dims, str, box);
}

}  // namespace fd

using namespace fd;

extern "C" int fd_attn_fwd(const FdAttnArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(a && a->q && a->k && a->v && a->o, "fd_attn_fwd: null tensor");
    FD_CHECK_ARG(a->B > 0 && a->H > 0 && a->Nq > 0 && a->Nkv > 0, "fd_attn_fwd: bad sizes");
    FD_CHECK_ARG(a->ldo % 8 == 0 && a->o_batch_stride % 8 == 0, "fd_attn_fwd: o strides must be multiples of 8");
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = make_qkv_tmap(&tq, a->q, a->H, a->Nq, a->B, a->ldq, a->q_batch_stride))) return rc;
    if ((rc = make_qkv_tmap(&tk, a->k, a->H, a->Nkv, a->B, a->ldk, a->k_batch_stride))) return rc;
    if ((rc = make_qkv_tmap(&tv, a->v, a->H, a->Nkv, a->B, a->ldv, a->v_batch_stride))) return rc;
    AttnKParams p;
    p.Nq = a->Nq;
    p.Nkv = a->Nkv;
    p.scale_log2 = a->scale * 1.4426950408889634f;
    p.o = (bf16*)a->o;
    p.ldo = a->ldo;
    p.o_batch_stride = a->o_batch_stride;
    p.lse = a->lse;
    p.H = a->H;
    static bool attr_set = false;
    if (!attr_set) {
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        attr_set = true;
    }
    dim3 grid((a->Nq + ATT_BM - 1) / ATT_BM, a->H, a->B);
    ProfScope prof(stream, PROF_ATTN_FWD, 4.0 * (double)a->B * a->H * (double)a->Nq * (double)a->Nkv * ATT_D);
    attn_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tq, tk, tv, p);
    FD_CHECK_LAUNCH();
    return 0;
}
