// fd_attn_generic.cu — FlashAttention forward for head dims other than 64 (any multiple of 16 up to 192) and for
// key-padding masks: SD1.5 (d = 40/80/160, examples/train_flash_sd.py:76,89), PixArt-alpha (d = 72, masked T5 context,
// examples/train_flash_pixart.py:68,73).  Head dims that are not multiples of 16 are zero-padded by the weight packs
// (zero q/k/v rows, zero out-proj columns), which leaves softmax(QK^T)V unchanged.
//
// One CTA per (128-query tile, head, batch), 6 warps (TMA, MMA, 4 softmax).  Q/K/V tiles are NS sub-tiles of
// [128 rows][64 columns] (128B swizzle); S = Q K^T runs d/16 k-steps, O += P V uses an MN-major V operand of N = d
// columns; S [0,128) and O [128,128+d) live in TMEM.  Same math and masking rules as fd_attn.cu; tuned for
// correctness and generality, not peak (the d = 64 fast path is fd_attn.cu).
#include "fd_common.cuh"
#include "fd_host.h"

namespace fd {

constexpr int AG_TILE = 128 * 64 * 2;  // one [128][64] bf16 sub-tile

struct AttnGParams {
    int Nq, Nkv, d, H;
    float scale_log2;
    bf16* o;
    long long ldo, o_batch_stride;
    float* lse;
    const int* kv_len;   // optional [B]: keys >= kv_len[b] are masked
};

__device__ __forceinline__ void ag_tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) { tmem_st_32x16(taddr, r); }

template <int NS, int STG>
__global__ void __launch_bounds__(192, 1)
attn_fwd_generic_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const AttnGParams p) {
    constexpr int T = NS * AG_TILE;  // bytes of one Q / K / V tile
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + T;
    uint8_t* sV = sK + STG * T;
    uint8_t* sP = sV + STG * T;                   // 2 sub-tiles [128][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AG_TILE);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;        // [STG]
    uint64_t* kv_empty = kv_full + STG;  // [STG]
    uint64_t* s_full = kv_empty + STG;
    uint64_t* p_ready = s_full + 1;
    uint64_t* o_done = p_ready + 1;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(o_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int nkv = p.kv_len != nullptr ? min(p.Nkv, max(1, p.kv_len[batch])) : p.Nkv;
    const int n_kv_tiles = (nkv + 127) / 128;
    const int d = p.d;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < STG; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_ready, 128);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_holder, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, T);
            for (int s = 0; s < NS; ++s)
                tma_load_3d(&tmQ, q_full, sQ + s * AG_TILE, head * d + s * 64, q_tile * 128, batch);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < n_kv_tiles; ++j) {
                mbar_wait(&kv_empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&kv_full[st], 2 * T);
                for (int s = 0; s < NS; ++s) {
                    tma_load_3d(&tmK, &kv_full[st], sK + st * T + s * AG_TILE, head * d + s * 64, j * 128, batch);
                    tma_load_3d(&tmV, &kv_full[st], sV + st * T + s * AG_TILE, head * d + s * 64, j * 128, batch);
                }
                if (++st == STG) {
                    st = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
            const uint32_t idesc_pv = make_idesc_bf16(128, d, 0, 1);
            mbar_wait(q_full, 0);
            const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < n_kv_tiles; ++j) {
                mbar_wait(&kv_full[st], ph);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sK + st * T), v_addr = smem_u32(sV + st * T);
                for (int k = 0; k < d / 16; ++k)
                    tc_mma_bf16(tmem_S, make_desc_k_sw128(q_addr + (k >> 2) * AG_TILE + (k & 3) * 32),
                                make_desc_k_sw128(k_addr + (k >> 2) * AG_TILE + (k & 3) * 32), idesc_qk,
                                k != 0 ? 1u : 0u);
                tc_commit(s_full);
                mbar_wait(p_ready, j & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_bf16(tmem_O, make_desc_k_sw128(p_addr + (k >> 2) * AG_TILE + (k & 3) * 32),
                                make_desc_mn_sw128(v_addr + k * 2048, AG_TILE, 1024), idesc_pv, (j | k) != 0 ? 1u : 0u);
                tc_commit(&kv_empty[st]);
                tc_commit(o_done);
                if (++st == STG) {
                    st = 0;
                    ph ^= 1u;
                }
            }
        }
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_kv_tiles; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int kv_valid = min(128, nkv - j * 128);
            uint32_t sr[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_S + lane_base + c * 32, sr[c]);
            tmem_ld_wait();
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (c * 32 + i >= kv_valid) sr[c][i] = 0xff800000u;
                    mx = fmaxf(mx, __uint_as_float(sr[c][i]));
                }
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float alpha = exp2f(m_run - m_new);
            if (j > 0) {
                mbar_wait(o_done, (j - 1) & 1);
                tc_fence_after();
                for (int c = 0; c < d / 16; ++c) {
                    uint32_t r[16];
                    tmem_ld_32x16(tmem_O + lane_base + c * 16, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                    ag_tmem_st_32x16(tmem_O + lane_base + c * 16, r);
                }
                tmem_st_wait();
            }
            float psum = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float p0 = exp2f(__uint_as_float(sr[c][i]) * p.scale_log2 - m_new);
                    const float p1 = exp2f(__uint_as_float(sr[c][i + 1]) * p.scale_log2 - m_new);
                    psum += p0 + p1;
                    pk[i >> 1] = pack_bf16x2(p0, p1);
                }
                uint8_t* sub = sP + (c >> 1) * AG_TILE + row * 128;
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
        mbar_wait(o_done, (n_kv_tiles - 1) & 1);
        tc_fence_after();
        const int q_row = q_tile * 128 + row;
        const float inv_l = 1.f / l_run;
        bf16* orow = p.o + (long long)batch * p.o_batch_stride + (long long)q_row * p.ldo + head * d;
        for (int c = 0; c < d / 16; ++c) {
            uint32_t r[16];
            tmem_ld_32x16(tmem_O + lane_base + c * 16, r);
            tmem_ld_wait();
            if (q_row < p.Nq) {
#pragma unroll
                for (int q4 = 0; q4 < 2; ++q4) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(r[8 * q4 + 0]) * inv_l, __uint_as_float(r[8 * q4 + 1]) * inv_l);
                    u.y = pack_bf16x2(__uint_as_float(r[8 * q4 + 2]) * inv_l, __uint_as_float(r[8 * q4 + 3]) * inv_l);
                    u.z = pack_bf16x2(__uint_as_float(r[8 * q4 + 4]) * inv_l, __uint_as_float(r[8 * q4 + 5]) * inv_l);
                    u.w = pack_bf16x2(__uint_as_float(r[8 * q4 + 6]) * inv_l, __uint_as_float(r[8 * q4 + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + c * 16 + q4 * 8) = u;
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
        tmem_dealloc(tmem_base, 512);
    }
}

static int ag_tmap(CUtensorMap* m, const void* base, int HD, int N, int B, int64_t ld, int64_t bs) {
    const uint64_t dims[3] = {(uint64_t)HD, (uint64_t)N, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)bs * 2};
    const uint32_t box[3] = {64u, 128u, 1u};
    return encode_tmap_bf16(m, base, 3, dims, str, box);
}

template <int NS, int STG>
static int ag_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnGParams& p,
                     dim3 grid, cudaStream_t stream) {
    constexpr int SMEM = NS * AG_TILE * (1 + 2 * STG) + 2 * AG_TILE + 256 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_generic_kernel<NS, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    attn_fwd_generic_kernel<NS, STG><<<grid, 192, SMEM, stream>>>(tq, tk, tv, p);
    FD_CHECK_LAUNCH();
    return 0;
}

}  // namespace fd

using namespace fd;

extern "C" int fd_attn_fwd_generic(const FdAttnArgs* a, int32_t head_dim, const int32_t* kv_len, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(a && a->q && a->k && a->v && a->o, "fd_attn_fwd_generic: null tensor");
    FD_CHECK_ARG(head_dim % 16 == 0 && head_dim >= 16 && head_dim <= 192,
                 "fd_attn_fwd_generic: head_dim=%d must be a multiple of 16 in [16,192]", head_dim);
    FD_CHECK_ARG(a->B > 0 && a->H > 0 && a->Nq > 0 && a->Nkv > 0, "fd_attn_fwd_generic: bad sizes");
    FD_CHECK_ARG(a->ldo % 8 == 0 && a->o_batch_stride % 8 == 0, "fd_attn_fwd_generic: o strides must be multiples of 8");
    const int HD = a->H * head_dim;
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = ag_tmap(&tq, a->q, HD, a->Nq, a->B, a->ldq, a->q_batch_stride))) return rc;
    if ((rc = ag_tmap(&tk, a->k, HD, a->Nkv, a->B, a->ldk, a->k_batch_stride))) return rc;
    if ((rc = ag_tmap(&tv, a->v, HD, a->Nkv, a->B, a->ldv, a->v_batch_stride))) return rc;
    AttnGParams p;
    p.Nq = a->Nq; p.Nkv = a->Nkv; p.d = head_dim; p.H = a->H;
    p.scale_log2 = a->scale * 1.4426950408889634f;
    p.o = (bf16*)a->o; p.ldo = a->ldo; p.o_batch_stride = a->o_batch_stride;
    p.lse = a->lse;
    p.kv_len = kv_len;
    dim3 grid((a->Nq + 127) / 128, a->H, a->B);
    if (head_dim <= 64) return ag_launch<1, 2>(tq, tk, tv, p, grid, stream);
    if (head_dim <= 128) return ag_launch<2, 2>(tq, tk, tv, p, grid, stream);
    return ag_launch<3, 1>(tq, tk, tv, p, grid, stream);
}
