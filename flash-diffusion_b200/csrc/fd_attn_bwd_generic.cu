// fd_attn_bwd_generic.cu — attention backward for head dims other than 64 and for key-padding masks (sm_100a).
//
// Two kernels behind one entry point (fd_attn_bwd_generic):
//  * attn_bwd_g_kernel<NS, STG>: the tcgen05 backward of fd_attn_bwd.cu generalised to a head dim d that is a
//    multiple of 16 up to 80 (SD1.5 d=40 -> 48 / 80, PixArt-alpha d=72 -> 80) and to kv_len[B] masks.  Q, K, V, dO tiles
//    are NS sub-tiles of [128][64] (128-byte swizzle); the contraction over d issues d/16 MMAs (the columns a 64-wide
//    TMA box reads past d belong to the next head and are never multiplied), the products whose N dimension is d use
//    instruction N = d over MN-major descriptors that span the sub-tiles.  TMEM: S [0,128), dP [128,256),
//    dV [256,256+d), dK [384,384+d); dQ re-uses the S columns [0,d) once the softmax warps have consumed S.
//  * attn_bwd_small_*: three CUDA-core passes (P / dS into a global scratch, then dK|dV and dQ) for larger head dims at
//    short sequences — SD1.5's d = 160 levels: 16x16 and 8x8 tokens, 77 text keys, ~1% of that UNet's attention FLOPs.
//
// UPSTREAM math: autograd of F.scaled_dot_product_attention(q, k, v, attn_mask=key padding) under the student LoRA
// backward (reference src/flash/models/unets/unet.py:108-119, transformers/tranformers.py:58-92).
#include "fd_common.cuh"
#include "fd_host.h"

namespace fd {

constexpr int GB_T = 128;
constexpr int GB_TILE = GB_T * 64 * 2;   // one [128][64] bf16 sub-tile, 16 KB
constexpr int GB_THREADS = 448;   // warps: 0 TMA, 1 MMA, 2-9 softmax / dS (two per TMEM lane quarter), 10-13 dQ drain

struct AttnBwdGParams {
    int Nq, Nkv, H, d;
    float scale, scale_log2;
    const float* lse;      // [B,H,Nq]
    const float* delta;    // [B,H,Nq]
    const int* kv_len;     // [B] or null
    float* dq_accum;       // [B,Nq,H*d] fp32
    bf16* dk; long long lddk, dk_bs;
    bf16* dv; long long lddv, dv_bs;
};

__device__ __forceinline__ float gb_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void attn_delta_g_kernel(const bf16* __restrict__ o, long long ldo, long long o_bs,
                                    const bf16* __restrict__ d_o, long long lddo, long long do_bs,
                                    float* __restrict__ delta, int B, int H, int Nq, int d) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * H * Nq;
    if (idx >= total) return;
    const int row = (int)(idx % Nq);
    const int h = (int)((idx / Nq) % H);
    const int b = (int)(idx / ((long long)Nq * H));
    const uint4* po = reinterpret_cast<const uint4*>(o + (long long)b * o_bs + (long long)row * ldo + h * d);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + (long long)b * do_bs + (long long)row * lddo + h * d);
    float acc = 0.f;
    for (int i = 0; i < d / 8; ++i) {
        const uint4 a = po[i], c = pd[i];
        float2 x, y;
        x = unpack_bf16x2(a.x); y = unpack_bf16x2(c.x); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16x2(a.y); y = unpack_bf16x2(c.y); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16x2(a.z); y = unpack_bf16x2(c.z); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16x2(a.w); y = unpack_bf16x2(c.w); acc += x.x * y.x + x.y * y.y;
    }
    delta[idx] = acc;
}

__global__ void attn_dq_convert_g_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long lddq,
                                         long long dq_bs, int Nq, int HD, long long total4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const long long e = i * 4;
    const int c = (int)(e % HD);
    const long long r = e / HD;
    const int row = (int)(r % Nq);
    const long long b = r / Nq;
    const float4 v = *reinterpret_cast<const float4*>(acc + e);
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dq + b * dq_bs + (long long)row * lddq + c) = u;
}

template <int NS, int STG>
__global__ void __launch_bounds__(GB_THREADS, 1)
attn_bwd_g_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                  const AttnBwdGParams p) {
    constexpr int T = NS * GB_TILE;   // bytes of one Q / K / V / dO tile
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sK = smem;
    uint8_t* sV = sK + T;
    uint8_t* sQ = sV + T;                 // STG stages
    uint8_t* sDO = sQ + STG * T;          // STG stages
    uint8_t* sP = sDO + STG * T;          // [q][kv] as 2 sub-tiles of [128][64]
    uint8_t* sDS = sP + 2 * GB_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 2 * GB_TILE);
    uint64_t* kv_full = bars;
    uint64_t* qdo_full = bars + 1;    // [2]
    uint64_t* qdo_empty = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* pds_ready = bars + 6;
    uint64_t* mma2_done = bars + 7;
    uint64_t* dq_full = bars + 8;
    uint64_t* dq_empty = bars + 9;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kv_tile = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int n_q_tiles = (p.Nq + GB_T - 1) / GB_T;
    const int d = p.d;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmDO);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(kv_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&qdo_full[s], 1);
            mbar_init(&qdo_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(pds_ready, 256);
        mbar_init(mma2_done, 1);
        mbar_init(dq_full, 1);
        mbar_init(dq_empty, 128);
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
    const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dV = tmem_base + 256,
                   tmem_dK = tmem_base + 384, tmem_dQ = tmem_base;   // dQ aliases S

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * T);
            for (int s = 0; s < NS; ++s) {
                tma_load_3d(&tmK, kv_full, sK + s * GB_TILE, head * d + s * 64, kv_tile * GB_T, batch);
                tma_load_3d(&tmV, kv_full, sV + s * GB_TILE, head * d + s * 64, kv_tile * GB_T, batch);
            }
            int st = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_q_tiles; ++i) {
                mbar_wait(&qdo_empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&qdo_full[st], 2 * T);
                for (int s = 0; s < NS; ++s) {
                    tma_load_3d(&tmQ, &qdo_full[st], sQ + st * T + s * GB_TILE, head * d + s * 64, i * GB_T, batch);
                    tma_load_3d(&tmDO, &qdo_full[st], sDO + st * T + s * GB_TILE, head * d + s * 64, i * GB_T, batch);
                }
                if (++st == STG) {
                    st = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);    // Q K^T, dO V^T
            const uint32_t id_t = make_idesc_bf16(128, d, 1, 1);      // P^T dO, dS^T Q (both MN-major)
            const uint32_t id_q = make_idesc_bf16(128, d, 0, 1);      // dS K  (A K-major, B MN-major)
            mbar_wait(kv_full, 0);
            const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
            const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
            int st = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_q_tiles; ++i) {
                mbar_wait(&qdo_full[st], ph);
                if (i > 0) mbar_wait(dq_empty, (i - 1) & 1);      // dQ(i-1) drained: the S columns are free again
                tc_fence_after();
                const uint32_t q_addr = smem_u32(sQ + st * T);
                const uint32_t do_addr = smem_u32(sDO + st * T);
                for (int k = 0; k < d / 16; ++k)
                    tc_mma_bf16(tmem_S, make_desc_k_sw128(q_addr + (k >> 2) * GB_TILE + (k & 3) * 32),
                                make_desc_k_sw128(k_addr + (k >> 2) * GB_TILE + (k & 3) * 32), id_s, k != 0 ? 1u : 0u);
                for (int k = 0; k < d / 16; ++k)
                    tc_mma_bf16(tmem_dP, make_desc_k_sw128(do_addr + (k >> 2) * GB_TILE + (k & 3) * 32),
                                make_desc_k_sw128(v_addr + (k >> 2) * GB_TILE + (k & 3) * 32), id_s, k != 0 ? 1u : 0u);
                tc_commit(s_full);
                mbar_wait(pds_ready, i & 1);
                tc_fence_after();
                // dV += P^T dO ; dK += dS^T Q      (contraction over the 128 query rows, 16 per MMA)
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_bf16(tmem_dV, make_desc_mn_sw128(p_addr + k * 2048, GB_TILE, 1024),
                                make_desc_mn_sw128(do_addr + k * 2048, GB_TILE, 1024), id_t, (i | k) != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_bf16(tmem_dK, make_desc_mn_sw128(ds_addr + k * 2048, GB_TILE, 1024),
                                make_desc_mn_sw128(q_addr + k * 2048, GB_TILE, 1024), id_t, (i | k) != 0 ? 1u : 0u);
                // dQ_i = dS K                    (contraction over the 128 keys) into the consumed S columns
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_bf16(tmem_dQ, make_desc_k_sw128(ds_addr + (k >> 2) * GB_TILE + (k & 3) * 32),
                                make_desc_mn_sw128(k_addr + k * 2048, GB_TILE, 1024), id_q, k != 0 ? 1u : 0u);
                tc_commit(&qdo_empty[st]);
                tc_commit(mma2_done);
                tc_commit(dq_full);
                if (++st == STG) {
                    st = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp < 10) {
        // softmax / dS warps: thread (row, half) owns 64 keys of one query row
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const int nkv = p.kv_len != nullptr ? min(p.Nkv, max(1, p.kv_len[batch])) : p.Nkv;
        const int kv_valid = max(0, min(GB_T, nkv - kv_tile * GB_T));
        const float* lse_bh = p.lse + ((long long)batch * p.H + head) * p.Nq;
        const float* delta_bh = p.delta + ((long long)batch * p.H + head) * p.Nq;
        for (int i = 0; i < n_q_tiles; ++i) {
            const int q_row = i * GB_T + row;
            const bool q_ok = q_row < p.Nq;
            const float lse2 = q_ok ? lse_bh[q_row] * 1.4426950408889634f : 0.f;
            const float dlt = q_ok ? delta_bh[q_row] : 0.f;
            const bool full = (i + 1) * GB_T <= p.Nq && kv_valid == GB_T;      // block-uniform: no masking needed
            mbar_wait(s_full, i & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c = half * 2 + cc;
                uint32_t rs[32], rp[32];
                tmem_ld_32x32(tmem_S + lane_base + c * 32, rs);
                tmem_ld_32x32(tmem_dP + lane_base + c * 32, rp);
                tmem_ld_wait();
                uint32_t pk[16], dk_[16];
                if (full) {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float p0 = gb_ex2(fmaf(__uint_as_float(rs[j]), p.scale_log2, -lse2));
                        const float p1 = gb_ex2(fmaf(__uint_as_float(rs[j + 1]), p.scale_log2, -lse2));
                        const float d0 = p0 * (__uint_as_float(rp[j]) - dlt) * p.scale;
                        const float d1 = p1 * (__uint_as_float(rp[j + 1]) - dlt) * p.scale;
                        pk[j >> 1] = pack_bf16x2(p0, p1);
                        dk_[j >> 1] = pack_bf16x2(d0, d1);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
                        if (q_ok && c * 32 + j < kv_valid) {
                            p0 = gb_ex2(fmaf(__uint_as_float(rs[j]), p.scale_log2, -lse2));
                            d0 = p0 * (__uint_as_float(rp[j]) - dlt) * p.scale;
                        }
                        if (q_ok && c * 32 + j + 1 < kv_valid) {
                            p1 = gb_ex2(fmaf(__uint_as_float(rs[j + 1]), p.scale_log2, -lse2));
                            d1 = p1 * (__uint_as_float(rp[j + 1]) - dlt) * p.scale;
                        }
                        pk[j >> 1] = pack_bf16x2(p0, p1);
                        dk_[j >> 1] = pack_bf16x2(d0, d1);
                    }
                }
                if (cc == 0 && i > 0) mbar_wait(mma2_done, (i - 1) & 1);   // P / dS buffers free again
                uint8_t* subp = sP + half * GB_TILE + row * 128;
                uint8_t* subd = sDS + half * GB_TILE + row * 128;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int chunk = (cc * 4 + q4) ^ (row & 7);
                    *reinterpret_cast<uint4*>(subp + chunk * 16) =
                        make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
                    *reinterpret_cast<uint4*>(subd + chunk * 16) =
                        make_uint4(dk_[4 * q4], dk_[4 * q4 + 1], dk_[4 * q4 + 2], dk_[4 * q4 + 3]);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(pds_ready);
        }
        // epilogue: dK, dV of this key tile
        mbar_wait(mma2_done, (n_q_tiles - 1) & 1);
        tc_fence_after();
        const int kv_row = kv_tile * GB_T + row;
        bf16* dk_row = p.dk + (long long)batch * p.dk_bs + (long long)kv_row * p.lddk + head * d;
        bf16* dv_row = p.dv + (long long)batch * p.dv_bs + (long long)kv_row * p.lddv + head * d;
        const int nch = d / 16, split = (nch + 1) / 2;
#pragma unroll 1
        for (int c = half ? split : 0; c < (half ? nch : split); ++c) {
            uint32_t rk[16], rv[16];
            tmem_ld_32x16(tmem_dK + lane_base + c * 16, rk);
            tmem_ld_32x16(tmem_dV + lane_base + c * 16, rv);
            tmem_ld_wait();
            if (kv_row < p.Nkv) {
#pragma unroll
                for (int q4 = 0; q4 < 2; ++q4) {
                    uint4 u, w;
                    u.x = pack_bf16x2(__uint_as_float(rk[8 * q4 + 0]), __uint_as_float(rk[8 * q4 + 1]));
                    u.y = pack_bf16x2(__uint_as_float(rk[8 * q4 + 2]), __uint_as_float(rk[8 * q4 + 3]));
                    u.z = pack_bf16x2(__uint_as_float(rk[8 * q4 + 4]), __uint_as_float(rk[8 * q4 + 5]));
                    u.w = pack_bf16x2(__uint_as_float(rk[8 * q4 + 6]), __uint_as_float(rk[8 * q4 + 7]));
                    w.x = pack_bf16x2(__uint_as_float(rv[8 * q4 + 0]), __uint_as_float(rv[8 * q4 + 1]));
                    w.y = pack_bf16x2(__uint_as_float(rv[8 * q4 + 2]), __uint_as_float(rv[8 * q4 + 3]));
                    w.z = pack_bf16x2(__uint_as_float(rv[8 * q4 + 4]), __uint_as_float(rv[8 * q4 + 5]));
                    w.w = pack_bf16x2(__uint_as_float(rv[8 * q4 + 6]), __uint_as_float(rv[8 * q4 + 7]));
                    *reinterpret_cast<uint4*>(dk_row + c * 16 + q4 * 8) = u;
                    *reinterpret_cast<uint4*>(dv_row + c * 16 + q4 * 8) = w;
                }
            }
        }
    } else {
        // dQ drain warps 10..13: TMEM -> fp32 atomics
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        for (int i = 0; i < n_q_tiles; ++i) {
            mbar_wait(dq_full, i & 1);
            tc_fence_after();
            const int q_row = i * GB_T + row;
            float* dst = p.dq_accum + ((long long)batch * p.Nq + q_row) * ((long long)p.H * d) + head * d;
#pragma unroll 1
            for (int c = 0; c < d / 16; ++c) {
                uint32_t r[16];
                tmem_ld_32x16(tmem_dQ + lane_base + c * 16, r);
                tmem_ld_wait();
                if (q_row < p.Nq) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        atomicAdd(reinterpret_cast<float4*>(dst + c * 16 + j),
                                  make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                              __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
                }
            }
            tc_fence_before();
            mbar_arrive(dq_empty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------- short sequences
struct AttnBwdSParams {
    int Nq, Nkv, H, d;
    float scale;
    const bf16 *q, *k, *v, *d_o;
    long long ldq, q_bs, ldk, k_bs, ldv, v_bs, lddo, do_bs;
    const float* lse;
    const float* delta;
    const int* kv_len;
    bf16 *dq, *dk, *dv;
    long long lddq, dq_bs, lddk, dk_bs, lddv, dv_bs;
};

__device__ __forceinline__ float dot_bf16(const bf16* a, const bf16* b, int d) {
    float acc = 0.f;
    for (int c = 0; c < d; c += 8) {
        const uint4 x = *reinterpret_cast<const uint4*>(a + c), y = *reinterpret_cast<const uint4*>(b + c);
        float2 u, w;
        u = unpack_bf16x2(x.x); w = unpack_bf16x2(y.x); acc += u.x * w.x + u.y * w.y;
        u = unpack_bf16x2(x.y); w = unpack_bf16x2(y.y); acc += u.x * w.x + u.y * w.y;
        u = unpack_bf16x2(x.z); w = unpack_bf16x2(y.z); acc += u.x * w.x + u.y * w.y;
        u = unpack_bf16x2(x.w); w = unpack_bf16x2(y.w); acc += u.x * w.x + u.y * w.y;
    }
    return acc;
}

// pass 1: P and dS [B,H,Nq,Nkv] fp32 into global scratch, one (query, key) pair per thread
__global__ void __launch_bounds__(256)
attn_bwd_small_p_kernel(const AttnBwdSParams p, float* __restrict__ gP, float* __restrict__ gDS) {
    const int head = blockIdx.y, batch = blockIdx.z;
    const int Nq = p.Nq, Nkv = p.Nkv, d = p.d;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)Nq * Nkv) return;
    const int i = (int)(e / Nkv), j = (int)(e - (long long)i * Nkv);
    const bf16* q = p.q + (long long)batch * p.q_bs + head * d;
    const bf16* k = p.k + (long long)batch * p.k_bs + head * d;
    const bf16* v = p.v + (long long)batch * p.v_bs + head * d;
    const bf16* dO = p.d_o + (long long)batch * p.do_bs + head * d;
    const long long bh = (long long)batch * p.H + head;
    const int nkv = p.kv_len != nullptr ? min(Nkv, max(1, p.kv_len[batch])) : Nkv;
    float pr = 0.f, ds = 0.f;
    if (j < nkv) {
        const float s = dot_bf16(q + (long long)i * p.ldq, k + (long long)j * p.ldk, d) * p.scale;
        pr = __expf(s - p.lse[bh * Nq + i]);
        const float dp = dot_bf16(dO + (long long)i * p.lddo, v + (long long)j * p.ldv, d);
        ds = pr * (dp - p.delta[bh * Nq + i]) * p.scale;
    }
    gP[bh * Nq * Nkv + e] = pr;
    gDS[bh * Nq * Nkv + e] = ds;
}

// pass 2: dV[j, c] = sum_i P[i, j] dO[i, c] ; dK[j, c] = sum_i dS[i, j] Q[i, c]   (8 channels per thread)
__global__ void __launch_bounds__(128)
attn_bwd_small_kv_kernel(const AttnBwdSParams p, const float* __restrict__ gP, const float* __restrict__ gDS) {
    const int head = blockIdx.y, batch = blockIdx.z;
    const int Nq = p.Nq, Nkv = p.Nkv, d = p.d, dv8 = d / 8;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Nkv * dv8) return;
    const int j = e / dv8, c = (e - j * dv8) * 8;
    const bf16* q = p.q + (long long)batch * p.q_bs + head * d;
    const bf16* dO = p.d_o + (long long)batch * p.do_bs + head * d;
    const long long bh = (long long)batch * p.H + head;
    const float* P = gP + bh * Nq * Nkv;
    const float* DS = gDS + bh * Nq * Nkv;
    float av[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ak[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < Nq; ++i) {
        const float pr = P[(long long)i * Nkv + j], ds = DS[(long long)i * Nkv + j];
        const uint4 x = *reinterpret_cast<const uint4*>(dO + (long long)i * p.lddo + c);
        const uint4 y = *reinterpret_cast<const uint4*>(q + (long long)i * p.ldq + c);
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 a = unpack_bf16x2(xs[t]), b = unpack_bf16x2(ys[t]);
            av[2 * t] += pr * a.x; av[2 * t + 1] += pr * a.y;
            ak[2 * t] += ds * b.x; ak[2 * t + 1] += ds * b.y;
        }
    }
    uint4 u, w;
    u.x = pack_bf16x2(av[0], av[1]); u.y = pack_bf16x2(av[2], av[3]);
    u.z = pack_bf16x2(av[4], av[5]); u.w = pack_bf16x2(av[6], av[7]);
    w.x = pack_bf16x2(ak[0], ak[1]); w.y = pack_bf16x2(ak[2], ak[3]);
    w.z = pack_bf16x2(ak[4], ak[5]); w.w = pack_bf16x2(ak[6], ak[7]);
    *reinterpret_cast<uint4*>(p.dv + (long long)batch * p.dv_bs + (long long)j * p.lddv + head * d + c) = u;
    *reinterpret_cast<uint4*>(p.dk + (long long)batch * p.dk_bs + (long long)j * p.lddk + head * d + c) = w;
}

// pass 3: dQ[i, c] = sum_j dS[i, j] K[j, c]
__global__ void __launch_bounds__(128)
attn_bwd_small_q_kernel(const AttnBwdSParams p, const float* __restrict__ gDS) {
    const int head = blockIdx.y, batch = blockIdx.z;
    const int Nq = p.Nq, Nkv = p.Nkv, d = p.d, dv8 = d / 8;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Nq * dv8) return;
    const int i = e / dv8, c = (e - i * dv8) * 8;
    const bf16* k = p.k + (long long)batch * p.k_bs + head * d;
    const long long bh = (long long)batch * p.H + head;
    const float* DS = gDS + bh * Nq * Nkv + (long long)i * Nkv;
    float aq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < Nkv; ++j) {
        const float ds = DS[j];
        const uint4 y = *reinterpret_cast<const uint4*>(k + (long long)j * p.ldk + c);
        const uint32_t ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float2 b = unpack_bf16x2(ys[t]);
            aq[2 * t] += ds * b.x; aq[2 * t + 1] += ds * b.y;
        }
    }
    uint4 u;
    u.x = pack_bf16x2(aq[0], aq[1]); u.y = pack_bf16x2(aq[2], aq[3]);
    u.z = pack_bf16x2(aq[4], aq[5]); u.w = pack_bf16x2(aq[6], aq[7]);
    *reinterpret_cast<uint4*>(p.dq + (long long)batch * p.dq_bs + (long long)i * p.lddq + head * d + c) = u;
}

static int gb_tmap(CUtensorMap* m, const void* base, int HD, int N, int B, int64_t ld, int64_t bs) {
    const uint64_t dims[3] = {(uint64_t)HD, (uint64_t)N, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)bs * 2};
    const uint32_t box[3] = {64u, 128u, 1u};
    return encode_tmap_bf16(m, base, 3, dims, str, box);
}

template <int NS, int STG>
static int gb_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                     const AttnBwdGParams& p, dim3 grid, cudaStream_t stream) {
    constexpr int SMEM = NS * GB_TILE * (2 + 2 * STG) + 4 * GB_TILE + 256 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_g_kernel<NS, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    attn_bwd_g_kernel<NS, STG><<<grid, GB_THREADS, SMEM, stream>>>(tq, tk, tv, tdo, p);
    FD_CHECK_LAUNCH();
    return 0;
}

}  // namespace fd

using namespace fd;

extern "C" int fd_attn_bwd_generic(const FdAttnBwdArgs* a, int32_t head_dim, const int32_t* kv_len, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(a != nullptr, "fd_attn_bwd_generic: null args");
    const FdAttnArgs& f = a->f;
    FD_CHECK_ARG(f.q && f.k && f.v && f.o && f.lse && a->d_o && a->dq && a->dk && a->dv && a->delta,
                 "fd_attn_bwd_generic: null tensor");
    FD_CHECK_ARG(f.B > 0 && f.H > 0 && f.Nq > 0 && f.Nkv > 0, "fd_attn_bwd_generic: bad sizes");
    const int d = head_dim;
    FD_CHECK_ARG(d % 16 == 0 && d >= 16 && d <= 192, "fd_attn_bwd_generic: head_dim=%d must be a multiple of 16 in [16,192]", d);
    FD_CHECK_ARG(f.ldq % 8 == 0 && f.ldk % 8 == 0 && f.ldv % 8 == 0 && f.ldo % 8 == 0 && a->lddo % 8 == 0 &&
                     a->lddq % 8 == 0 && a->lddk % 8 == 0 && a->lddv % 8 == 0,
                 "fd_attn_bwd_generic: row strides must be multiples of 8");
    const long long total = (long long)f.B * f.H * f.Nq;
    attn_delta_g_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(
        (const bf16*)f.o, f.ldo, f.o_batch_stride, (const bf16*)a->d_o, a->lddo, a->do_batch_stride, a->delta, f.B,
        f.H, f.Nq, d);
    FD_CHECK_LAUNCH();

    if (d > 80) {
        // CUDA-core path for short sequences; dq_accum is the P / dS scratch (2 * B*H*Nq*Nkv floats)
        FD_CHECK_ARG(a->dq_accum != nullptr, "fd_attn_bwd_generic: scratch required");
        FD_CHECK_ARG((long long)f.Nq * f.Nkv <= (1LL << 20),
                     "fd_attn_bwd_generic: head_dim %d > 80 runs the short-sequence kernel (Nq*Nkv <= 2^20, got %d x %d)",
                     d, f.Nq, f.Nkv);
        AttnBwdSParams s;
        s.Nq = f.Nq; s.Nkv = f.Nkv; s.H = f.H; s.d = d; s.scale = f.scale;
        s.q = (const bf16*)f.q; s.k = (const bf16*)f.k; s.v = (const bf16*)f.v; s.d_o = (const bf16*)a->d_o;
        s.ldq = f.ldq; s.q_bs = f.q_batch_stride; s.ldk = f.ldk; s.k_bs = f.k_batch_stride;
        s.ldv = f.ldv; s.v_bs = f.v_batch_stride; s.lddo = a->lddo; s.do_bs = a->do_batch_stride;
        s.lse = f.lse; s.delta = a->delta; s.kv_len = kv_len;
        s.dq = (bf16*)a->dq; s.dk = (bf16*)a->dk; s.dv = (bf16*)a->dv;
        s.lddq = a->lddq; s.dq_bs = a->dq_batch_stride; s.lddk = a->lddk; s.dk_bs = a->dk_batch_stride;
        s.lddv = a->lddv; s.dv_bs = a->dv_batch_stride;
        float* gP = a->dq_accum;
        float* gDS = gP + (size_t)f.B * f.H * f.Nq * f.Nkv;
        ProfScope prof(stream, PROF_ATTN_BWD, 10.0 * (double)f.B * f.H * (double)f.Nq * (double)f.Nkv * d);
        const long long pairs = (long long)f.Nq * f.Nkv;
        attn_bwd_small_p_kernel<<<dim3((unsigned)((pairs + 255) / 256), f.H, f.B), 256, 0, stream>>>(s, gP, gDS);
        FD_CHECK_LAUNCH();
        attn_bwd_small_kv_kernel<<<dim3((f.Nkv * (d / 8) + 127) / 128, f.H, f.B), 128, 0, stream>>>(s, gP, gDS);
        FD_CHECK_LAUNCH();
        attn_bwd_small_q_kernel<<<dim3((f.Nq * (d / 8) + 127) / 128, f.H, f.B), 128, 0, stream>>>(s, gDS);
        FD_CHECK_LAUNCH();
        return 0;
    }

    FD_CHECK_ARG(a->dq_accum != nullptr, "fd_attn_bwd_generic: dq_accum scratch required");
    const size_t dq_elems = (size_t)f.B * f.Nq * f.H * d;
    FD_CHECK_CUDA(cudaMemsetAsync(a->dq_accum, 0, dq_elems * sizeof(float), stream));
    const int HD = f.H * d;
    CUtensorMap tq, tk, tv, tdo;
    int rc;
    if ((rc = gb_tmap(&tq, f.q, HD, f.Nq, f.B, f.ldq, f.q_batch_stride))) return rc;
    if ((rc = gb_tmap(&tk, f.k, HD, f.Nkv, f.B, f.ldk, f.k_batch_stride))) return rc;
    if ((rc = gb_tmap(&tv, f.v, HD, f.Nkv, f.B, f.ldv, f.v_batch_stride))) return rc;
    if ((rc = gb_tmap(&tdo, a->d_o, HD, f.Nq, f.B, a->lddo, a->do_batch_stride))) return rc;
    AttnBwdGParams p;
    p.Nq = f.Nq; p.Nkv = f.Nkv; p.H = f.H; p.d = d;
    p.scale = f.scale;
    p.scale_log2 = f.scale * 1.4426950408889634f;
    p.lse = f.lse;
    p.delta = a->delta;
    p.kv_len = kv_len;
    p.dq_accum = a->dq_accum;
    p.dk = (bf16*)a->dk; p.lddk = a->lddk; p.dk_bs = a->dk_batch_stride;
    p.dv = (bf16*)a->dv; p.lddv = a->lddv; p.dv_bs = a->dv_batch_stride;
    dim3 grid((f.Nkv + GB_T - 1) / GB_T, f.H, f.B);
    {
        ProfScope prof(stream, PROF_ATTN_BWD, 10.0 * (double)f.B * f.H * (double)f.Nq * (double)f.Nkv * d);
        if (d <= 64)
            rc = gb_launch<1, 2>(tq, tk, tv, tdo, p, grid, stream);
        else
            rc = gb_launch<2, 1>(tq, tk, tv, tdo, p, grid, stream);
    }
    if (rc) return rc;
    FD_CHECK_ARG(a->dq_batch_stride % 4 == 0, "fd_attn_bwd_generic: dq batch stride must be a multiple of 4");
    const long long total4 = (long long)dq_elems / 4;
    attn_dq_convert_g_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, stream>>>(
        a->dq_accum, (bf16*)a->dq, a->lddq, a->dq_batch_stride, f.Nq, HD, total4);
    FD_CHECK_LAUNCH();
    return 0;
}
