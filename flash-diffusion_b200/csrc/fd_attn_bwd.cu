// fd_attn_bwd.cu — FlashAttention-style backward for sm_100a (head dim 64, bf16, fp32 accumulate).
//
// One CTA per (128-key tile, head, batch); it loops over the 128-query tiles:
//   S  = Q K^T            P  = exp(S*scale - LSE)
//   dP = dO V^T           dS = P * (dP - delta) * scale        delta = rowsum(O * dO)
//   dV += P^T dO          dK += dS^T Q          dQ += dS K  (fp32 atomics into dq_accum)
// All five products run on tcgen05 with TMEM accumulators:
//   columns [0,128) S, [128,256) dP, [256,320) dV, [320,384) dK, [384,448) dQ.
// P and dS are written once to shared memory as 128B-swizzled [query][key] tiles and are consumed
// both K-major (dQ = dS K) and MN-major (dV = P^T dO, dK = dS^T Q) — no transposes are materialised;
// Q, dO, K are likewise consumed MN-major where the contraction runs over their row index.
// Warps: 0 TMA producer, 1 MMA issuer, 2-9 softmax/dS (two warps per TMEM lane quarter, 64 keys of one query row per
// thread: the exp / dS arithmetic is the critical path of a tile, ncu r01), 10-13 dQ drain.
//
// UPSTREAM math: autograd of F.scaled_dot_product_attention (student-LoRA backward and the GAN
// generator path through the frozen teacher, reference flash_diffusion_model.py:260-265,563-592).
#include "fd_common.cuh"
#include "fd_host.h"

namespace fd {

// dq (bf16, row stride lddq, batch stride dq_bs) <- dq_accum (fp32, contiguous [B, Nq, HD])
__global__ void attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long lddq,
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

constexpr int AB_T = 128;                   // tile edge (queries and keys)
constexpr int AB_D = 64;
constexpr int AB_TILE = AB_T * AB_D * 2;    // 16 KB
constexpr int AB_THREADS = 448;
// K, V (resident) + 2 x (Q, dO) + P + dS
constexpr int AB_SMEM = 2 * AB_TILE + 4 * AB_TILE + 2 * (2 * AB_TILE) + 256 + 1024;

struct AttnBwdKParams {
    int Nq, Nkv, H;
    float scale, scale_log2;
    const float* lse;     // [B,H,Nq]
    const float* delta;   // [B,H,Nq]
    float* dq_accum;      // [B,Nq,H*64] fp32
    bf16* dk; long long lddk, dk_bs;
    bf16* dv; long long lddv, dv_bs;
};

__device__ __forceinline__ float ab_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void attn_delta_kernel(const bf16* __restrict__ o, long long ldo, long long o_bs,
                                  const bf16* __restrict__ d_o, long long lddo, long long do_bs,
                                  float* __restrict__ delta, int B, int H, int Nq) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * H * Nq;
    if (idx >= total) return;
    const int row = (int)(idx % Nq);
    const int h = (int)((idx / Nq) % H);
    const int b = (int)(idx / ((long long)Nq * H));
    const uint4* po = reinterpret_cast<const uint4*>(o + (long long)b * o_bs + (long long)row * ldo + h * AB_D);
    const uint4* pd = reinterpret_cast<const uint4*>(d_o + (long long)b * do_bs + (long long)row * lddo + h * AB_D);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 a = po[i], c = pd[i];
        float2 x, y;
        x = unpack_bf16x2(a.x); y = unpack_bf16x2(c.x); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16x2(a.y); y = unpack_bf16x2(c.y); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16x2(a.z); y = unpack_bf16x2(c.z); acc += x.x * y.x + x.y * y.y;
        x = unpack_bf16x2(a.w); y = unpack_bf16x2(c.w); acc += x.x * y.x + x.y * y.y;
    }
    delta[idx] = acc;
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const AttnBwdKParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sK = smem;
    uint8_t* sV = sK + AB_TILE;
    uint8_t* sQ = sV + AB_TILE;            // 2 stages
    uint8_t* sDO = sQ + 2 * AB_TILE;       // 2 stages
    uint8_t* sP = sDO + 2 * AB_TILE;       // [q][kv] as 2 sub-tiles of [128][64]
    uint8_t* sDS = sP + 2 * AB_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + 2 * AB_TILE);
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
    const int n_q_tiles = (p.Nq + AB_T - 1) / AB_T;

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
                   tmem_dK = tmem_base + 320, tmem_dQ = tmem_base + 384;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(kv_full, 2 * AB_TILE);
            tma_load_3d(&tmK, kv_full, sK, head * AB_D, kv_tile * AB_T, batch);
            tma_load_3d(&tmV, kv_full, sV, head * AB_D, kv_tile * AB_T, batch);
            for (int i = 0; i < n_q_tiles; ++i) {
                const int st = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                mbar_wait(&qdo_empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&qdo_full[st], 2 * AB_TILE);
                tma_load_3d(&tmQ, &qdo_full[st], sQ + st * AB_TILE, head * AB_D, i * AB_T, batch);
                tma_load_3d(&tmDO, &qdo_full[st], sDO + st * AB_TILE, head * AB_D, i * AB_T, batch);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);    // Q K^T, dO V^T
            constexpr uint32_t id_t = make_idesc_bf16(128, 64, 1, 1);     // P^T dO, dS^T Q (both MN-major)
            constexpr uint32_t id_q = make_idesc_bf16(128, 64, 0, 1);     // dS K  (A K-major, B MN-major)
            mbar_wait(kv_full, 0);
            const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
            const uint32_t p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
            for (int i = 0; i < n_q_tiles; ++i) {
                const int st = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                mbar_wait(&qdo_full[st], ph);
                tc_fence_after();
                const uint32_t q_addr = smem_u32(sQ + st * AB_TILE);
                const uint32_t do_addr = smem_u32(sDO + st * AB_TILE);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_bf16(tmem_S, make_desc_k_sw128(q_addr + k * 32), make_desc_k_sw128(k_addr + k * 32),
                                id_s, k != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_bf16(tmem_dP, make_desc_k_sw128(do_addr + k * 32), make_desc_k_sw128(v_addr + k * 32),
                                id_s, k != 0 ? 1u : 0u);
                tc_commit(s_full);
                mbar_wait(pds_ready, i & 1);
                tc_fence_after();
                // dV += P^T dO ; dK += dS^T Q      (contraction over the 128 query rows, 16 per MMA)
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_bf16(tmem_dV, make_desc_mn_sw128(p_addr + k * 2048, AB_TILE, 1024),
                                make_desc_mn_sw128(do_addr + k * 2048, 0, 1024), id_t, (i | k) != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_bf16(tmem_dK, make_desc_mn_sw128(ds_addr + k * 2048, AB_TILE, 1024),
                                make_desc_mn_sw128(q_addr + k * 2048, 0, 1024), id_t, (i | k) != 0 ? 1u : 0u);
                // dQ_i = dS K                    (contraction over the 128 keys)
                if (i > 0) {
                    mbar_wait(dq_empty, (i - 1) & 1);
                    tc_fence_after();
                }
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    tc_mma_bf16(tmem_dQ, make_desc_k_sw128(ds_addr + (k >> 2) * AB_TILE + (k & 3) * 32),
                                make_desc_mn_sw128(k_addr + k * 2048, 0, 1024), id_q, k != 0 ? 1u : 0u);
                tc_commit(&qdo_empty[st]);
                tc_commit(mma2_done);
                tc_commit(dq_full);
            }
        }
    } else if (warp < 10) {
        // softmax / dS warps: thread (row, half) owns 64 keys of one query row
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const int kv_valid = min(AB_T, p.Nkv - kv_tile * AB_T);
        const float* lse_bh = p.lse + ((long long)batch * p.H + head) * p.Nq;
        const float* delta_bh = p.delta + ((long long)batch * p.H + head) * p.Nq;
        int q_row = row;
        float lse_next = q_row < p.Nq ? lse_bh[q_row] : 0.f;
        float dlt_next = q_row < p.Nq ? delta_bh[q_row] : 0.f;
        for (int i = 0; i < n_q_tiles; ++i) {
            q_row = i * AB_T + row;
            const bool q_ok = q_row < p.Nq;
            const float lse2 = lse_next * 1.4426950408889634f;
            const float dlt = dlt_next;
            if (i + 1 < n_q_tiles) {           // next tile's row scalars: off the critical path
                const int nr = q_row + AB_T;
                lse_next = nr < p.Nq ? lse_bh[nr] : 0.f;
                dlt_next = nr < p.Nq ? delta_bh[nr] : 0.f;
            }
            const bool full = (i + 1) * AB_T <= p.Nq && kv_valid == AB_T;      // block-uniform: no masking needed
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
                        const float p0 = ab_ex2(fmaf(__uint_as_float(rs[j]), p.scale_log2, -lse2));
                        const float p1 = ab_ex2(fmaf(__uint_as_float(rs[j + 1]), p.scale_log2, -lse2));
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
                            p0 = ab_ex2(fmaf(__uint_as_float(rs[j]), p.scale_log2, -lse2));
                            d0 = p0 * (__uint_as_float(rp[j]) - dlt) * p.scale;
                        }
                        if (q_ok && c * 32 + j + 1 < kv_valid) {
                            p1 = ab_ex2(fmaf(__uint_as_float(rs[j + 1]), p.scale_log2, -lse2));
                            d1 = p1 * (__uint_as_float(rp[j + 1]) - dlt) * p.scale;
                        }
                        pk[j >> 1] = pack_bf16x2(p0, p1);
                        dk_[j >> 1] = pack_bf16x2(d0, d1);
                    }
                }
                if (cc == 0 && i > 0) mbar_wait(mma2_done, (i - 1) & 1);   // P / dS buffers free again
                uint8_t* subp = sP + half * AB_TILE + row * 128;
                uint8_t* subd = sDS + half * AB_TILE + row * 128;
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
        const int kv_row = kv_tile * AB_T + row;
        bf16* dk_row = p.dk + (long long)batch * p.dk_bs + (long long)kv_row * p.lddk + head * AB_D;
        bf16* dv_row = p.dv + (long long)batch * p.dv_bs + (long long)kv_row * p.lddv + head * AB_D;
        {
            const int c = half;
            uint32_t rk[32], rv[32];
            tmem_ld_32x32(tmem_dK + lane_base + c * 32, rk);
            tmem_ld_32x32(tmem_dV + lane_base + c * 32, rv);
            tmem_ld_wait();
            if (kv_row < p.Nkv) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 u, w;
                    u.x = pack_bf16x2(__uint_as_float(rk[8 * q4 + 0]), __uint_as_float(rk[8 * q4 + 1]));
                    u.y = pack_bf16x2(__uint_as_float(rk[8 * q4 + 2]), __uint_as_float(rk[8 * q4 + 3]));
                    u.z = pack_bf16x2(__uint_as_float(rk[8 * q4 + 4]), __uint_as_float(rk[8 * q4 + 5]));
                    u.w = pack_bf16x2(__uint_as_float(rk[8 * q4 + 6]), __uint_as_float(rk[8 * q4 + 7]));
                    w.x = pack_bf16x2(__uint_as_float(rv[8 * q4 + 0]), __uint_as_float(rv[8 * q4 + 1]));
                    w.y = pack_bf16x2(__uint_as_float(rv[8 * q4 + 2]), __uint_as_float(rv[8 * q4 + 3]));
                    w.z = pack_bf16x2(__uint_as_float(rv[8 * q4 + 4]), __uint_as_float(rv[8 * q4 + 5]));
                    w.w = pack_bf16x2(__uint_as_float(rv[8 * q4 + 6]), __uint_as_float(rv[8 * q4 + 7]));
                    *reinterpret_cast<uint4*>(dk_row + c * 32 + q4 * 8) = u;
                    *reinterpret_cast<uint4*>(dv_row + c * 32 + q4 * 8) = w;
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
            const int q_row = i * AB_T + row;
            float* dst = p.dq_accum + ((long long)batch * p.Nq + q_row) * (p.H * AB_D) + head * AB_D;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_dQ + lane_base + c * 32, r);
                tmem_ld_wait();
                if (q_row < p.Nq) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        atomicAdd(reinterpret_cast<float4*>(dst + c * 32 + j),
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

static int make_tmap3(CUtensorMap* m, const void* base, int H, int N, int B, int64_t ld, int64_t bs) {
    const uint64_t dims[3] = {(uint64_t)H * AB_D, (uint64_t)N, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)bs * 2};
    const uint32_t box[3] = {(uint32_t)AB_D, 128u, 1u};
    return encode_tmap_bf16(m, base, 3, dims, str, box);
}

}  // namespace fd

using namespace fd;

extern "C" int fd_attn_bwd(const FdAttnBwdArgs* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(a != nullptr, "fd_attn_bwd: null args");
    const FdAttnArgs& f = a->f;
    FD_CHECK_ARG(f.q && f.k && f.v && f.o && f.lse && a->d_o && a->dq && a->dk && a->dv && a->delta && a->dq_accum,
                 "fd_attn_bwd: null tensor");
    FD_CHECK_ARG(f.B > 0 && f.H > 0 && f.Nq > 0 && f.Nkv > 0, "fd_attn_bwd: bad sizes");
    FD_CHECK_ARG(f.ldo % 8 == 0 && a->lddo % 8 == 0 && a->lddk % 8 == 0 && a->lddv % 8 == 0,
                 "fd_attn_bwd: row strides must be multiples of 8");
    const long long total = (long long)f.B * f.H * f.Nq;
    attn_delta_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(
        (const bf16*)f.o, f.ldo, f.o_batch_stride, (const bf16*)a->d_o, a->lddo, a->do_batch_stride, a->delta, f.B,
        f.H, f.Nq);
    FD_CHECK_LAUNCH();
    const size_t dq_elems = (size_t)f.B * f.Nq * f.H * AB_D;
    FD_CHECK_CUDA(cudaMemsetAsync(a->dq_accum, 0, dq_elems * sizeof(float), stream));
    CUtensorMap tq, tk, tv, tdo;
    int rc;
    if ((rc = make_tmap3(&tq, f.q, f.H, f.Nq, f.B, f.ldq, f.q_batch_stride))) return rc;
    if ((rc = make_tmap3(&tk, f.k, f.H, f.Nkv, f.B, f.ldk, f.k_batch_stride))) return rc;
    if ((rc = make_tmap3(&tv, f.v, f.H, f.Nkv, f.B, f.ldv, f.v_batch_stride))) return rc;
    if ((rc = make_tmap3(&tdo, a->d_o, f.H, f.Nq, f.B, a->lddo, a->do_batch_stride))) return rc;
    AttnBwdKParams p;
    p.Nq = f.Nq; p.Nkv = f.Nkv; p.H = f.H;
    p.scale = f.scale;
    p.scale_log2 = f.scale * 1.4426950408889634f;
    p.lse = f.lse;
    p.delta = a->delta;
    p.dq_accum = a->dq_accum;
    p.dk = (bf16*)a->dk; p.lddk = a->lddk; p.dk_bs = a->dk_batch_stride;
    p.dv = (bf16*)a->dv; p.lddv = a->lddv; p.dv_bs = a->dv_batch_stride;
    static bool attr_set = false;
    if (!attr_set) {
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
        attr_set = true;
    }
    dim3 grid((f.Nkv + AB_T - 1) / AB_T, f.H, f.B);
    {
    ProfScope prof(stream, PROF_ATTN_BWD, 10.0 * (double)f.B * f.H * (double)f.Nq * (double)f.Nkv * AB_D);
    attn_bwd_kernel<<<grid, AB_THREADS, AB_SMEM, stream>>>(tq, tk, tv, tdo, p);
    }
    FD_CHECK_LAUNCH();
    FD_CHECK_ARG(a->lddq % 4 == 0 && a->dq_batch_stride % 4 == 0, "fd_attn_bwd: dq strides must be multiples of 4");
    const long long total4 = (long long)dq_elems / 4;
    attn_dq_convert_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, stream>>>(
        a->dq_accum, (bf16*)a->dq, a->lddq, a->dq_batch_stride, f.Nq, f.H * AB_D, total4);
    FD_CHECK_LAUNCH();
    return 0;
}
