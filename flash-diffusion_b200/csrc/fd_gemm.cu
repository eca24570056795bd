// fd_gemm.cu — persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   acc[M,N] = A1[M,K1] B1[N,K1]^T (+ A2[M,K2] B2[N,K2]^T),  bf16 operands, fp32 accumulate in TMEM.
//
// Structure (one CTA per SM, 256 threads):
//   warp 0    TMA producer   : cp.async.bulk.tensor (2D plain / 4D NHWC conv taps) -> smem ring
//   warp 1    MMA issuer     : one lane issues tcgen05.mma (128 x BN x 16), commits to mbarriers
//   warp 2    TMEM allocator
//   warps 4-7 epilogue       : tcgen05.ld 32x32b -> bias / rowvec / GEGLU / residual -> global
// TMEM holds two accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the MMA
// main loop of tile i+1.  Operand tiles are 64 bf16 (=128 B) wide in K with the 128-byte swizzle,
// written by TMA and consumed through K-major shared-memory descriptors.
//
// Reference math replaced: torch linear / conv2d calls under
// src/flash/models/unets/unet.py:108-119 (see include/flashb200.h, fd_gemm).
#include <stdlib.h>

#include "fd_common.cuh"
#include "fd_host.h"

namespace fd {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 384;   // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: epilogue (2 per lane quarter)

struct GemmKParams {
    int M, N, kb1, kb2, num_m_tiles, num_n_tiles, group_m;
    int conv_taps, cblocks, H, W, tile_w, tile_h;
    int tap_dn[FD_MAX_TAPS], tap_dh[FD_MAX_TAPS], tap_dw[FD_MAX_TAPS];
    const float* bias;
    const float* rowvec;
    int rows_per_group;
    long long ldrv;
    int geglu;
    const bf16* residual;
    long long ldr;
    void* out;
    long long ldo;
    int out_fp32;
    const float* ln_stats;
    const float* ln_colsum;
    float ln_inv_c, ln_eps;
    float* rowstats_out;
    int act;
    const float* rowscale;
    int rows_per_group_scale;
    long long ldrs;
};

template <int BN>
struct GemmCfg {
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int TMEM_COLS = 2 * BN;  // 128 / 256 / 512: powers of two
    static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 256 + 1024;
};

// Tile rasterisation: consecutive tile indices (= what the 148 persistent CTAs work on at the same time) walk
// GROUP_M m-tiles x all n-tiles, m fastest, so a wave touches ~16 A row-panels and ~9 B panels instead of
// 148 A panels and 1 B panel: each A panel is fetched from DRAM once per group and re-used out of L2.
__device__ __forceinline__ void tile_coords(const GemmKParams& p, int tile, int& mt, int& nt) {
    const int per_group = p.group_m * p.num_n_tiles;
    const int g = tile / per_group;
    const int r = tile - g * per_group;
    const int gm = min(p.group_m, p.num_m_tiles - g * p.group_m);
    nt = r / gm;
    mt = g * p.group_m + (r - nt * gm);
}

// per-row state carried across the 32-column chunks of one tile
struct RowState {
    float ln_mean, ln_rstd;   // LayerNorm fold inputs for this row
    float s1, s2;             // running (sum, sum of squares) of the stored outputs (rowstats_out)
};

__device__ __forceinline__ void row_state_init(const GemmKParams& p, int row, RowState& rs) {
    rs.s1 = rs.s2 = 0.f;
    rs.ln_mean = 0.f;
    rs.ln_rstd = 1.f;
    if (p.ln_stats != nullptr && row < p.M) {
        const float2 st = *reinterpret_cast<const float2*>(p.ln_stats + 2 * (long long)row);
        const float mean = st.x * p.ln_inv_c;
        const float var = fmaxf(st.y * p.ln_inv_c - mean * mean, 0.f);
        rs.ln_mean = mean;
        rs.ln_rstd = rsqrtf(var + p.ln_eps);
    }
}

__device__ __forceinline__ void row_state_flush(const GemmKParams& p, int row, const RowState& rs) {
    if (p.rowstats_out != nullptr && row < p.M) {
        atomicAdd(p.rowstats_out + 2 * (long long)row, rs.s1);
        atomicAdd(p.rowstats_out + 2 * (long long)row + 1, rs.s2);
    }
}

__device__ __forceinline__ void stats_of_bf16x2(uint32_t u, RowState& rs) {
    const float2 f = unpack_bf16x2(u);
    rs.s1 += f.x + f.y;
    rs.s2 += f.x * f.x + f.y * f.y;
}

__device__ __forceinline__ void epilogue_chunk(const GemmKParams& p, int row, int col0,
                                               uint32_t (&r)[32], RowState& rs) {
    // r holds acc[row, col0 .. col0+31] as fp32 bit patterns
    const bool row_ok = row < p.M;
    const int N = p.N;
    if (col0 >= N) return;
    const bool full = (col0 + 32 <= N);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

    if (p.ln_stats != nullptr) {
        // LayerNorm(x) W^T = rstd * (x W'^T - mean * colsum(W'))
        const float nm = -rs.ln_mean;
        if (full) {
            const float4* c4 = reinterpret_cast<const float4*>(p.ln_colsum + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 c = __ldg(c4 + j);
                v[4 * j + 0] = rs.ln_rstd * fmaf(nm, c.x, v[4 * j + 0]);
                v[4 * j + 1] = rs.ln_rstd * fmaf(nm, c.y, v[4 * j + 1]);
                v[4 * j + 2] = rs.ln_rstd * fmaf(nm, c.z, v[4 * j + 2]);
                v[4 * j + 3] = rs.ln_rstd * fmaf(nm, c.w, v[4 * j + 3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < N) v[j] = rs.ln_rstd * fmaf(nm, __ldg(p.ln_colsum + col0 + j), v[j]);
        }
    }

    if (p.bias != nullptr) {
        if (full) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 b = __ldg(b4 + j);
                v[4 * j + 0] += b.x;
                v[4 * j + 1] += b.y;
                v[4 * j + 2] += b.z;
                v[4 * j + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < N) v[j] += __ldg(p.bias + col0 + j);
        }
    }
    if (p.rowvec != nullptr && row_ok) {
        const float* rv = p.rowvec + (long long)(row / p.rows_per_group) * p.ldrv + col0;
        if (full && (p.ldrv & 3) == 0) {
            const float4* b4 = reinterpret_cast<const float4*>(rv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 b = __ldg(b4 + j);
                v[4 * j + 0] += b.x;
                v[4 * j + 1] += b.y;
                v[4 * j + 2] += b.z;
                v[4 * j + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < N) v[j] += __ldg(rv + j);
        }
    }
    if (!row_ok) return;

    if (p.act == 1) {   // gelu (tanh approximation): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float x = v[j];
            const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
            v[j] = 0.5f * x * (1.0f + tanhf(u));
        }
    }
    if (p.rowscale != nullptr) {
        const float* sv = p.rowscale + (long long)(row / p.rows_per_group_scale) * p.ldrs + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j < N) v[j] *= __ldg(sv + j);
    }

    if (p.geglu) {
        // interleaved packing: cols [0,16) value, [16,32) gate -> 16 outputs at col0/2
        const int oc0 = col0 >> 1;
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = v[j] * gelu_erf(v[16 + j]);
        if (p.residual != nullptr) {
            const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + (long long)row * p.ldr + oc0);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint4 u = __ldg(r4 + j);
                float2 f;
                f = unpack_bf16x2(u.x); o[8 * j + 0] += f.x; o[8 * j + 1] += f.y;
                f = unpack_bf16x2(u.y); o[8 * j + 2] += f.x; o[8 * j + 3] += f.y;
                f = unpack_bf16x2(u.z); o[8 * j + 4] += f.x; o[8 * j + 5] += f.y;
                f = unpack_bf16x2(u.w); o[8 * j + 6] += f.x; o[8 * j + 7] += f.y;
            }
        }
        if (p.out_fp32) {
            float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + oc0);
#pragma unroll
            for (int j = 0; j < 4; ++j) o4[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        } else {
            uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + oc0);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint4 u = make_uint4(pack_bf16x2(o[8 * j], o[8 * j + 1]), pack_bf16x2(o[8 * j + 2], o[8 * j + 3]),
                                           pack_bf16x2(o[8 * j + 4], o[8 * j + 5]), pack_bf16x2(o[8 * j + 6], o[8 * j + 7]));
                o4[j] = u;
                if (p.rowstats_out != nullptr) {
                    stats_of_bf16x2(u.x, rs); stats_of_bf16x2(u.y, rs); stats_of_bf16x2(u.z, rs); stats_of_bf16x2(u.w, rs);
                }
            }
        }
        return;
    }

    const bool vec_ok = full && ((p.ldo & 7) == 0) && (p.residual == nullptr || (p.ldr & 7) == 0);
    if (vec_ok) {
        if (p.residual != nullptr) {
            const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + (long long)row * p.ldr + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 u = __ldg(r4 + j);
                float2 f;
                f = unpack_bf16x2(u.x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
                f = unpack_bf16x2(u.y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
                f = unpack_bf16x2(u.z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
                f = unpack_bf16x2(u.w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
            }
        }
        if (p.out_fp32) {
            float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
            uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 u = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                o4[j] = u;
                if (p.rowstats_out != nullptr) {
                    stats_of_bf16x2(u.x, rs); stats_of_bf16x2(u.y, rs); stats_of_bf16x2(u.z, rs); stats_of_bf16x2(u.w, rs);
                }
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (col0 + j < N) {
                float x = v[j];
                if (p.residual != nullptr)
                    x += __bfloat162float(p.residual[(long long)row * p.ldr + col0 + j]);
                if (p.out_fp32)
                    reinterpret_cast<float*>(p.out)[(long long)row * p.ldo + col0 + j] = x;
                else {
                    const bf16 xb = __float2bfloat16(x);
                    reinterpret_cast<bf16*>(p.out)[(long long)row * p.ldo + col0 + j] = xb;
                    const float xf = __bfloat162float(xb);
                    rs.s1 += xf;
                    rs.s2 += xf * xf;
                }
            }
        }
    }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
            const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
            const GemmKParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (required by the 128-byte swizzle atoms)
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB1);
        if (p.kb2 > 0) {
            tma_prefetch_desc(&tmA2);
            tma_prefetch_desc(&tmB2);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 256);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();                  // everything above overlapped the previous kernel's tail (see launch_gemm)
    pdl_launch_dependents();

    const int total_tiles = p.num_m_tiles * p.num_n_tiles;
    const int kb_total = p.kb1 + p.kb2;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int mt, nt;
                tile_coords(p, tile, mt, nt);
                int n0 = 0, h0 = 0, w0 = 0;
                if (p.conv_taps) {
                    const int row0 = mt * BM;
                    const int hw = p.H * p.W;
                    n0 = row0 / hw;
                    const int rem = row0 - n0 * hw;
                    h0 = rem / p.W;
                    w0 = rem - h0 * p.W;
                }
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
                    uint8_t* a_dst = sA + stage * Cfg::A_BYTES;
                    uint8_t* b_dst = sB + stage * Cfg::B_BYTES;
                    if (kb < p.kb1) {
                        if (p.conv_taps) {
                            const int tap = kb / p.cblocks;
                            const int cb = kb - tap * p.cblocks;
                            tma_load_4d(&tmA1, &full[stage], a_dst, cb * BK, w0 + p.tap_dw[tap],
                                        h0 + p.tap_dh[tap], n0 + p.tap_dn[tap]);
                        } else {
                            tma_load_2d(&tmA1, &full[stage], a_dst, kb * BK, mt * BM);
                        }
                        tma_load_2d(&tmB1, &full[stage], b_dst, kb * BK, nt * BN);
                    } else {
                        const int k2 = kb - p.kb1;
                        tma_load_2d(&tmA2, &full[stage], a_dst, k2 * BK, mt * BM);
                        tma_load_2d(&tmB2, &full[stage], b_dst, k2 * BK, nt * BN);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adesc = make_desc_k_sw128(a_addr + k * 32);
                        const uint64_t bdesc = make_desc_k_sw128(b_addr + k * 32);
                        tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(&empty[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tc_commit(&tfull[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        const int q = (warp - 4) & 3;      // TMEM lane quarter == warp_id % 4
        const int half = (warp - 4) >> 2;  // which half of the tile's columns this warp drains
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int mt, nt;
            tile_coords(p, tile, mt, nt);
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int row = mt * BM + q * 32 + lane;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            RowState rs;
            row_state_init(p, row, rs);
#pragma unroll 1
            for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
                uint32_t r[32];
                tmem_ld_32x32(t_addr + c * 32, r);
                tmem_ld_wait();
                epilogue_chunk(p, row, nt * BN + c * 32, r, rs);
            }
            row_state_flush(p, row, rs);
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ CTA-pair kernel
// 256 x BN output tile per CTA PAIR (cluster of 2, tcgen05 cta_group::2).  Each CTA stages its own 128 rows of A
// and its own BN/2 rows of B; the pair's tensor cores read both halves, so per-SM shared-memory traffic per MMA
// (and L2->SM operand traffic) is half that of the single-CTA 128 x BN tile: 64 B/clk read + 64 B/clk TMA fill
// at BN = 256 instead of 96 + 96.  The leader CTA (rank 0) issues every MMA; both CTAs run a TMA producer and an
// epilogue for their own 128 accumulator rows (TMEM lanes).
template <int BN>
struct PairCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int BH_BYTES = (BN / 2) * BK * 2;
    static constexpr int STAGES = (BN == 256) ? 6 : (BN == 160 ? 7 : 8);
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;   // power of two >= 2 accumulator stages
    static constexpr int SMEM_BYTES = STAGES * (A_BYTES + BH_BYTES) + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const GemmKParams p) {
    using Cfg = PairCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::BH_BYTES);
    uint64_t* full = bars;               // used in the leader CTA only
    uint64_t* empty = bars + STAGES;     // per CTA
    uint64_t* tfull = bars + 2 * STAGES; // per CTA
    uint64_t* tempty = tfull + 2;        // used in the leader CTA only
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB1);
        if (p.kb2 > 0) {
            tma_prefetch_desc(&tmA2);
            tma_prefetch_desc(&tmB2);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 2);      // leader's arrive.expect_tx + peer's remote arrive
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 16);   // one arrival per epilogue warp (8) of each CTA
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_holder, Cfg::TMEM_COLS);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();
    pdl_launch_dependents();

    const int num_m2 = (p.M + 2 * BM - 1) / (2 * BM);
    const int total_tiles = num_m2 * p.num_n_tiles;
    const int kb_total = p.kb1 + p.kb2;
    const int pair_id = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    GemmKParams pp = p;
    pp.num_m_tiles = num_m2;             // tile_coords works on pair tiles

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
                int mt2, nt;
                tile_coords(pp, tile, mt2, nt);
                const int row0 = mt2 * 2 * BM + (int)rank * BM;
                int n0 = 0, h0 = 0, w0 = 0;
                if (p.conv_taps) {
                    const int hw = p.H * p.W;
                    n0 = row0 / hw;
                    const int rem = row0 - n0 * hw;
                    h0 = rem / p.W;
                    w0 = rem - h0 * p.W;
                }
                const int brow = nt * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), 0);
                    if (leader)
                        mbar_arrive_expect_tx(&full[stage], 2 * (Cfg::A_BYTES + Cfg::BH_BYTES));
                    else
                        mbar_arrive_cluster(full_leader);
                    uint8_t* a_dst = sA + stage * Cfg::A_BYTES;
                    uint8_t* b_dst = sB + stage * Cfg::BH_BYTES;
                    if (kb < p.kb1) {
                        if (p.conv_taps) {
                            const int tap = kb / p.cblocks;
                            const int cb = kb - tap * p.cblocks;
                            tma_load_4d_pair(&tmA1, full_leader, a_dst, cb * BK, w0 + p.tap_dw[tap],
                                             h0 + p.tap_dh[tap], n0 + p.tap_dn[tap]);
                        } else {
                            tma_load_2d_pair(&tmA1, full_leader, a_dst, kb * BK, row0);
                        }
                        tma_load_2d_pair(&tmB1, full_leader, b_dst, kb * BK, brow);
                    } else {
                        const int k2 = kb - p.kb1;
                        tma_load_2d_pair(&tmA2, full_leader, a_dst, k2 * BK, row0);
                        tma_load_2d_pair(&tmB2, full_leader, b_dst, k2 * BK, brow);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < kb_total; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::BH_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc_mma_bf16_pair(d_tmem, make_desc_k_sw128(a_addr + k * 32),
                                         make_desc_k_sw128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit_pair(&empty[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tc_commit_pair(&tfull[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        const int q = (warp - 4) & 3;
        const int half = (warp - 4) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
            int mt2, nt;
            tile_coords(pp, tile, mt2, nt);
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int row = mt2 * 2 * BM + (int)rank * BM + q * 32 + lane;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            RowState rs;
            row_state_init(p, row, rs);
            constexpr int NCH = BN / 32, SPLIT = (NCH + 1) / 2;
#pragma unroll 1
            for (int c = half ? SPLIT : 0; c < (half ? NCH : SPLIT); ++c) {
                uint32_t r[32];
                tmem_ld_32x32(t_addr + c * 32, r);
                tmem_ld_wait();
                epilogue_chunk(p, row, nt * BN + c * 32, r, rs);
            }
            row_state_flush(p, row, rs);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader)
                    mbar_arrive(&tempty[acc]);
                else
                    mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ host

// Tile configuration: BN in {64,128,256} single-CTA (128-row tiles) or {128,256} CTA-pair (256-row tiles,
// returned as 512 + BN).  Cost = waves x (tile rows x BN) / efficiency, efficiencies from the smem-traffic
// model calibrated on B200.
static int choose_bn(int M, int N, int force) {
    if (force == 64 || force == 128 || force == 256 || force == 512 + 128 || force == 512 + 160 ||
        force == 512 + 256)
        return force;
    const int sms = num_sms();
    struct Cand { int code, rows, bn; double eff; };
    static const Cand cands[6] = {{512 + 256, 256, 256, 1.00}, {512 + 160, 256, 160, 0.80}, {512 + 128, 256, 128, 0.80},
                                  {256, 128, 256, 0.72}, {128, 128, 128, 0.58}, {64, 128, 64, 0.36}};
    static const bool no_pair = getenv("FD_NO_PAIR") != nullptr;
    int best = 128;
    double best_cost = 1e30;
    for (int i = no_pair ? 3 : 0; i < 6; ++i) {
        const Cand& c = cands[i];
        const long long tiles = (long long)((M + c.rows - 1) / c.rows) * ((N + c.bn - 1) / c.bn);
        const int units = c.rows == 256 ? sms / 2 : sms;
        const long long waves = (tiles + units - 1) / units;
        const double cost = (double)waves * c.rows * c.bn / c.eff / (c.rows == 256 ? 2.0 : 1.0);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = c.code;
        }
    }
    return best;
}

// Programmatic dependent launch of the GEMM kernels (FD_PDL=1): the kernels call griddepcontrol.wait after their
// prologue, so with the launch attribute the prologue may overlap the previous kernel's tail.  Measured on B200 (SDXL
// teacher evaluation replayed from a CUDA graph, batch 8 and 16): no difference (68.53 vs 68.52 ms), so it is off by
// default; without the attribute griddepcontrol.wait returns immediately.
static bool use_pdl() {
    static const bool on = [] {
        const char* e = getenv("FD_PDL");
        return e != nullptr && atoi(e) != 0;
    }();
    return on;
}

template <int BN>
static int launch_gemm_pair(const CUtensorMap& tA1, const CUtensorMap& tB1, const CUtensorMap& tA2,
                            const CUtensorMap& tB2, const GemmKParams& p, cudaStream_t stream, bool pdl) {
    using Cfg = PairCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        FD_CHECK_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int total = ((p.M + 2 * BM - 1) / (2 * BM)) * p.num_n_tiles;
    int pairs = num_sms() / 2;
    if (total < pairs) pairs = total;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_pair_kernel<BN>, tA1, tB1, tA2, tB2, p));
    FD_CHECK_LAUNCH();
    return 0;
}

template <int BN>
static int launch_gemm(const CUtensorMap& tA1, const CUtensorMap& tB1, const CUtensorMap& tA2,
                       const CUtensorMap& tB2, const GemmKParams& p, cudaStream_t stream, bool pdl) {
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        FD_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int total = p.num_m_tiles * p.num_n_tiles;
    const int grid = total < num_sms() ? total : num_sms();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<BN>, tA1, tB1, tA2, tB2, p));
    FD_CHECK_LAUNCH();
    return 0;
}

}  // namespace fd

extern "C" int fd_gemm(const FdGemmArgs* a, void* stream_) {
    using namespace fd;
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(a != nullptr, "fd_gemm: null args");
    FD_CHECK_ARG(a->M > 0 && a->N > 0, "fd_gemm: bad M/N %d/%d", a->M, a->N);
    FD_CHECK_ARG(a->K1 > 0 && a->K2 >= 0, "fd_gemm: bad K1/K2 %d/%d", a->K1, a->K2);
    // K itself may be ragged (TMA zero-fills the tail of the last 64-wide block of BOTH operands);
    // only the row strides must be 16-byte multiples (checked when the tensor maps are encoded).
    FD_CHECK_ARG(a->a1 && a->b1 && a->out, "fd_gemm: null operand");
    FD_CHECK_ARG(!a->geglu || (a->N % 32 == 0), "fd_gemm: geglu needs N %% 32 == 0");
    FD_CHECK_ARG(!a->rowvec || a->rows_per_group > 0, "fd_gemm: rowvec needs rows_per_group");

    const int code = choose_bn(a->M, a->N, a->force_bn);
    const bool pair = code >= 512;
    const int BN = pair ? code - 512 : code;
    const uint32_t b_box_rows = pair ? BN / 2 : BN;
    GemmKParams p;
    memset(&p, 0, sizeof(p));
    p.M = a->M;
    p.N = a->N;
    p.num_m_tiles = (a->M + BM - 1) / BM;
    p.num_n_tiles = (a->N + BN - 1) / BN;
    {
        static const int env_gm = getenv("FD_GROUP_M") ? atoi(getenv("FD_GROUP_M")) : 0;
        p.group_m = env_gm > 0 ? env_gm : (pair ? 8 : 16);   // pair tiles are 256 rows: 8 x 256 = 16 x 128
    }
    p.bias = a->bias;
    p.rowvec = a->rowvec;
    p.rows_per_group = a->rows_per_group;
    p.ldrv = a->ldrv > 0 ? a->ldrv : a->N;
    p.geglu = a->geglu;
    p.residual = (const bf16*)a->residual;
    p.ldr = a->ldr;
    p.out = a->out;
    p.ldo = a->ldo;
    p.out_fp32 = a->out_fp32;
    p.ln_stats = a->ln_stats;
    p.ln_colsum = a->ln_colsum;
    p.ln_inv_c = a->ln_inv_c;
    p.ln_eps = a->ln_eps;
    p.rowstats_out = a->rowstats_out;
    p.act = a->act;
    p.rowscale = a->rowscale;
    p.rows_per_group_scale = a->rows_per_group_scale;
    p.ldrs = a->ldrs > 0 ? a->ldrs : a->N;
    FD_CHECK_ARG(!a->rowscale || a->rows_per_group_scale > 0, "fd_gemm: rowscale needs rows_per_group_scale");
    FD_CHECK_ARG(!a->ln_stats || (a->ln_colsum && a->K2 == 0), "fd_gemm: LayerNorm fold needs ln_colsum and no K2 segment");
    FD_CHECK_ARG(!a->rowstats_out || !a->out_fp32, "fd_gemm: rowstats_out needs a bf16 output");
    if (a->rowstats_out) FD_CHECK_CUDA(cudaMemsetAsync(a->rowstats_out, 0, sizeof(float) * 2 * (size_t)a->M, stream));

    CUtensorMap tA1, tB1, tA2, tB2;
    int rc;
    if (a->conv_taps > 0) {
        FD_CHECK_ARG(a->conv_taps <= FD_MAX_TAPS, "fd_gemm: too many taps");
        FD_CHECK_ARG(a->C > 0 && a->C % 8 == 0, "fd_gemm: conv C=%d must be a multiple of 8", a->C);
        FD_CHECK_ARG(a->H > 0 && a->W > 0 && a->NB_in > 0, "fd_gemm: bad conv geometry");
        const int cpad = ((a->C + BK - 1) / BK) * BK;
        FD_CHECK_ARG(a->K1 == a->conv_taps * cpad,
                     "fd_gemm: conv K1=%d must equal taps*ceil64(C)=%d", a->K1, a->conv_taps * cpad);
        int tile_w = a->W < BM ? a->W : BM;
        FD_CHECK_ARG(BM % tile_w == 0 && a->W % tile_w == 0,
                     "fd_gemm: conv W=%d must be a power of two <=128 or a multiple of 128", a->W);
        int tile_h = BM / tile_w;
        if (tile_h > a->H) tile_h = a->H;
        FD_CHECK_ARG(a->H % tile_h == 0 && BM % (tile_w * tile_h) == 0,
                     "fd_gemm: conv H=%d incompatible with tile (%d x %d)", a->H, tile_h, tile_w);
        const int tile_n = BM / (tile_w * tile_h);
        FD_CHECK_ARG((long long)a->M % ((long long)a->H * a->W) == 0,
                     "fd_gemm: conv M must be a multiple of H*W");
        p.conv_taps = a->conv_taps;
        p.cblocks = cpad / BK;
        p.H = a->H;
        p.W = a->W;
        p.tile_w = tile_w;
        p.tile_h = tile_h;
        for (int t = 0; t < a->conv_taps; ++t) {
            p.tap_dn[t] = a->tap_dn[t];
            p.tap_dh[t] = a->tap_dh[t];
            p.tap_dw[t] = a->tap_dw[t];
        }
        p.kb1 = a->conv_taps * p.cblocks;
        const uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->NB_in};
        const uint64_t str[3] = {(uint64_t)a->C * 2, (uint64_t)a->W * a->C * 2,
                                 (uint64_t)a->H * a->W * a->C * 2};
        const uint32_t box[4] = {(uint32_t)BK, (uint32_t)tile_w, (uint32_t)tile_h, (uint32_t)tile_n};
        rc = encode_tmap_bf16(&tA1, a->a1, 4, dims, str, box);
        if (rc) return rc;
    } else {
        p.kb1 = (a->K1 + BK - 1) / BK;
        const uint64_t dims[2] = {(uint64_t)a->K1, (uint64_t)a->M};
        const uint64_t str[1] = {(uint64_t)a->lda1 * 2};
        const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
        rc = encode_tmap_bf16(&tA1, a->a1, 2, dims, str, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)a->K1, (uint64_t)a->N};
        const uint64_t str[1] = {(uint64_t)a->ldb1 * 2};
        const uint32_t box[2] = {(uint32_t)BK, b_box_rows};
        rc = encode_tmap_bf16(&tB1, a->b1, 2, dims, str, box);
        if (rc) return rc;
    }
    if (a->K2 > 0) {
        FD_CHECK_ARG(a->a2 && a->b2, "fd_gemm: K2>0 but null segment-2 operands");
        p.kb2 = (a->K2 + BK - 1) / BK;
        const uint64_t dimsa[2] = {(uint64_t)a->K2, (uint64_t)a->M};
        const uint64_t stra[1] = {(uint64_t)a->lda2 * 2};
        const uint32_t boxa[2] = {(uint32_t)BK, (uint32_t)BM};
        rc = encode_tmap_bf16(&tA2, a->a2, 2, dimsa, stra, boxa);
        if (rc) return rc;
        const uint64_t dimsb[2] = {(uint64_t)a->K2, (uint64_t)a->N};
        const uint64_t strb[1] = {(uint64_t)a->ldb2 * 2};
        const uint32_t boxb[2] = {(uint32_t)BK, b_box_rows};
        rc = encode_tmap_bf16(&tB2, a->b2, 2, dimsb, strb, boxb);
        if (rc) return rc;
    } else {
        tA2 = tA1;
        tB2 = tB1;
    }
    ProfScope prof(stream, a->conv_taps > 0 ? PROF_CONV : PROF_GEMM,
                   2.0 * (double)a->M * (double)a->N *
                       ((a->conv_taps > 0 ? (double)a->conv_taps * a->C : (double)a->K1) + (double)a->K2),
                   a->M, a->N, a->K1 + a->K2);
    // the programmatic edge needs a KERNEL as the previous stream operation: not after the row-statistics memset
    const bool pdl = use_pdl() && a->rowstats_out == nullptr && !profiling_on();
    if (pair) {
        if (BN == 256) return launch_gemm_pair<256>(tA1, tB1, tA2, tB2, p, stream, pdl);
        if (BN == 160) return launch_gemm_pair<160>(tA1, tB1, tA2, tB2, p, stream, pdl);
        return launch_gemm_pair<128>(tA1, tB1, tA2, tB2, p, stream, pdl);
    }
    if (BN == 256) return launch_gemm<256>(tA1, tB1, tA2, tB2, p, stream, pdl);
    if (BN == 128) return launch_gemm<128>(tA1, tB1, tA2, tB2, p, stream, pdl);
    return launch_gemm<64>(tA1, tB1, tA2, tB2, p, stream, pdl);
}
