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
    // stream-K (pair kernel): the k-block units of all tiles are dealt evenly to the CTA pairs; a tile cut between two
    // (or more) pairs is summed through `sk_ws` (fp32 partial accumulators, one slot per CTA) guarded by `sk_flags`.
    int sk;               // 1: hybrid stream-K (tiles >= sk_first are cut along K between pairs)
    int sk_first;         // first tile of the stream-K region (a multiple of the pair count)
    float* sk_ws;
    int* sk_flags;
    int tail_first;       // sk == 0: tiles >= tail_first (the last, partial wave) are cut into tail_split column slices
    int tail_split;
    int tma_out;          // 1: bf16 output leaves through shared memory + TMA store (cp.async.bulk.tensor ... global)
    int tma_res;          // 1: the residual chunk is TMA-loaded into the output staging buffer ahead of its use
    float* colstats_out;  // [M / colstats_rows, N, 2] per-image column (sum, sum of squares) for the next GroupNorm
    int colstats_rows;
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

// erf-GELU for the GEGLU epilogue: erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the bf16 output
// rounding) — two MUFU ops (rcp, ex2) and ~12 FMA-pipe ops instead of erff()'s ~30-instruction branchy expansion.
// The GEGLU GEMMs (8192x10240x1280, 32768x5120x640: 21 % of a teacher evaluation) evaluate it 128 times per thread
// per tile, which made their epilogue as long as the K = 640 main loop.
__device__ __forceinline__ float gelu_erf_fast(float x) {
    // x Phi(x) with Phi by A&S 26.2.17 (= 7.1.26 at x / sqrt2): t = 1 / (1 + p |x| / sqrt2), Phi(|x|) = 1 - poly(t) e^{-x^2/2}
    const float ax = fabsf(x);
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164190f, ax, 1.0f)));      // 0.3275911 / sqrt(2)
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752f));             // e^{-x^2/2}
    const float erf_abs = fmaf(-poly, e, 1.0f);                 // erf(|x| / sqrt2)
    return 0.5f * fmaf(ax, erf_abs, x);                         // 0.5 x (1 + sign(x) erf(|x| / sqrt2))
}

// Epilogue stage 1: r = acc[row, col0 .. col0+31] (fp32 bit patterns) -> v = final fp32 outputs (LayerNorm fold, bias,
// row vector, activation, gate, GEGLU, residual).  With GEGLU the 16 outputs are v[0..15] (output column col0/2 + j).
// Every lane of the warp runs it (rows >= M compute on zero accumulators and read nothing).
__device__ __forceinline__ void epilogue_values(const GemmKParams& p, int row, int col0, const uint32_t (&r)[32],
                                                const RowState& rs, float (&v)[32], bool skip_residual = false) {
    const bool row_ok = row < p.M;
    const int N = p.N;
    const bool full = (col0 + 32 <= N);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

    if (p.ln_stats != nullptr && p.bias != nullptr && full) {
        // LayerNorm(x) W^T + b' = rstd * acc + (-rstd * mean * colsum(W') + b'): two FMAs per element
        const float rstd = rs.ln_rstd, nmr = -rs.ln_mean * rs.ln_rstd;
        const float4* c4 = reinterpret_cast<const float4*>(p.ln_colsum + col0);
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 c = __ldg(c4 + j);
            const float4 b = __ldg(b4 + j);
            v[4 * j + 0] = fmaf(v[4 * j + 0], rstd, fmaf(nmr, c.x, b.x));
            v[4 * j + 1] = fmaf(v[4 * j + 1], rstd, fmaf(nmr, c.y, b.y));
            v[4 * j + 2] = fmaf(v[4 * j + 2], rstd, fmaf(nmr, c.z, b.z));
            v[4 * j + 3] = fmaf(v[4 * j + 3], rstd, fmaf(nmr, c.w, b.w));
        }
    } else {
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
    if (p.act == 1) {   // gelu (tanh approximation): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float x = v[j];
            const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
            v[j] = 0.5f * x * (1.0f + tanhf(u));
        }
    }
    if (p.act == 2) {   // ReLU (VGG16 feature stack of the LPIPS distillation loss)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.rowscale != nullptr && row_ok) {
        const float* sv = p.rowscale + (long long)(row / p.rows_per_group_scale) * p.ldrs + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (col0 + j < N) v[j] *= __ldg(sv + j);
    }
    if (p.geglu) {
        // interleaved packing: cols [0,16) value, [16,32) gate -> 16 outputs at col0/2  (N % 32 == 0: always full)
        const int oc0 = col0 >> 1;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = v[j] * gelu_erf_fast(v[16 + j]);
        if (p.residual != nullptr && row_ok && !skip_residual) {
            const bf16* rp = p.residual + (long long)row * p.ldr + oc0;
            if ((p.ldr & 7) == 0) {
                const uint4* r4 = reinterpret_cast<const uint4*>(rp);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint4 u = __ldg(r4 + j);
                    float2 f;
                    f = unpack_bf16x2(u.x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
                    f = unpack_bf16x2(u.y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
                    f = unpack_bf16x2(u.z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
                    f = unpack_bf16x2(u.w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += __bfloat162float(rp[j]);
            }
        }
        return;
    }
    if (p.residual != nullptr && row_ok && !skip_residual) {
        const bf16* rp = p.residual + (long long)row * p.ldr + col0;
        if (full && (p.ldr & 7) == 0) {
            const uint4* r4 = reinterpret_cast<const uint4*>(rp);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 u = __ldg(r4 + j);
                float2 f;
                f = unpack_bf16x2(u.x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
                f = unpack_bf16x2(u.y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
                f = unpack_bf16x2(u.z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
                f = unpack_bf16x2(u.w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (col0 + j < N) v[j] += __bfloat162float(rp[j]);
        }
    }
}

// Epilogue stage 2 (bf16 outputs): round to bf16 pairs, accumulate the row statistics of the ROUNDED values (what
// the next LayerNorm-folded GEMM will read), columns >= n_out count as zero.  nout = 32 (16 with GEGLU).
template <int NOUT>
__device__ __forceinline__ void epilogue_pack(const GemmKParams& p, bool row_ok, int oc0, int n_out, float (&v)[32],
                                              uint32_t (&pk)[16], RowState& rs) {
    if (oc0 + NOUT > n_out) {
#pragma unroll
        for (int j = 0; j < NOUT; ++j)
            if (oc0 + j >= n_out) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < NOUT / 2; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
    if (p.rowstats_out != nullptr && row_ok) {
#pragma unroll
        for (int j = 0; j < NOUT / 2; ++j) stats_of_bf16x2(pk[j], rs);
    }
}

// GroupNorm statistics from the producer: the warp holds a 32-row x 32-column block of ROUNDED outputs (lane = row,
// pk = 16 bf16 pairs).  A butterfly of 31 shuffles per quantity leaves lane l with the sum over the 32 rows of column
// col0 + l (stage `o` keeps the half of the columns whose bit `o` equals the lane's), then one 8-byte reduction per
// lane into colstats[image][col0 + lane] — the reduction pass of the following GroupNorm over the whole activation
// (gn_reduce_kernel: 16-33 us per call, 61 calls per SDXL evaluation) disappears.
__device__ __forceinline__ void colstats_accumulate(const GemmKParams& p, bool row_ok, int row_base, int col0,
                                                    const uint32_t (&pk)[16], int lane) {
    float s[32], q[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 f = unpack_bf16x2(pk[j]);
        const float a = row_ok ? f.x : 0.f, b = row_ok ? f.y : 0.f;
        s[2 * j] = a;
        s[2 * j + 1] = b;
        q[2 * j] = a * a;
        q[2 * j + 1] = b * b;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float ks = up ? s[i + o] : s[i], ss = up ? s[i] : s[i + o];
            const float kq = up ? q[i + o] : q[i], sq = up ? q[i] : q[i + o];
            s[i] = ks + __shfl_xor_sync(0xffffffffu, ss, o);
            q[i] = kq + __shfl_xor_sync(0xffffffffu, sq, o);
        }
    }
    const int img = row_base / p.colstats_rows;
    float2* dst = reinterpret_cast<float2*>(p.colstats_out) + (long long)img * p.N + col0 + lane;
    atomicAdd(dst, make_float2(s[0], q[0]));
}

// Direct (register -> global) store of one chunk; used by the single-CTA kernel and whenever the output is fp32 or
// its row stride is not a TMA-legal multiple of 16 bytes.
__device__ __forceinline__ void epilogue_chunk(const GemmKParams& p, int row, int col0, const uint32_t (&r)[32],
                                               RowState& rs) {
    if (col0 >= p.N) return;
    float v[32];
    epilogue_values(p, row, col0, r, rs, v);
    if (row >= p.M) return;
    const int n_out = p.geglu ? (p.N >> 1) : p.N;
    const int oc0 = p.geglu ? (col0 >> 1) : col0;
    const int nout = p.geglu ? 16 : 32;
    const bool full = oc0 + nout <= n_out;
    if (p.out_fp32) {
        float* orow = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + oc0;
        if (full && (p.ldo & 3) == 0) {
            float4* o4 = reinterpret_cast<float4*>(orow);
            if (p.geglu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nout && oc0 + j < n_out) orow[j] = v[j];
        }
        return;
    }
    uint32_t pk[16];
    if (p.geglu)
        epilogue_pack<16>(p, true, oc0, n_out, v, pk, rs);
    else
        epilogue_pack<32>(p, true, oc0, n_out, v, pk, rs);
    bf16* orow = reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldo + oc0;
    if (full && (p.ldo & 7) == 0) {
        uint4* o4 = reinterpret_cast<uint4*>(orow);
        o4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        o4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (!p.geglu) {
            o4[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
            o4[3] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j < nout && oc0 + j < n_out) {
                const uint32_t w = pk[j >> 1];
                const unsigned short h = (j & 1) ? (unsigned short)(w >> 16) : (unsigned short)(w & 0xffffu);
                reinterpret_cast<unsigned short*>(orow)[j] = h;
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
//
// STREAM-K.  The work of a launch is U = tiles x k-blocks units, dealt evenly and CONTIGUOUSLY to the P pairs: pair p
// owns units [p U / P, (p+1) U / P) — the tail of one tile, whole tiles, the head of another.  160 tiles on 74 pairs
// (8192 x 1280 x K) therefore cost 2.16 tile-times instead of 3 waves.  A segment that does not start at k-block 0
// is a PARTIAL: it is always the first thing its pair does, its raw fp32 accumulators go to the pair's workspace slot
// and a flag is released.  The segment that starts at k-block 0 FINISHES the tile: it is the last thing ITS pair does
// for that tile, so by then the partials of the following pair(s) have long been written; it acquires their flags,
// adds their slots and runs the normal epilogue.  All CTAs are co-resident (grid <= SM count, 1 CTA / SM), so the
// flag wait cannot dead-lock.  Flags are reset by the reader: the workspace stays all-zero between launches.
//
// EPILOGUE.  bf16 outputs leave through shared memory and the TMA store unit: every epilogue warp owns two 2 KB
// staging buffers (32 rows x 64 B, 64-byte swizzle = conflict-free 16-byte stores), writes a 32 x 32 chunk, fences
// the async proxy and one lane issues cp.async.bulk.tensor.2d.global.shared::cta; the tensor map clips rows >= M and
// columns >= N, so ragged edges need no predicates.  fp32 outputs / non-TMA-legal strides use direct stores.
template <int BN>
struct PairCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int BH_BYTES = (BN / 2) * BK * 2;
    static constexpr int STAGES = (BN == 256) ? 6 : (BN == 160 ? 7 : 8);
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;   // power of two >= 2 accumulator stages
    static constexpr int OUT_STAGE_BYTES = 8 * 2 * 2048;            // 8 epilogue warps x 2 buffers x (32 rows x 64 B)
    static constexpr int SMEM_BYTES = STAGES * (A_BYTES + BH_BYTES) + OUT_STAGE_BYTES + 512 + 1024;
};

__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tmap), "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // 8 epilogue warps

// first stream-K unit (k-block of the stream-K region) of pair q
__device__ __forceinline__ long long sk_begin(long long total_units, int num_pairs, int q) {
    return (total_units * q) / num_pairs;
}

// The sequence of work items of one CTA pair, identical in the producer, the MMA issuer and the epilogue warps:
//   * whole tiles dealt round-robin (item i of pair p is tile p + i P: the P pairs work on P CONSECUTIVE tiles of the
//     L2-friendly raster at any moment);
//   * sk == 0: the tiles of the last, partial wave are cut into `tail_split` column slices (sub, nsub) so that the
//     tail occupies ~all pairs for 1/nsub of a tile time instead of a few pairs for a whole one;
//   * sk == 1: the tiles from sk_first on are a stream-K region: its k-blocks are dealt evenly and contiguously, pair
//     p owning [p U / P, (p+1) U / P) — the tail of one tile, maybe a whole tile, the head of the next.
struct WorkIter {
    int pair_id, num_pairs, total_tiles, kb_total;
    int sk, sk_first, tail_first, tail_split;
    int i;
    long long u, u_end, sk_units;
    int tile, kb0, kb1, sub, nsub;
    __device__ __forceinline__ void init(const GemmKParams& p, int pair_id_, int num_pairs_, int total_tiles_,
                                         int kb_total_) {
        pair_id = pair_id_; num_pairs = num_pairs_; total_tiles = total_tiles_; kb_total = kb_total_;
        sk = p.sk; sk_first = p.sk_first; tail_first = p.tail_first; tail_split = p.tail_split;
        i = 0;
        sk_units = sk ? (long long)(total_tiles - sk_first) * kb_total : 0;
        u = sk ? sk_begin(sk_units, num_pairs, pair_id) : 0;
        u_end = sk ? sk_begin(sk_units, num_pairs, pair_id + 1) : 0;
    }
    __device__ __forceinline__ bool next() {
        const int item = pair_id + i * num_pairs;
        if (!sk) {
            const int n_items = tail_first + (total_tiles - tail_first) * tail_split;
            if (item >= n_items) return false;
            ++i;
            if (item < tail_first) {
                tile = item; sub = 0; nsub = 1;
            } else {
                const int t = item - tail_first;
                tile = tail_first + t / tail_split; sub = t % tail_split; nsub = tail_split;
            }
            kb0 = 0; kb1 = kb_total;
            return true;
        }
        if (item < sk_first) {
            ++i;
            tile = item; sub = 0; nsub = 1; kb0 = 0; kb1 = kb_total;
            return true;
        }
        if (u >= u_end) return false;
        const int t = (int)(u / kb_total);
        kb0 = (int)(u - (long long)t * kb_total);
        kb1 = (int)min((long long)kb_total, kb0 + (u_end - u));
        u += kb1 - kb0;
        tile = sk_first + t; sub = 0; nsub = 1;
        return true;
    }
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmB1s, const __grid_constant__ CUtensorMap tmB2s,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                 const GemmKParams p) {
    using Cfg = PairCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    uint8_t* sOut = sB + STAGES * Cfg::BH_BYTES;                       // 1024-byte aligned (stage sizes are)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + Cfg::OUT_STAGE_BYTES);
    uint64_t* full = bars;               // used in the leader CTA only
    uint64_t* empty = bars + STAGES;     // per CTA
    uint64_t* tfull = bars + 2 * STAGES; // per CTA
    uint64_t* tempty = tfull + 2;        // used in the leader CTA only
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);
    uint64_t* res_bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [8 warps][2 buffers]

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
        if (p.tma_out) tma_prefetch_desc(&tmOut);
        if (p.tma_res) tma_prefetch_desc(&tmRes);
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
        for (int a = 0; a < 16; ++a) mbar_init(&res_bars[a], 1);
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
    WorkIter wi;
    wi.init(p, pair_id, num_pairs, total_tiles, kb_total);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            while (wi.next()) {
                const int kb0 = wi.kb0, kb1 = wi.kb1;
                int mt2, nt;
                tile_coords(pp, wi.tile, mt2, nt);
                const int row0 = mt2 * 2 * BM + (int)rank * BM;
                int n0 = 0, h0 = 0, w0 = 0;
                if (p.conv_taps) {
                    const int hw = p.H * p.W;
                    n0 = row0 / hw;
                    const int rem = row0 - n0 * hw;
                    h0 = rem / p.W;
                    w0 = rem - h0 * p.W;
                }
                // column slice `sub` of `nsub`: this CTA stages (BN / nsub) / 2 rows of B through the slice tensor maps
                const int n_ext = BN / wi.nsub;
                const int brow = nt * BN + wi.sub * n_ext + (int)rank * (n_ext / 2);
                const bool sliced = wi.nsub > 1;
                const CUtensorMap* mB1 = sliced ? &tmB1s : &tmB1;
                const CUtensorMap* mB2 = sliced ? &tmB2s : &tmB2;
                const uint32_t tx_bytes = 2u * (uint32_t)(Cfg::A_BYTES + Cfg::BH_BYTES / wi.nsub);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), 0);
                    if (leader)
                        mbar_arrive_expect_tx(&full[stage], tx_bytes);
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
                        tma_load_2d_pair(mB1, full_leader, b_dst, kb * BK, brow);
                    } else {
                        const int k2 = kb - p.kb1;
                        tma_load_2d_pair(&tmA2, full_leader, a_dst, k2 * BK, row0);
                        tma_load_2d_pair(mB2, full_leader, b_dst, k2 * BK, brow);
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
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            while (wi.next()) {
                const int kb0 = wi.kb0, kb1 = wi.kb1;
                const uint32_t idesc = make_idesc_bf16(2 * BM, BN / wi.nsub);
                mbar_wait(&tempty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::BH_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc_mma_bf16_pair(d_tmem, make_desc_k_sw128(a_addr + k * 32),
                                         make_desc_k_sw128(b_addr + k * 32), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
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
        const int ew = warp - 4;          // epilogue warp 0..7
        const int q = ew & 3;             // TMEM lane quarter == warp_id % 4
        const int half = ew >> 2;         // which half of the tile's 32-column chunks this warp drains
        constexpr int NCH = BN / 32;
        uint8_t* my_stage = sOut + ew * 4096;
        uint32_t n_stored = 0;            // TMA stores issued by this warp (buffer = n_stored & 1)
        // RESIDUAL PREFETCH (tma_res): the residual chunk a store will need is TMA-loaded into that store's staging
        // buffer ahead of time — the first chunk of a tile before the wait for the accumulator (i.e. during the main
        // loop), chunk c+1 while chunk c is processed — so the epilogue never sits on a global-load latency (ncu r02:
        // 12-25 % of the samples of the residual GEMMs were long-scoreboard stalls on exactly that load).
        uint64_t* my_res = res_bars + ew * 2;
        uint32_t res_phase = 0;           // bit b: parity to wait for on my_res[b]
        auto issue_res = [&](uint32_t k, int col0, int row0) {     // lane 0: residual of stored-chunk number k
            const uint32_t b = k & 1u;
            mbar_arrive_expect_tx(&my_res[b], 2048);
            tma_load_2d(&tmRes, &my_res[b], my_stage + b * 2048, col0, row0);
        };
        const int n_out = p.geglu ? (p.N >> 1) : p.N;
        float* ws_mine = p.sk_ws + ((size_t)pair_id * 2 + rank) * (size_t)(BM * BN);
        int acc = 0;
        uint32_t acc_phase = 0;
        while (wi.next()) {
            const int tile = wi.tile, kb0 = wi.kb0, kb1 = wi.kb1;
            int mt2, nt;
            tile_coords(pp, tile, mt2, nt);
            // a column slice has BN / nsub columns; its 32-column chunks are split between the two warp halves
            const int nch = NCH / wi.nsub, split = (nch + 1) / 2;
            const int c_lo = half ? split : 0, c_hi = half ? nch : split;
            const int col_base = nt * BN + wi.sub * (BN / wi.nsub);
            const int row0 = mt2 * 2 * BM + (int)rank * BM + q * 32;
            if (p.tma_res && kb0 == 0 && c_lo < c_hi && col_base + c_lo * 32 < p.N) {
                if (lane == 0) {
                    tma_store_wait_read<1>();                    // the store that last used this buffer has read it
                    issue_res(n_stored, col_base + c_lo * 32, row0);
                }
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            if (kb0 != 0) {
                // PARTIAL segment: raw accumulators -> this CTA's workspace slot ([chunk][quarter][reg][lane]: every
                // warp-wide store / later load is one contiguous 128-byte line), then release the flag.
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_addr + c * 32, r);
                    tmem_ld_wait();
                    float* dst = ws_mine + ((size_t)(c * 4 + q) * 32) * 32 + lane;
#pragma unroll
                    for (int j = 0; j < 32; ++j) __stcg(dst + j * 32, __uint_as_float(r[j]));
                }
                __threadfence();
                epi_bar_sync();
                if (ew == 0 && lane == 0) st_release_gpu(p.sk_flags + pair_id * 2 + rank, 1);
            } else {
                // FINISHER (or a whole tile).  Partials of this tile live in the slots of the following pairs.
                int n_part = 0;
                if (kb1 < kb_total) {
                    const long long tile_end = (long long)(tile - wi.sk_first + 1) * kb_total;
                    while (pair_id + 1 + n_part < num_pairs &&
                           sk_begin(wi.sk_units, num_pairs, pair_id + 1 + n_part) < tile_end)
                        ++n_part;
                    if (ew == 0 && lane == 0) {
                        for (int i = 0; i < n_part; ++i) {
                            int* f = p.sk_flags + (pair_id + 1 + i) * 2 + rank;
                            uint32_t spins = 0;
                            while (ld_acquire_gpu(f) == 0) {
                                if (++spins > FD_SPIN_LIMIT) {
                                    printf("fd: stream-K flag timeout pair %d waits for %d\n", pair_id, pair_id + 1 + i);
                                    __trap();
                                }
                            }
                            *f = 0;       // reader resets: the workspace is all-zero again when the launch ends
                        }
                    }
                    epi_bar_sync();
                }
                const int row = mt2 * 2 * BM + (int)rank * BM + q * 32 + lane;
                const bool row_ok = row < p.M;
                RowState rs;
                row_state_init(p, row, rs);
#pragma unroll 1
                for (int c = c_lo; c < c_hi; ++c) {
                    const int col0 = col_base + c * 32;
                    if (col0 >= p.N) break;                       // warp-uniform
                    uint32_t r[32];
                    tmem_ld_32x32(t_addr + c * 32, r);
                    tmem_ld_wait();
                    for (int i = 0; i < n_part; ++i) {
                        const float* src = p.sk_ws + ((size_t)(pair_id + 1 + i) * 2 + rank) * (size_t)(BM * BN) +
                                           ((size_t)(c * 4 + q) * 32) * 32 + lane;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldcg(src + j * 32));
                    }
                    if (!p.tma_out) {
                        epilogue_chunk(p, row, col0, r, rs);
                        continue;
                    }
                    if (p.tma_res && lane == 0 && c + 1 < c_hi && col0 + 32 < p.N) {
                        tma_store_wait_read<0>();                 // store n_stored-1 has read the other buffer
                        issue_res(n_stored + 1, col0 + 32, row0);
                    }
                    float v[32];
                    epilogue_values(p, row, col0, r, rs, v, p.tma_res != 0);
                    uint32_t pk[16];
                    uint8_t* buf = my_stage + (n_stored & 1u) * 2048;
                    if (p.tma_res) {
                        const uint32_t b = n_stored & 1u;
                        mbar_wait(&my_res[b], (res_phase >> b) & 1u);
                        res_phase ^= 1u << b;
                        const uint8_t* src = buf + lane * 64;
                        const int sw = (lane >> 1) & 3;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 u = *reinterpret_cast<const uint4*>(src + ((j ^ sw) << 4));
                            float2 f;
                            f = unpack_bf16x2(u.x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
                            f = unpack_bf16x2(u.y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
                            f = unpack_bf16x2(u.z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
                            f = unpack_bf16x2(u.w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
                        }
                        __syncwarp();                             // every lane has read before anyone overwrites
                    } else if (n_stored >= 2) {                   // the store issued from this buffer has read it
                        if (lane == 0) tma_store_wait_read<1>();
                        __syncwarp();
                    }
                    if (p.geglu) {
                        const int oc0 = col0 >> 1;
                        epilogue_pack<16>(p, row_ok, oc0, n_out, v, pk, rs);
                        // 32 rows x 32 B, 32-byte swizzle: 16-byte chunk index ^= bit 7 of the address = (row >> 2) & 1
                        uint8_t* dst = buf + lane * 32;
                        const int sw = (lane >> 2) & 1;
                        *reinterpret_cast<uint4*>(dst + ((0 ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(dst + ((1 ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmOut, buf, oc0, row - lane);
                            tma_store_commit();
                        }
                    } else {
                        epilogue_pack<32>(p, row_ok, col0, n_out, v, pk, rs);
                        // 32 rows x 64 B, 64-byte swizzle: chunk index ^= address bits [7,9) = (row >> 1) & 3
                        uint8_t* dst = buf + lane * 64;
                        const int sw = (lane >> 1) & 3;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<uint4*>(dst + ((j ^ sw) << 4)) =
                                make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmOut, buf, col0, row - lane);
                            tma_store_commit();
                        }
                        if (p.colstats_out != nullptr && row - lane < p.M)
                            colstats_accumulate(p, row_ok, row - lane, col0, pk, lane);
                    }
                    ++n_stored;
                }
                row_state_flush(p, row, rs);
            }
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
        if (lane == 0) tma_store_wait_all();     // shared memory must outlive the bulk stores that read it
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ host

// Tile configuration: BN in {64,128,256} single-CTA (128-row tiles) or {128,160,256} CTA-pair (256-row tiles,
// returned as 512 + BN).  Cost = waves x (tile rows x BN) / efficiency, efficiencies from the smem-traffic
// model calibrated on B200.  With a stream-K workspace the pair kernels pay FRACTIONAL waves.
static int choose_bn(int M, int N, int force, bool streamk) {
    force &= 1023;      // bits 1024 / 2048 of force_bn select the work split (see fd_gemm), not the tile
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
        double waves = (double)((tiles + units - 1) / units);
        (void)streamk;
        if (c.rows == 256 && tiles > units && tiles % units != 0 && (tiles % units) * 2 <= units)
            waves -= 0.5;                         // the partial wave runs as column slices (~half a tile time)
        const double cost = waves * c.rows * c.bn / c.eff / (c.rows == 256 ? 2.0 : 1.0);
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
                            const CUtensorMap& tB2, const CUtensorMap& tB1s, const CUtensorMap& tB2s,
                            const CUtensorMap& tOut, const CUtensorMap& tRes, const GemmKParams& p, int pairs,
                            cudaStream_t stream, bool pdl) {
    using Cfg = PairCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        FD_CHECK_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg::SMEM_BYTES));
        attr_set = true;
    }
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
    FD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_pair_kernel<BN>, tA1, tB1, tA2, tB2, tB1s, tB2s, tOut, tRes, p));
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

extern "C" size_t fd_gemm_workspace_bytes(void) {
    // 4 KB of flags + one 128 x 256 fp32 partial-accumulator slot per CTA of the persistent pair grid
    return 4096 + (size_t)fd::num_sms() * (size_t)(fd::BM * 256) * sizeof(float);
}

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

    // stream-K needs the caller's workspace (fd_gemm_workspace_bytes) and 16-byte alignment of it
    static const bool env_no_sk = getenv("FD_NO_STREAMK") != nullptr;
    static const bool env_no_tma_store = getenv("FD_NO_TMA_STORE") != nullptr;
    const bool ws_ok = a->workspace != nullptr && a->workspace_bytes >= (long long)fd_gemm_workspace_bytes() &&
                       ((uintptr_t)a->workspace & 15) == 0 && !env_no_sk;
    int code = choose_bn(a->M, a->N, a->force_bn, ws_ok);
    if (a->colstats_out != nullptr) {
        // the column statistics live in the TMA-store epilogue of the CTA-pair kernel only
        FD_CHECK_ARG(!a->out_fp32 && !a->geglu && a->N % 32 == 0 && a->colstats_rows > 0 && a->colstats_rows % 32 == 0 &&
                         a->M % a->colstats_rows == 0 && (a->ldo % 8) == 0 && ((uintptr_t)a->out & 15) == 0 &&
                         !env_no_tma_store,
                     "fd_gemm: colstats_out needs a bf16 TMA-storable output, N %% 32 == 0, rows per image %% 32 == 0");
        if (code < 512) code = 512 + 128;
    }
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
    p.colstats_out = a->colstats_out;
    p.colstats_rows = a->colstats_rows;
    p.act = a->act;
    p.rowscale = a->rowscale;
    p.rows_per_group_scale = a->rows_per_group_scale;
    p.ldrs = a->ldrs > 0 ? a->ldrs : a->N;
    FD_CHECK_ARG(!a->rowscale || a->rows_per_group_scale > 0, "fd_gemm: rowscale needs rows_per_group_scale");
    FD_CHECK_ARG(!a->ln_stats || (a->ln_colsum && a->K2 == 0), "fd_gemm: LayerNorm fold needs ln_colsum and no K2 segment");
    FD_CHECK_ARG(!a->rowstats_out || !a->out_fp32, "fd_gemm: rowstats_out needs a bf16 output");
    if (a->rowstats_out && !a->rowstats_prezeroed)
        FD_CHECK_CUDA(cudaMemsetAsync(a->rowstats_out, 0, sizeof(float) * 2 * (size_t)a->M, stream));

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
    // pair kernel: work split (tail column slices or hybrid stream-K) and TMA-store epilogue
    CUtensorMap tOut = tA1, tRes = tA1, tB1s = tB1, tB2s = tB2;
    int pairs = 0;
    if (pair) {
        const int tiles = ((a->M + 2 * BM - 1) / (2 * BM)) * p.num_n_tiles;
        const int kb_total = p.kb1 + p.kb2;
        const int max_pairs = num_sms() / 2;
        pairs = tiles < max_pairs ? tiles : max_pairs;
        p.tail_first = tiles;
        p.tail_split = 1;
        const int rem = tiles % max_pairs;               // tiles of the last, partial wave
        static const int env_tail = getenv("FD_TAIL_SPLIT") ? atoi(getenv("FD_TAIL_SPLIT")) : -1;   // 0 = off
        static const int env_sk_min_ = getenv("FD_SK_MIN") ? atoi(getenv("FD_SK_MIN")) : 24;
        // force_bn | 1024: stream-K whenever it is legal (tests); force_bn | 2048: whole tiles only, no slices
        const int env_sk_min = (a->force_bn & 1024) ? 0 : env_sk_min_;
        const bool plain = (a->force_bn & 2048) != 0;
        // (1) hybrid stream-K pays for its partial-tile exchange (~4 us on the kernel's tail) only when the idle part of
        //     the last wave is long: (1 - rem/P) tile-times of kb_total k-blocks each, in units of BN = 256 k-blocks
        if (!plain && ws_ok && tiles > max_pairs && rem != 0 && kb_total >= 8 &&
            (double)(max_pairs - rem) / max_pairs * kb_total * BN / 256.0 >= (double)env_sk_min) {
            p.sk = 1;
            p.sk_first = (tiles / max_pairs - 1) * max_pairs;      // last full wave + the partial wave are streamed
            p.sk_flags = reinterpret_cast<int*>(a->workspace);
            p.sk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->workspace) + 4096);
        } else if (!plain && rem != 0 && env_tail != 0) {
            // (2) otherwise cut the tiles of the partial wave into column slices (multiples of 32 columns) while they
            //     still fit one wave: the tail then lasts ~1/split of a tile time
            int split = 1;
            for (int sp = 2; sp <= 4; sp *= 2)
                if (BN % (32 * sp) == 0 && (long long)rem * sp <= max_pairs) split = sp;
            if (env_tail > 0 && BN % (32 * env_tail) == 0 && (long long)rem * env_tail <= max_pairs) split = env_tail;
            if (split > 1) {
                p.tail_first = tiles - rem;
                p.tail_split = split;
                if (tiles < max_pairs) pairs = rem * split;
                const uint32_t rows = (uint32_t)(BN / split / 2);
                {
                    const uint64_t dims[2] = {(uint64_t)a->K1, (uint64_t)a->N};
                    const uint64_t str[1] = {(uint64_t)a->ldb1 * 2};
                    const uint32_t box[2] = {(uint32_t)BK, rows};
                    rc = encode_tmap_bf16(&tB1s, a->b1, 2, dims, str, box);
                    if (rc) return rc;
                }
                if (a->K2 > 0) {
                    const uint64_t dims[2] = {(uint64_t)a->K2, (uint64_t)a->N};
                    const uint64_t str[1] = {(uint64_t)a->ldb2 * 2};
                    const uint32_t box[2] = {(uint32_t)BK, rows};
                    rc = encode_tmap_bf16(&tB2s, a->b2, 2, dims, str, box);
                    if (rc) return rc;
                } else {
                    tB2s = tB1s;
                }
            }
        }
        const int n_out = a->geglu ? a->N / 2 : a->N;
        if (!a->out_fp32 && !env_no_tma_store && (a->ldo % 8) == 0 && ((uintptr_t)a->out & 15) == 0) {
            const uint64_t dims[2] = {(uint64_t)n_out, (uint64_t)a->M};
            const uint64_t str[1] = {(uint64_t)a->ldo * 2};
            const uint32_t box[2] = {a->geglu ? 16u : 32u, 32u};
            rc = encode_tmap_bf16_sw(&tOut, a->out, 2, dims, str, box, a->geglu ? 32 : 64);
            if (rc) return rc;
            p.tma_out = 1;
            static const bool env_no_tma_res = getenv("FD_NO_TMA_RES") != nullptr;
            if (a->residual && !a->geglu && !env_no_tma_res && (a->ldr % 8) == 0 && ((uintptr_t)a->residual & 15) == 0) {
                const uint64_t rstr[1] = {(uint64_t)a->ldr * 2};
                rc = encode_tmap_bf16_sw(&tRes, a->residual, 2, dims, rstr, box, 64);
                if (rc) return rc;
                p.tma_res = 1;
            }
        }
    }
    ProfScope prof(stream, a->conv_taps > 0 ? PROF_CONV : PROF_GEMM,
                   2.0 * (double)a->M * (double)a->N *
                       ((a->conv_taps > 0 ? (double)a->conv_taps * a->C : (double)a->K1) + (double)a->K2),
                   a->M, a->N, a->K1 + a->K2);
    // the programmatic edge needs a KERNEL as the previous stream operation: not after the row-statistics memset
    const bool pdl = use_pdl() && a->rowstats_out == nullptr && !profiling_on();
    if (pair) {
        if (BN == 256) return launch_gemm_pair<256>(tA1, tB1, tA2, tB2, tB1s, tB2s, tOut, tRes, p, pairs, stream, pdl);
        if (BN == 160) return launch_gemm_pair<160>(tA1, tB1, tA2, tB2, tB1s, tB2s, tOut, tRes, p, pairs, stream, pdl);
        return launch_gemm_pair<128>(tA1, tB1, tA2, tB2, tB1s, tB2s, tOut, tRes, p, pairs, stream, pdl);
    }
    if (BN == 256) return launch_gemm<256>(tA1, tB1, tA2, tB2, p, stream, pdl);
    if (BN == 128) return launch_gemm<128>(tA1, tB1, tA2, tB2, p, stream, pdl);
    return launch_gemm<64>(tA1, tB1, tA2, tB2, p, stream, pdl);
}
