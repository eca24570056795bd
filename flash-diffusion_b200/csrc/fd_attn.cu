// fd_attn.cu — fused FlashAttention-style forward for sm_100a (head dim 64, bf16, fp32 softmax).
//
//   O = softmax(Q K^T * scale) V       per (batch, head), Q/K/V read in place from [B, N, H*64]
//
// One CTA per (256 queries = two 128-row tiles, head, batch); 18 warps:
//   warp 0    TMA producer  : Q0,Q1 once, K/V tiles (128 keys) through a 3-stage ring
//   warp 1    MMA issuer    : S_w = Q_w K^T (tcgen05 128x128x64) and O_w += P_w V (128x64x128) for w = 0,1,
//                             interleaved so that the softmax of one tile overlaps the MMAs of the other
//   warps 2-17 softmax      : per query tile 8 warps = 2 column halves x 4 TMEM lane quarters (a thread owns 64 of
//                             its row's 128 scores; the halves exchange the row max through smem); the scores go
//                             to registers in one tcgen05.ld round trip and the S buffer is released at once
//                             (S(j+1) is issued while softmax(j) runs);
//                             FMNMX3 row max, FFMA2 + ex2.approx + FADD2, P -> smem (bf16, 128B-swizzled K-major);
//                             O rescaled in TMEM only when the running maximum grew by more than 2^8
// TMEM: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512) (P = bf16 pairs, the A
// operand of O += P V is read from TMEM, so P never touches shared memory).  V is an MN-major B operand straight
// from its row-major [key, d] tile, so no transpose is materialised.
//
// UPSTREAM math: diffusers Attention + AttnProcessor2_0 -> F.scaled_dot_product_attention
// (SURVEY.md §2.2); reference call path src/flash/models/unets/unet.py:108-119.
#include <stdlib.h>

#include "fd_common.cuh"
#include "fd_host.h"

#ifndef FD_ATTN_P_TMEM
#define FD_ATTN_P_TMEM 1   // P operand of the P V product lives in TMEM (tcgen05.mma A-from-TMEM) instead of smem
#endif

namespace fd {

constexpr int ATT_BM = 256;   // queries per CTA: two 128-row tiles, one softmax warpgroup each
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_STAGES = 3;
constexpr int ATT_THREADS = 64 + 512;         // TMA warp, MMA warp, 2 tiles x 2 column halves x 4 softmax warps
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB
// Q0,Q1 | K,V x stages | P0,P1 (2 sub-tiles each) | barriers
constexpr int ATT_SMEM = 2 * ATT_TILE_BYTES + 2 * ATT_STAGES * ATT_TILE_BYTES + 4 * ATT_TILE_BYTES + 256 + 4096 + 1024;

struct AttnKParams {
    int Nq, Nkv;
    float scale_log2;
    bf16* o;
    long long ldo, o_batch_stride;
    float* lse;  // [B,H,Nq] or null
    int H;
    int bar_all;  // attn_fwd1: 1 = one 256-thread barrier for the row-maximum exchange (r02 A/B, FD_ATTN_BAR256)
    int diag;     // attn_fwd1: FD_ATTN_DIAG bits — timing diagnostic that BREAKS the result (bit 0: no row-maximum exchange)
    int spin;     // attn_fwd1: 1 = the S / P ping-pong waits poll with test_wait instead of suspending in try_wait (FD_ATTN_SPIN)
    int early;    // attn_fwd1: 1 = first TMA loads issued before the TMEM allocation / block sync (FD_ATTN_EARLY=0: off)
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

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2): halves the issue slots of the softmax inner loop
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    uint64_t ra, rb, rc, rd;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    uint64_t ra, rb, rd;
    asm("mov.b64 %0, {%1,%2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

// exp2 on the FMA pipe (Cody-Waite range reduction + cubic), two values at a time.  The MUFU (XU) pipe issues only
// 16 ex2 per clock per SM and is the softmax bottleneck (ncu: mio_throttle is the top stall), so every
// FD_ATTN_POLY_MOD-th pair of scores takes this path instead.  Relative error < 7e-4 (bf16 P has 2^-9 = 2e-3).
#ifndef FD_ATTN_POLY_MOD
#define FD_ATTN_POLY_MOD 4
#endif
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
    const float kMagic = 12582912.0f;   // 1.5 * 2^23
    x.x = fmaxf(x.x, -126.0f);
    x.y = fmaxf(x.y, -126.0f);
    const float2 t = fadd2(x, make_float2(kMagic, kMagic));
    const float2 n = fadd2(t, make_float2(-kMagic, -kMagic));
    const float2 f = fadd2(x, make_float2(-n.x, -n.y));
    float2 pz = ffma2(f, make_float2(0.05550411f, 0.05550411f), make_float2(0.24022651f, 0.24022651f));
    pz = ffma2(pz, f, make_float2(0.69314718f, 0.69314718f));
    pz = ffma2(pz, f, make_float2(1.0f, 1.0f));
    float2 r;
    r.x = __int_as_float(__float_as_int(pz.x) + (__float_as_int(t.x) << 23));
    r.y = __int_as_float(__float_as_int(pz.y) + (__float_as_int(t.y) << 23));
    return r;
}

// Rescale threshold (log2 domain): O and l are only rescaled when the running row maximum grew by more
// than this; otherwise the stale maximum keeps being used (P <= 2^8, exact in fp32 accumulation, and the
// final O / l is invariant to the reference maximum).
constexpr float ATT_RESCALE_THRESHOLD = 8.0f;

// POLY: a share of the exponentials runs on the FMA pipe (cubic, relative error < 7e-4, always an UNDER-estimate).
// That is fine for an inference-only forward, but when the log-sum-exp is kept for a backward the row sum l must be
// the sum of the very probabilities the backward recomputes (ex2.approx of S - LSE): with the polynomial share the
// recomputed row sums were off by ~1e-4, which is multiplied by the COMMON-MODE of dP and of the keys in
// dQ = sum_j dS_j K_j and showed up as cos 0.979 (instead of 0.998) on the q/k LoRA gradients of deep, near-uniform
// attention blocks at SDXL size (tests/test_sdxl_parity_gpu.py).  So the training forward (lse != NULL) is all-MUFU.
template <bool POLY>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnKParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sQ = smem;                                   // 2 tiles
    uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;                // ATT_STAGES tiles
    uint8_t* sV = sK + ATT_STAGES * ATT_TILE_BYTES;       // ATT_STAGES tiles
    uint8_t* sP = sV + ATT_STAGES * ATT_TILE_BYTES;       // 2 x (2 sub-tiles of [128][64])
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * ATT_TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;                  // [ATT_STAGES]
    uint64_t* kv_empty = kv_full + ATT_STAGES;     // [ATT_STAGES]
    uint64_t* s_full = kv_empty + ATT_STAGES;      // [2]  S_w(j) written by the tensor core
    uint64_t* s_free = s_full + 2;                 // [2]  S_w(j) copied to registers: S_w(j+1) may be issued
    uint64_t* p_ready = s_free + 2;                // [2]  P_w(j) in smem
    uint64_t* pv_done = p_ready + 2;               // [2]  O_w += P_w(j) V(j) complete
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_done + 2);
    float* mx_buf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [tile][parity][half][128]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_blk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int n_kv_tiles = (p.Nkv + ATT_BN - 1) / ATT_BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int w = 0; w < 2; ++w) {
            mbar_init(&s_full[w], 1);
            mbar_init(&s_free[w], 256);
            mbar_init(&p_ready[w], 256);
            mbar_init(&pv_done[w], 1);
        }
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
    // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
            tma_load_3d(&tmQ, q_full, sQ, head * ATT_D, q_blk * ATT_BM, batch);
            tma_load_3d(&tmQ, q_full, sQ + ATT_TILE_BYTES, head * ATT_D, q_blk * ATT_BM + 128, batch);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < n_kv_tiles; ++j) {
                mbar_wait(&kv_empty[st], ph ^ 1u);
                mbar_arrive_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
                tma_load_3d(&tmK, &kv_full[st], sK + st * ATT_TILE_BYTES, head * ATT_D, j * ATT_BN, batch);
                tma_load_3d(&tmV, &kv_full[st], sV + st * ATT_TILE_BYTES, head * ATT_D, j * ATT_BN, batch);
                if (++st == ATT_STAGES) {
                    st = 0;
                    ph ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_BN, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, 0, 1);  // B (=V) MN-major
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t p_addr = smem_u32(sP);
            auto issue_s = [&](int w, int st) {
                const uint32_t k_addr = smem_u32(sK + st * ATT_TILE_BYTES);
                const uint32_t qa = q_addr + w * ATT_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k)
                    tc_mma_bf16(tmem_base + w * 128, make_desc_k_sw128(qa + k * 32),
                                make_desc_k_sw128(k_addr + k * 32), idesc_qk, k != 0 ? 1u : 0u);
                tc_commit(&s_full[w]);
            };
            auto issue_pv = [&](int w, int st, int j) {
                const uint32_t v_addr = smem_u32(sV + st * ATT_TILE_BYTES);
#if FD_ATTN_P_TMEM
                // A operand (P, bf16) straight from TMEM: 8 columns (= 16 keys) per MMA
#pragma unroll
                for (int k = 0; k < ATT_BN / 16; ++k)
                    tc_mma_bf16_ts(tmem_base + 256 + w * 64, tmem_base + 384 + w * 64 + k * 8,
                                   make_desc_mn_sw128(v_addr + k * 2048, 0, 1024), idesc_pv, (j | k) != 0 ? 1u : 0u);
#else
                const uint32_t pa = p_addr + w * 2 * ATT_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < ATT_BN / 16; ++k)
                    tc_mma_bf16(tmem_base + 256 + w * 64,
                                make_desc_k_sw128(pa + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32),
                                make_desc_mn_sw128(v_addr + k * 2048, 0, 1024), idesc_pv, (j | k) != 0 ? 1u : 0u);
#endif
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);
            issue_s(1, 0);
            // Event-driven issue: per query tile w, S_w(j+1) may go as soon as S_w(j) was copied to registers
            // (s_free) and K(j+1) has landed; P_w(j) V(j) as soon as P_w(j) is in smem (p_ready).  Whichever
            // event fires first is served first, so a late warpgroup never delays the other one.
            int next_s[2] = {1, 1};      // next S tile to issue
            int next_pv[2] = {0, 0};     // next P V tile to issue
            int kv_seen = 1;             // number of K/V tiles known to have landed
            while (next_pv[0] < n_kv_tiles || next_pv[1] < n_kv_tiles) {
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int js = next_s[w];
                    if (js < n_kv_tiles) {
                        // needs: s_free[w](js-1) and kv_full(js)
                        if (mbar_test_wait(&s_free[w], (js - 1) & 1)) {
                            bool kv_ok = js < kv_seen;
                            if (!kv_ok && js == kv_seen) {
                                const int stg = js % ATT_STAGES;
                                if (mbar_test_wait(&kv_full[stg], (js / ATT_STAGES) & 1)) {
                                    kv_seen = js + 1;
                                    kv_ok = true;
                                }
                            }
                            if (kv_ok) {
                                tc_fence_after();
                                issue_s(w, js % ATT_STAGES);
                                next_s[w] = js + 1;
                            }
                        }
                    }
                    const int jp = next_pv[w];
                    if (jp < n_kv_tiles && mbar_test_wait(&p_ready[w], jp & 1)) {
                        tc_fence_after();
                        issue_pv(w, jp % ATT_STAGES, jp);
                        tc_commit(&pv_done[w]);
                        next_pv[w] = jp + 1;
                        // the K/V stage of tile jp is free once BOTH query tiles have issued their P V on it
                        if (next_pv[w ^ 1] > jp) tc_commit(&kv_empty[jp % ATT_STAGES]);
                    }
                }
            }
        }
    } else {
        // 16 softmax warps: query tile w (0/1) x column half h (0/1) x TMEM lane quarter (= warp % 4).
        // Thread (w, h, row) owns 64 of the 128 score columns of its row: half the serial work per thread and
        // twice the warps per scheduler to hide the TMEM / MUFU / barrier latencies.  The two halves of a row
        // exchange their partial row maximum through shared memory once per key tile.
        const int g = (warp - 2) >> 2;
        const int w = g >> 1, h = g & 1;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;                       // row within this 128-query tile
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const uint32_t tmem_S = tmem_base + w * 128 + h * 64;
        const uint32_t tmem_O = tmem_base + 256 + w * 64 + h * 32;
#if !FD_ATTN_P_TMEM
        uint8_t* sPw = sP + w * 2 * ATT_TILE_BYTES + h * ATT_TILE_BYTES;   // this half's [128][64] sub-tile
#endif
        float* mxw = mx_buf + w * 512;
        const int bar_id = 1 + w;
        float m_used = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_kv_tiles; ++j) {
            mbar_wait(&s_full[w], j & 1);
            tc_fence_after();
            const int kv_valid = min(ATT_BN, p.Nkv - j * ATT_BN) - h * 64;   // valid columns within this half
            uint32_t sr[2][32];
            tmem_ld_32x32(tmem_S + lane_base, sr[0]);
            tmem_ld_32x32(tmem_S + lane_base + 32, sr[1]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&s_free[w]);
            if (kv_valid < 64) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i >= kv_valid) sr[c][i] = 0xff800000u;      // -inf: exp2 -> 0
            }
            float mxs[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < 32; i += 2)
                    mxs[c] = max3(mxs[c], __uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1]));
            float mx = fmaxf(mxs[0], mxs[1]);
            float* slot = mxw + (j & 1) * 256;
            slot[h * 128 + row] = mx;
            asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
            mx = fmaxf(mx, slot[(h ^ 1) * 128 + row]);
            const float m_new = mx * p.scale_log2;
            const bool grow = m_new > m_used + ATT_RESCALE_THRESHOLD;      // first tile: m_used = -inf -> true
            const float alpha = (grow && j > 0) ? fast_exp2(m_used - m_new) : 1.0f;
            if (grow) m_used = m_new;
            l_run *= alpha;
            const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
            const float2 nm2 = make_float2(-m_used, -m_used);
            float2 ps2 = make_float2(0.f, 0.f);
            uint32_t pk[2][16];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float2 x = ffma2(make_float2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), sc2, nm2);
                    float2 e;
                    if (POLY && FD_ATTN_POLY_MOD > 0 && ((i >> 1) % (FD_ATTN_POLY_MOD > 0 ? FD_ATTN_POLY_MOD : 1)) == 1)
                        e = exp2_poly2(x);
                    else if (p.diag & 2)    // diag bit 1 (TIMING DIAGNOSTIC ONLY): no MUFU, one FMA per element
                        e = ffma2(x, make_float2(0.001f, 0.001f), make_float2(1.f, 1.f));
                    else
                        e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                    ps2 = fadd2(ps2, e);
                    pk[c][i >> 1] = pack_bf16x2(e.x, e.y);
                }
            }
            l_run += ps2.x + ps2.y;
            // P_w(j-1) V(j-1) must have completed before O is rescaled or the P buffer is overwritten
            if (j > 0) {
                mbar_wait(&pv_done[w], (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_O + lane_base, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                    tmem_st_32x32(tmem_O + lane_base, r);
                    tmem_st_wait();
                }
            }
#if FD_ATTN_P_TMEM
            // P (bf16 pairs) goes to its own TMEM columns [384 + 64 w + 32 h, +32): no shared-memory round trip
            {
                uint32_t pflat[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    pflat[i] = pk[0][i];
                    pflat[16 + i] = pk[1][i];
                }
                tmem_st_32x32(tmem_base + 384 + w * 64 + h * 32 + lane_base, pflat);
                tmem_st_wait();
            }
#else
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint8_t* sub = sPw + row * 128;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int chunk = (c * 4 + q4) ^ (row & 7);
                    *reinterpret_cast<uint4*>(sub + chunk * 16) =
                        make_uint4(pk[c][4 * q4], pk[c][4 * q4 + 1], pk[c][4 * q4 + 2], pk[c][4 * q4 + 3]);
                }
            }
            fence_proxy_async();
#endif
            tc_fence_before();
            mbar_arrive(&p_ready[w]);
        }
        // epilogue: combine the two halves' row sums, then O / l -> bf16 (each half stores 32 of the 64 channels)
        float* slot = mxw + (n_kv_tiles & 1) * 256;
        slot[h * 128 + row] = l_run;
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
        l_run += slot[(h ^ 1) * 128 + row];
        mbar_wait(&pv_done[w], (n_kv_tiles - 1) & 1);
        tc_fence_after();
        const int q_row = q_blk * ATT_BM + w * 128 + row;
        const float inv_l = 1.f / l_run;
        bf16* orow = p.o + (long long)batch * p.o_batch_stride + (long long)q_row * p.ldo + head * ATT_D + h * 32;
        {
            uint32_t r[32];
            tmem_ld_32x32(tmem_O + lane_base, r);
            tmem_ld_wait();
            if (q_row < p.Nq) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(r[8 * q4 + 0]) * inv_l, __uint_as_float(r[8 * q4 + 1]) * inv_l);
                    u.y = pack_bf16x2(__uint_as_float(r[8 * q4 + 2]) * inv_l, __uint_as_float(r[8 * q4 + 3]) * inv_l);
                    u.z = pack_bf16x2(__uint_as_float(r[8 * q4 + 4]) * inv_l, __uint_as_float(r[8 * q4 + 5]) * inv_l);
                    u.w = pack_bf16x2(__uint_as_float(r[8 * q4 + 6]) * inv_l, __uint_as_float(r[8 * q4 + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + q4 * 8) = u;
                }
            }
        }
        if (h == 0 && p.lse != nullptr && q_row < p.Nq)
            p.lse[((long long)batch * p.H + head) * p.Nq + q_row] = (m_used + log2f(l_run)) * 0.69314718055994531f;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------ 2 CTAs / SM variant
// One 128-query tile per CTA (10 warps, <= 96 registers, 82 KB shared memory, 256 TMEM columns) so that TWO CTAs are
// resident per SM.  The two resident CTAs play the role of the two ping-ponged tiles of attn_fwd_kernel — while one
// is in its softmax the other's MMAs run — and, unlike there, the prologue (barrier init, TMEM allocation, first
// Q/K/V loads: ~3 us) and the epilogue (last P V, normalise, store, teardown, next block launch: ~3 us) of one CTA
// overlap the main loop of its neighbour.  At 1024 keys (8 key tiles; 60 of SDXL's 70 attention blocks) those 6.3 us
// were a third of a CTA's life (r01: 444 TFLOP/s at 1024 keys vs 665 at 4096).  Work items are half as large, so
// the last partial wave is also half as long.  K/V tiles are staged per CTA (2-stage ring): twice the L2 -> smem
// operand traffic of the shared ring, 64 B/clk/SM, still inside the budget.
// TMEM: S [0,128)  O [128,192)  P [192,256) (bf16 pairs; A operand of O += P V read from TMEM).
constexpr int ATT1_KST = 3;                   // K ring: a stage is released as soon as S = Q K^T has read it
constexpr int ATT1_VST = 2;                   // V ring: released after O += P V
constexpr int ATT1_THREADS = 64 + 256;        // TMA warp, MMA warp, 2 column halves x 4 softmax warps
constexpr int ATT1_SMEM = ATT_TILE_BYTES + (ATT1_KST + ATT1_VST) * ATT_TILE_BYTES + 256 + 2048 + 1024;

// POLY 0: every exponential on the MUFU; m > 0: every m-th pair of scores on the FMA pipe (cubic); -1: NO exponential
// (FD_ATTN_POLY=-1: timing diagnostic with a wrong result — the floor of everything that is not an exponential).
template <int POLY>
__global__ void __launch_bounds__(ATT1_THREADS, 2)
attn_fwd1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnKParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + ATT_TILE_BYTES;
    uint8_t* sV = sK + ATT1_KST * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT1_VST * ATT_TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* k_full = bars + 1;                   // [ATT1_KST]
    uint64_t* k_empty = k_full + ATT1_KST;         // [ATT1_KST]
    uint64_t* v_full = k_empty + ATT1_KST;         // [ATT1_VST]
    uint64_t* v_empty = v_full + ATT1_VST;         // [ATT1_VST]
    uint64_t* s_full = v_empty + ATT1_VST;         // S(j) written by the tensor core
    uint64_t* s_free = s_full + 1;                 // S(j) copied to registers: S(j+1) may be issued
    uint64_t* p_ready = s_free + 1;                // P(j) in TMEM
    uint64_t* pv_done = p_ready + 1;               // O += P(j) V(j) complete
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_done + 1);
    float* mx_buf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [parity][half][128]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_blk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int n_kv_tiles = (p.Nkv + ATT_BN - 1) / ATT_BN;

    // PROLOGUE.  The thread that initialises the barriers also issues the first TMA loads (Q, the first three key tiles
    // and the first two value tiles) BEFORE the TMEM allocation and the block-wide synchronisation: a CTA has only 8
    // key tiles at 1024 keys, so the ~1 us of TMA latency in front of its first S = Q K^T is worth hiding
    // (p.early = 0 keeps the r02a order).
    //
    // K / V PIPELINE.  Keys and values travel in SEPARATE rings: K(j) is dead once S(j) = Q K(j)^T has been computed
    // (one softmax EARLIER than V(j)), so its stage goes back to the TMA thread right then and K(j+3) is requested
    // more than two tile times before it is needed.  With the shared 2-stage ring of r02a the load of tile j+2 could
    // only start after O += P(j) V(j) — about when S(j+2) was already due — and the ~1 us TMA latency sat on every
    // tile: switching the MUFU work or the row-max exchange OFF changed the kernel time by 2 % each
    // (profiles/r02_attention.txt), the loop was waiting for keys.
    const int nk_pre = p.early ? min(ATT1_KST, n_kv_tiles) : 0;
    const int nv_pre = p.early ? min(ATT1_VST, n_kv_tiles) : 0;
    if (warp == (p.early ? 0 : 1) && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT1_KST; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
        }
        for (int s = 0; s < ATT1_VST; ++s) {
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 256);
        mbar_init(p_ready, 256);
        mbar_init(pv_done, 1);
        fence_barrier_init();
        fence_proxy_async();            // the initialised barriers are about to be used by the async proxy (TMA)
    }
    auto load_k = [&](int j) {
        const int st = j % ATT1_KST;
        mbar_arrive_expect_tx(&k_full[st], ATT_TILE_BYTES);
        tma_load_3d(&tmK, &k_full[st], sK + st * ATT_TILE_BYTES, head * ATT_D, j * ATT_BN, batch);
    };
    auto load_v = [&](int j) {
        const int st = j % ATT1_VST;
        mbar_arrive_expect_tx(&v_full[st], ATT_TILE_BYTES);
        tma_load_3d(&tmV, &v_full[st], sV + st * ATT_TILE_BYTES, head * ATT_D, j * ATT_BN, batch);
    };
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        if (p.early) {
            mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_3d(&tmQ, q_full, sQ, head * ATT_D, q_blk * 128, batch);
            load_k(0);
            if (nv_pre > 0) load_v(0);
            for (int j = 1; j < nk_pre; ++j) load_k(j);
            for (int j = 1; j < nv_pre; ++j) load_v(j);
        }
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(tmem_holder, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    // waits on the S / P ping-pong between the MMA thread and the softmax warps: polled (p.spin) or suspended try_wait
    auto wait_pp = [&](uint64_t* bar, uint32_t parity) {
        if (p.spin)
            mbar_wait_poll(bar, parity);
        else
            mbar_wait(bar, parity);
    };

    if (warp == 0) {
        if (lane == 0) {
            if (!p.early) {
                mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
                tma_load_3d(&tmQ, q_full, sQ, head * ATT_D, q_blk * 128, batch);
            }
            // keys run ahead of values: K(jk) is requested as soon as S(jk - 3) is done, V(jv) as soon as P V(jv - 2) is
            int jk = nk_pre, jv = nv_pre;
            while (jk < n_kv_tiles || jv < n_kv_tiles) {
                if (jk < n_kv_tiles) {
                    mbar_wait(&k_empty[jk % ATT1_KST], ((jk / ATT1_KST) & 1) ^ 1u);
                    load_k(jk);
                    ++jk;
                }
                if (jv < n_kv_tiles && (jv + 1 < jk || jk >= n_kv_tiles)) {
                    mbar_wait(&v_empty[jv % ATT1_VST], ((jv / ATT1_VST) & 1) ^ 1u);
                    load_v(jv);
                    ++jv;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_BN, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, 0, 1);  // B (=V) MN-major
            const uint32_t q_addr = smem_u32(sQ);
            auto issue_s = [&](int j) {
                const int st = j % ATT1_KST;
                const uint32_t k_addr = smem_u32(sK + st * ATT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k)
                    tc_mma_bf16(tmem_base, make_desc_k_sw128(q_addr + k * 32), make_desc_k_sw128(k_addr + k * 32),
                                idesc_qk, k != 0 ? 1u : 0u);
                tc_commit(s_full);
                tc_commit(&k_empty[st]);            // the key tile is dead once S(j) is complete
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < n_kv_tiles; ++j) {
                if (j + 1 < n_kv_tiles) {
                    // S(j+1) as soon as S(j) sits in the softmax warps' registers and K(j+1) has landed
                    wait_pp(s_free, j & 1);
                    mbar_wait(&k_full[(j + 1) % ATT1_KST], ((j + 1) / ATT1_KST) & 1);
                    tc_fence_after();
                    issue_s(j + 1);
                }
                wait_pp(p_ready, j & 1);
                mbar_wait(&v_full[j % ATT1_VST], (j / ATT1_VST) & 1);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sV + (j % ATT1_VST) * ATT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < ATT_BN / 16; ++k)
                    tc_mma_bf16_ts(tmem_base + 128, tmem_base + 192 + k * 8,
                                   make_desc_mn_sw128(v_addr + k * 2048, 0, 1024), idesc_pv, (j | k) != 0 ? 1u : 0u);
                tc_commit(pv_done);
                tc_commit(&v_empty[j % ATT1_VST]);
            }
        }
    } else {
        // 8 softmax warps: column half h (0/1) x TMEM lane quarter (= warp % 4); a thread owns 64 of its row's 128
        // scores, the halves exchange their row maximum through shared memory once per key tile.
        const int h = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const uint32_t tmem_S = tmem_base + h * 64;
        const uint32_t tmem_O = tmem_base + 128 + h * 32;
        float m_used = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_kv_tiles; ++j) {
            wait_pp(s_full, j & 1);
            tc_fence_after();
            const int kv_valid = min(ATT_BN, p.Nkv - j * ATT_BN) - h * 64;
            uint32_t sr[2][32];
            tmem_ld_32x32(tmem_S + lane_base, sr[0]);
            tmem_ld_32x32(tmem_S + lane_base + 32, sr[1]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);
            if (kv_valid < 64) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i >= kv_valid) sr[c][i] = 0xff800000u;
            }
            float mxs[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int i = 0; i < 32; i += 2)
                    mxs[c] = max3(mxs[c], __uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1]));
            float mx = fmaxf(mxs[0], mxs[1]);
            float* slot = mx_buf + (j & 1) * 256;
            if (!(p.diag & 1)) {        // diag bit 0 (TIMING DIAGNOSTIC ONLY, wrong results): no half-row exchange
            slot[h * 128 + row] = mx;
            // only the two warps of one lane quarter exchange: a 64-thread barrier per quarter, so a quarter never
            // waits for the slowest of the eight warps
            if (p.bar_all)
                asm volatile("bar.sync 1, 256;" ::: "memory");
            else
                asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
            mx = fmaxf(mx, slot[(h ^ 1) * 128 + row]);
            }
            const float m_new = mx * p.scale_log2;
            const bool grow = m_new > m_used + ATT_RESCALE_THRESHOLD;
            const float alpha = (grow && j > 0) ? fast_exp2(m_used - m_new) : 1.0f;
            if (grow) m_used = m_new;
            l_run *= alpha;
            const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
            const float2 nm2 = make_float2(-m_used, -m_used);
            float2 ps2 = make_float2(0.f, 0.f);
            uint32_t pk[2][16];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float2 x = ffma2(make_float2(__uint_as_float(sr[c][i]), __uint_as_float(sr[c][i + 1])), sc2, nm2);
                    float2 e;
                    if (POLY > 0 && ((i >> 1) % (POLY > 0 ? POLY : 1)) == (POLY > 1 ? 1 : 0))
                        e = exp2_poly2(x);
                    else if (POLY < 0)      // TIMING DIAGNOSTIC ONLY: no MUFU, one FMA per element
                        e = ffma2(x, make_float2(0.001f, 0.001f), make_float2(1.f, 1.f));
                    else
                        e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                    ps2 = fadd2(ps2, e);
                    pk[c][i >> 1] = pack_bf16x2(e.x, e.y);
                }
            }
            l_run += ps2.x + ps2.y;
            if (j > 0) {
                wait_pp(pv_done, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_O + lane_base, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                    tmem_st_32x32(tmem_O + lane_base, r);
                    tmem_st_wait();
                }
            }
            {
                uint32_t pflat[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    pflat[i] = pk[0][i];
                    pflat[16 + i] = pk[1][i];
                }
                tmem_st_32x32(tmem_base + 192 + h * 32 + lane_base, pflat);
                tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive(p_ready);
        }
        float* slot = mx_buf + (n_kv_tiles & 1) * 256;
        slot[h * 128 + row] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
        l_run += slot[(h ^ 1) * 128 + row];
        mbar_wait(pv_done, (n_kv_tiles - 1) & 1);
        tc_fence_after();
        const int q_row = q_blk * 128 + row;
        const float inv_l = 1.f / l_run;
        bf16* orow = p.o + (long long)batch * p.o_batch_stride + (long long)q_row * p.ldo + head * ATT_D + h * 32;
        {
            uint32_t r[32];
            tmem_ld_32x32(tmem_O + lane_base, r);
            tmem_ld_wait();
            if (q_row < p.Nq) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(r[8 * q4 + 0]) * inv_l, __uint_as_float(r[8 * q4 + 1]) * inv_l);
                    u.y = pack_bf16x2(__uint_as_float(r[8 * q4 + 2]) * inv_l, __uint_as_float(r[8 * q4 + 3]) * inv_l);
                    u.z = pack_bf16x2(__uint_as_float(r[8 * q4 + 4]) * inv_l, __uint_as_float(r[8 * q4 + 5]) * inv_l);
                    u.w = pack_bf16x2(__uint_as_float(r[8 * q4 + 6]) * inv_l, __uint_as_float(r[8 * q4 + 7]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + q4 * 8) = u;
                }
            }
        }
        if (h == 0 && p.lse != nullptr && q_row < p.Nq)
            p.lse[((long long)batch * p.H + head) * p.Nq + q_row] = (m_used + log2f(l_run)) * 0.69314718055994531f;
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
    const uint32_t box[3] = {(uint32_t)ATT_D, 128u, 1u};   // 128-row tiles for Q (two per CTA), K and V
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
    static const int bar_all = getenv("FD_ATTN_BAR256") != nullptr;
    p.bar_all = bar_all;
    static const int early = getenv("FD_ATTN_EARLY") ? atoi(getenv("FD_ATTN_EARLY")) : 1;
    p.early = early;
    static const int spin = getenv("FD_ATTN_SPIN") ? atoi(getenv("FD_ATTN_SPIN")) : 0;
    p.spin = spin;
    static const int diag = getenv("FD_ATTN_DIAG") ? atoi(getenv("FD_ATTN_DIAG")) : 0;
    p.diag = diag;
    static bool attr_set = false;
    if (!attr_set) {
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd1_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT1_SMEM));
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd1_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT1_SMEM));
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd1_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT1_SMEM));
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd1_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT1_SMEM));
        FD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd1_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT1_SMEM));
        attr_set = true;
    }
    ProfScope prof(stream, PROF_ATTN_FWD, 4.0 * (double)a->B * a->H * (double)a->Nq * (double)a->Nkv * ATT_D);
    // FD_ATTN_V=2: the two-tiles-per-CTA kernel (1 CTA / SM); default: one tile per CTA, two CTAs per SM
    static const int variant = getenv("FD_ATTN_V") ? atoi(getenv("FD_ATTN_V")) : 1;
    if (variant == 1) {
        dim3 grid((a->Nq + 127) / 128, a->H, a->B);
        // share of the exponentials on the FMA pipe in the inference forward (FD_ATTN_POLY = 4: 25 %, 3: 33 %, 2: 50 %).
        // Measured on B200 (profiles/r02_attention.txt): with two CTAs per SM the all-MUFU loop is the fastest (531 /
        // 710 TFLOP/s at 1024 / 4096 keys vs 507 / 690 at 25 %, and the 50 % variant falls off the register cliff),
        // so the default is 0 and the training and inference forwards are the same kernel.
        static const int poly = getenv("FD_ATTN_POLY") ? atoi(getenv("FD_ATTN_POLY")) : 0;
        if (poly < 0)
            attn_fwd1_kernel<-1><<<grid, ATT1_THREADS, ATT1_SMEM, stream>>>(tq, tk, tv, p);
        else if (a->lse != nullptr || poly == 0)
            attn_fwd1_kernel<0><<<grid, ATT1_THREADS, ATT1_SMEM, stream>>>(tq, tk, tv, p);
        else if (poly == 2)
            attn_fwd1_kernel<2><<<grid, ATT1_THREADS, ATT1_SMEM, stream>>>(tq, tk, tv, p);
        else if (poly == 3)
            attn_fwd1_kernel<3><<<grid, ATT1_THREADS, ATT1_SMEM, stream>>>(tq, tk, tv, p);
        else
            attn_fwd1_kernel<4><<<grid, ATT1_THREADS, ATT1_SMEM, stream>>>(tq, tk, tv, p);
        FD_CHECK_LAUNCH();
        return 0;
    }
    dim3 grid((a->Nq + ATT_BM - 1) / ATT_BM, a->H, a->B);
    if (a->lse != nullptr)
        attn_fwd_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tq, tk, tv, p);
    else
        attn_fwd_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tq, tk, tv, p);
    FD_CHECK_LAUNCH();
    return 0;
}
