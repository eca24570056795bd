// fd_norm.cu — HBM-bound normalisation kernels (GroupNorm / LayerNorm, forward + backward) on
// channels-last bf16 activations.  Vectorised 16-byte accesses, fp32 statistics, warp-shuffle
// reductions.  UPSTREAM math: torch.nn.GroupNorm / LayerNorm as used by diffusers ResnetBlock2D,
// Transformer2DModel and BasicTransformerBlock.
#include "fd_common.cuh"
#include "fd_host.h"

namespace fd {

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 t;
    t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                              pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ---------------------------------------------------------------------------- GroupNorm
// raw[n, g, 0..1] += (sum, sum of squares) over this block's rows.  blockDim = (nvec, rpi).
// Optionally a second input (dy) turns this into the backward reduction:
//   MODE 0: (sum x, sum x^2)
//   MODE 1: (sum dyg, sum dyg * xhat) with dyg = dy * act'(pre) * gamma
template <int MODE>
__global__ void gn_reduce_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                 const float* __restrict__ stats, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float* __restrict__ raw, int HW,
                                 int C, int G, int rows_per_block, int silu_act) {
    extern __shared__ float sm[];  // [2*G]
    const int n = blockIdx.y;
    const int cpg = C / G;
    const int vec = threadIdx.x;
    const int c0 = vec * 8;
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 2 * G; i += blockDim.x * blockDim.y) sm[i] = 0.f;
    __syncthreads();

    float sc[8], sh[8], mu[8], rs[8];
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = (c0 + j) / cpg;
            mu[j] = stats[(n * G + g) * 2 + 0];
            rs[j] = stats[(n * G + g) * 2 + 1];
            sc[j] = gamma[c0 + j];
            sh[j] = beta[c0 + j];
        }
    }
    const int r_begin = blockIdx.x * rows_per_block;
    const int r_end = min(HW, r_begin + rows_per_block);
    int r = r_begin + threadIdx.y;
    if (MODE == 0) {
        // four independent 16-byte loads in flight per thread: the kernel is latency-bound otherwise (r01: 2.2 TB/s)
        const int st = blockDim.y;
        for (; r + 3 * st < r_end; r += 4 * st) {
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                u[k] = *reinterpret_cast<const uint4*>(x + ((long long)n * HW + r + k * st) * C + c0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 t = unpack_bf16x2(w[j]);
                    a0[2 * j] += t.x;
                    a1[2 * j] += t.x * t.x;
                    a0[2 * j + 1] += t.y;
                    a1[2 * j + 1] += t.y * t.y;
                }
            }
        }
    }
    for (; r < r_end; r += blockDim.y) {
        const long long off = ((long long)n * HW + r) * C + c0;
        float xv[8];
        load8(x + off, xv);
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a0[j] += xv[j];
                a1[j] += xv[j] * xv[j];
            }
        } else {
            float dv[8];
            load8(dy + off, dv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xh = (xv[j] - mu[j]) * rs[j];
                float d = dv[j];
                if (silu_act) d *= silu_grad(xh * sc[j] + sh[j]);
                d *= sc[j];
                a0[j] += d;
                a1[j] += d * xh;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int g = (c0 + j) / cpg;
        atomicAdd(&sm[2 * g], a0[j]);
        atomicAdd(&sm[2 * g + 1], a1[j]);
    }
    __syncthreads();
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 2 * G; i += blockDim.x * blockDim.y)
        atomicAdd(&raw[(long long)n * G * 2 + i], sm[i]);
}

__global__ void gn_finalize_kernel(const float* __restrict__ raw, float* __restrict__ stats, int total,
                                   float inv_count, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float mean = raw[2 * i] * inv_count;
    float var = raw[2 * i + 1] * inv_count - mean * mean;
    var = fmaxf(var, 0.f);
    stats[2 * i] = mean;
    stats[2 * i + 1] = rsqrtf(var + eps);
}

// y = act((x - mean) * rstd * gamma + beta).  grid (row chunks, NB), block (nvec, rpi).
// RAW: `stats` holds the raw (sum, sum of squares) of gn_reduce_kernel<0>; mean / rstd are formed here (the separate
// finalize launch is gone) and, when stats_out != NULL, written once per image for the backward.
// RAW == 2: `stats` holds PER-CHANNEL sums [NB, C, 2] left by the epilogue of the GEMM / conv that produced x
// (FdGemmArgs.colstats_out); every block first folds them into the G group sums in shared memory (eight partial sums per
// group, fixed order: deterministic), so no reduction pass over x is needed at all.
template <int RAW>
__global__ void gn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                bf16* __restrict__ y, int HW, int C, int G, int rows_per_block,
                                int silu_act, float inv_count, float eps, float* __restrict__ stats_out) {
    const int n = blockIdx.y;
    const int cpg = C / G;
    const int c0 = threadIdx.x * 8;
    if (RAW == 2) {
        extern __shared__ float gsum[];           // [G][2] group sums, then [G * 8][2] partial sums
        float* part = gsum + 2 * G;
        const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
        const float2* cs = reinterpret_cast<const float2*>(stats) + (long long)n * C;
        for (int t = tid; t < G * 8; t += nthr) {
            const int g = t >> 3, sub = t & 7;
            float a = 0.f, b = 0.f;
            for (int c = sub; c < cpg; c += 8) {
                const float2 v = __ldg(cs + g * cpg + c);
                a += v.x;
                b += v.y;
            }
            part[2 * t] = a;
            part[2 * t + 1] = b;
        }
        __syncthreads();
        for (int g = tid; g < G; g += nthr) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a += part[2 * (g * 8 + k)];
                b += part[2 * (g * 8 + k) + 1];
            }
            gsum[2 * g] = a;
            gsum[2 * g + 1] = b;
        }
        __syncthreads();
    }
    extern __shared__ float gsum_[];
    const float* gst = (RAW == 2) ? gsum_ : stats + (long long)n * G * 2;     // this image's [G][2]
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int g = (c0 + j) / cpg;
        float mean = gst[g * 2 + 0];
        float rstd = gst[g * 2 + 1];
        if (RAW) {
            mean *= inv_count;
            rstd = rsqrtf(fmaxf(rstd * inv_count - mean * mean, 0.f) + eps);
        }
        sc[j] = rstd * gamma[c0 + j];
        sh[j] = beta[c0 + j] - mean * sc[j];
    }
    if (RAW && stats_out != nullptr && blockIdx.x == 0) {
        for (int g = threadIdx.y * blockDim.x + threadIdx.x; g < G; g += blockDim.x * blockDim.y) {
            const float mean = gst[g * 2 + 0] * inv_count;
            const float var = fmaxf(gst[g * 2 + 1] * inv_count - mean * mean, 0.f);
            stats_out[(n * G + g) * 2 + 0] = mean;
            stats_out[(n * G + g) * 2 + 1] = rsqrtf(var + eps);
        }
    }
    const int r_begin = blockIdx.x * rows_per_block;
    const int r_end = min(HW, r_begin + rows_per_block);
    int r = r_begin + threadIdx.y;
    {
        const int st = blockDim.y;
        for (; r + 3 * st < r_end; r += 4 * st) {       // four loads in flight per thread
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                u[k] = *reinterpret_cast<const uint4*>(x + ((long long)n * HW + r + k * st) * C + c0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
                float v[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 t = unpack_bf16x2(w[j]);
                    const float t0 = t.x * sc[2 * j] + sh[2 * j], t1 = t.y * sc[2 * j + 1] + sh[2 * j + 1];
                    v[2 * j] = silu_act ? silu(t0) : t0;
                    v[2 * j + 1] = silu_act ? silu(t1) : t1;
                }
                store8(y + ((long long)n * HW + r + k * st) * C + c0, v);
            }
        }
    }
    for (; r < r_end; r += blockDim.y) {
        const long long off = ((long long)n * HW + r) * C + c0;
        float v[8];
        load8(x + off, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = v[j] * sc[j] + sh[j];
            v[j] = silu_act ? silu(t) : t;
        }
        store8(y + off, v);
    }
}

// dx = rstd * (dyg - (s1 + xhat * s2) / m)
__global__ void gn_bwd_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                    const float* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ raw,
                                    bf16* __restrict__ dx, int HW, int C, int G, int rows_per_block,
                                    int silu_act, float inv_m) {
    const int n = blockIdx.y;
    const int cpg = C / G;
    const int c0 = threadIdx.x * 8;
    float sc[8], sh[8], mu[8], rs[8], s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int g = (c0 + j) / cpg;
        mu[j] = stats[(n * G + g) * 2 + 0];
        rs[j] = stats[(n * G + g) * 2 + 1];
        s1[j] = raw[(n * G + g) * 2 + 0] * inv_m;
        s2[j] = raw[(n * G + g) * 2 + 1] * inv_m;
        sc[j] = gamma[c0 + j];
        sh[j] = beta[c0 + j];
    }
    const int r_begin = blockIdx.x * rows_per_block;
    const int r_end = min(HW, r_begin + rows_per_block);
    for (int r = r_begin + threadIdx.y; r < r_end; r += blockDim.y) {
        const long long off = ((long long)n * HW + r) * C + c0;
        float xv[8], dv[8];
        load8(x + off, xv);
        load8(dy + off, dv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = (xv[j] - mu[j]) * rs[j];
            float d = dv[j];
            if (silu_act) d *= silu_grad(xh * sc[j] + sh[j]);
            d *= sc[j];
            dv[j] = rs[j] * (d - s1[j] - xh * s2[j]);
        }
        store8(dx + off, dv);
    }
}

static void gn_geometry(int HW, int C, int NB, dim3& grid, dim3& block, int& rows_per_block) {
    const int nvec = C / 8;
    int rpi = 256 / nvec;
    if (rpi < 1) rpi = 1;
    if (rpi > HW) rpi = HW;
    block = dim3(nvec, rpi);
    int target_blocks = (8 * num_sms() + NB - 1) / NB;  // per image
    int max_chunks = (HW + rpi * 4 - 1) / (rpi * 4);    // at least 4 rows per thread
    if (max_chunks < 1) max_chunks = 1;
    int chunks = target_blocks < max_chunks ? target_blocks : max_chunks;
    if (chunks < 1) chunks = 1;
    rows_per_block = (HW + chunks - 1) / chunks;
    chunks = (HW + rows_per_block - 1) / rows_per_block;
    grid = dim3(chunks, NB);
}

// ---------------------------------------------------------------------------- LayerNorm
// one warp per row; C % 8 == 0, C <= 2048
constexpr int LN_MAXV = 8;  // vectors of 8 per lane -> C <= 2048

__global__ void ln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                              const float* __restrict__ beta, bf16* __restrict__ y,
                              float* __restrict__ stats, int rows, int C, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int nvec = C >> 3;
    const bf16* xr = x + (long long)warp * C;
    float v[LN_MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            load8(xr + vi * 8, v[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[i][j];
        }
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[i][j] - mean;
                q += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    if (stats != nullptr && lane == 0) {
        stats[2 * warp] = mean;
        stats[2 * warp + 1] = rstd;
    }
    bf16* yr = y + (long long)warp * C;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd;
            if (gamma != nullptr) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
                o[0] *= g0.x; o[1] *= g0.y; o[2] *= g0.z; o[3] *= g0.w;
                o[4] *= g1.x; o[5] *= g1.y; o[6] *= g1.z; o[7] *= g1.w;
                if (beta != nullptr) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
                    o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w;
                    o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
                }
            }
            store8(yr + vi * 8, o);
        }
    }
}

// y = LN(x) * (1 + scale[b]) + shift[b]   (AdaLN modulation, no affine), one warp per row
__global__ void ln_modulate_kernel(const bf16* __restrict__ x, const float* __restrict__ scale,
                                   const float* __restrict__ shift, long long ld_mod, bf16* __restrict__ y, int rows,
                                   int C, int rows_per_batch, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int nvec = C >> 3;
    const bf16* xr = x + (long long)warp * C;
    float v[LN_MAXV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            load8(xr + vi * 8, v[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[i][j];
        }
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[i][j] - mean;
                q += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    const float* sc = scale + (long long)(warp / rows_per_batch) * ld_mod;
    const float* sh = shift + (long long)(warp / rows_per_batch) * ld_mod;
    bf16* yr = y + (long long)warp * C;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(sc + vi * 8));
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(sc + vi * 8 + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(sh + vi * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(sh + vi * 8 + 4));
            const float sca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float shf[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * (1.0f + sca[j]) + shf[j];
            store8(yr + vi * 8, o);
        }
    }
}

__global__ void ln_bwd_kernel(const bf16* __restrict__ x, const float* __restrict__ stats,
                              const float* __restrict__ gamma, const bf16* __restrict__ dy,
                              bf16* __restrict__ dx, int rows, int C) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int nvec = C >> 3;
    const bf16* xr = x + (long long)warp * C;
    const bf16* dr = dy + (long long)warp * C;
    const float mean = stats[2 * warp], rstd = stats[2 * warp + 1];
    float xh[LN_MAXV][8], dg[LN_MAXV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            load8(xr + vi * 8, xh[i]);
            load8(dr + vi * 8, dg[i]);
            if (gamma != nullptr) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
                dg[i][0] *= g0.x; dg[i][1] *= g0.y; dg[i][2] *= g0.z; dg[i][3] *= g0.w;
                dg[i][4] *= g1.x; dg[i][5] *= g1.y; dg[i][6] *= g1.z; dg[i][7] *= g1.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                xh[i][j] = (xh[i][j] - mean) * rstd;
                s1 += dg[i][j];
                s2 += dg[i][j] * xh[i][j];
            }
        }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    bf16* dxr = dx + (long long)warp * C;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = rstd * (dg[i][j] - s1 - xh[i][j] * s2);
            store8(dxr + vi * 8, o);
        }
    }
}


// Backward of y = LN(x) * (1 + scale[b]) + shift[b] (no affine): recomputes the row statistics from x,
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * (1 + scale[b]),
//   dscale[b, c] += sum_rows dy * xhat,   dshift[b, c] += sum_rows dy.
// grid (chunks, B): a block owns a contiguous run of rows of ONE sample, accumulates the two column sums in shared
// memory (2*C floats) and adds them to global memory once at the end.
__global__ void ln_modulate_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                       const float* __restrict__ scale, long long ld_mod, bf16* __restrict__ dx,
                                       float* __restrict__ dscale, float* __restrict__ dshift, int C,
                                       int rows_per_batch, int rows_per_block, float eps) {
    extern __shared__ float ln_acc[];            // [2][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int b = blockIdx.y;
    const int nvec = C >> 3;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) ln_acc[i] = 0.f;
    __syncthreads();
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(r0 + rows_per_block, rows_per_batch);
    const float* sc = scale + (long long)b * ld_mod;
    for (int r = r0 + warp; r < r1; r += nwarps) {
        const long long row = (long long)b * rows_per_batch + r;
        const bf16* xr = x + row * C;
        const bf16* dr = dy + row * C;
        float xh[LN_MAXV][8], dg[LN_MAXV][8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int vi = lane + i * 32;
            if (vi < nvec) {
                load8(xr + vi * 8, xh[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += xh[i][j];
            }
        }
        const float mean = warp_sum(s) / C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int vi = lane + i * 32;
            if (vi < nvec) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xh[i][j] -= mean;
                    q += xh[i][j] * xh[i][j];
                }
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / C + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int vi = lane + i * 32;
            if (vi < nvec) {
                load8(dr + vi * 8, dg[i]);
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(sc + vi * 8));
                const float4 a1 = __ldg(reinterpret_cast<const float4*>(sc + vi * 8 + 4));
                const float sca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xh[i][j] *= rstd;
                    atomicAdd(&ln_acc[vi * 8 + j], dg[i][j] * xh[i][j]);
                    atomicAdd(&ln_acc[C + vi * 8 + j], dg[i][j]);
                    dg[i][j] *= 1.0f + sca[j];
                    s1 += dg[i][j];
                    s2 += dg[i][j] * xh[i][j];
                }
            }
        }
        s1 = warp_sum(s1) / C;
        s2 = warp_sum(s2) / C;
        bf16* dxr = dx + row * C;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int vi = lane + i * 32;
            if (vi < nvec) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = rstd * (dg[i][j] - s1 - xh[i][j] * s2);
                store8(dxr + vi * 8, o);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        atomicAdd(dscale + (long long)b * C + i, ln_acc[i]);
        atomicAdd(dshift + (long long)b * C + i, ln_acc[C + i]);
    }
}

// out = res + gate[b] * h ; backward dh = gate[b] * dout, dgate[b, c] += sum_rows dout * h   (AdaLN-Zero gates)
__global__ void gate_residual_kernel(const bf16* __restrict__ h, const float* __restrict__ gate, long long ld_gate,
                                     const bf16* __restrict__ res, bf16* __restrict__ out, long long rows, int C,
                                     int rows_per_batch) {
    const int cv = C >> 3;
    const long long total = rows * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / cv;
        const int v = (int)(i - row * cv);
        const float* g = gate + (row / rows_per_batch) * ld_gate + v * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(g));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float a[8], r[8];
        load8(h + row * C + v * 8, a);
        load8(res + row * C + v * 8, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += gg[j] * a[j];
        store8(out + row * C + v * 8, r);
    }
}

__global__ void gate_bwd_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ h,
                                const float* __restrict__ gate, long long ld_gate, bf16* __restrict__ dh,
                                float* __restrict__ dgate, int C, int rows_per_batch, int rows_per_block) {
    extern __shared__ float ln_acc[];            // [C]
    const int b = blockIdx.y;
    const int cv = C >> 3;
    for (int i = threadIdx.x; i < C; i += blockDim.x) ln_acc[i] = 0.f;
    __syncthreads();
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(r0 + rows_per_block, rows_per_batch);
    const float* gp = gate + (long long)b * ld_gate;
    // a thread keeps the same 8 columns for all its rows when blockDim.x is a multiple of cv; otherwise it just
    // walks the (row, vector) space: the shared-memory adds make either correct
    const long long total = (long long)(r1 - r0) * cv;
    for (long long i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = (int)(i / cv);
        const int v = (int)(i - (long long)r * cv);
        const long long row = (long long)b * rows_per_batch + r0 + r;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp + v * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gp + v * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float d[8], a[8], o[8];
        load8(dout + row * C + v * 8, d);
        load8(h + row * C + v * 8, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j] = gg[j] * d[j];
            atomicAdd(&ln_acc[v * 8 + j], d[j] * a[j]);
        }
        store8(dh + row * C + v * 8, o);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dgate + (long long)b * C + i, ln_acc[i]);
}

static void per_sample_geometry(int B, int rows_per_batch, int min_rows, dim3& grid, int& rows_per_block) {
    int chunks = (num_sms() * 4 + B - 1) / B;
    const int max_chunks = (rows_per_batch + min_rows - 1) / min_rows;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    rows_per_block = (rows_per_batch + chunks - 1) / chunks;
    chunks = (rows_per_batch + rows_per_block - 1) / rows_per_block;
    grid = dim3(chunks, B);
}

}  // namespace fd

using namespace fd;

extern "C" int fd_groupnorm_stats(const void* x, float* stats, int32_t NB, int32_t HW, int32_t C,
                                  int32_t G, float eps, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && C % G == 0 && C / 8 <= 1024, "fd_groupnorm_stats: bad C=%d G=%d", C, G);
    // stats doubles as the raw accumulation buffer (sum, sumsq) before finalisation
    FD_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * NB * G, stream));
    dim3 grid, block;
    int rpb;
    gn_geometry(HW, C, NB, grid, block, rpb);
    gn_reduce_kernel<0><<<grid, block, 2 * G * sizeof(float), stream>>>(
        (const bf16*)x, nullptr, nullptr, nullptr, nullptr, stats, HW, C, G, rpb, 0);
    FD_CHECK_LAUNCH();
    const int total = NB * G;
    gn_finalize_kernel<<<(total + 127) / 128, 128, 0, stream>>>(stats, stats, total,
                                                                1.0f / ((float)HW * (C / G)), eps);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_groupnorm_apply(const void* x, const float* stats, const float* gamma,
                                  const float* beta, void* y, int32_t NB, int32_t HW, int32_t C,
                                  int32_t G, int32_t silu_act, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && C % G == 0 && C / 8 <= 1024, "fd_groupnorm_apply: bad C=%d G=%d", C, G);
    dim3 grid, block;
    int rpb;
    gn_geometry(HW, C, NB, grid, block, rpb);
    gn_apply_kernel<0><<<grid, block, 0, stream>>>((const bf16*)x, stats, gamma, beta, (bf16*)y, HW, C, G,
                                                       rpb, silu_act, 0.f, 0.f, nullptr);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* raw,
                                float* stats_out, int32_t NB, int32_t HW, int32_t C, int32_t G, float eps,
                                int32_t silu_act, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && C % G == 0 && C / 8 <= 1024, "fd_groupnorm_fwd: bad C=%d G=%d", C, G);
    FD_CHECK_ARG(raw != nullptr && raw != stats_out, "fd_groupnorm_fwd: raw scratch must be a distinct buffer");
    FD_CHECK_CUDA(cudaMemsetAsync(raw, 0, sizeof(float) * 2 * NB * G, stream));
    dim3 grid, block;
    int rpb;
    gn_geometry(HW, C, NB, grid, block, rpb);
    gn_reduce_kernel<0><<<grid, block, 2 * G * sizeof(float), stream>>>(
        (const bf16*)x, nullptr, nullptr, nullptr, nullptr, raw, HW, C, G, rpb, 0);
    FD_CHECK_LAUNCH();
    gn_apply_kernel<1><<<grid, block, 0, stream>>>((const bf16*)x, raw, gamma, beta, (bf16*)y, HW, C, G, rpb,
                                                      silu_act, 1.0f / ((float)HW * (C / G)), eps, stats_out);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_groupnorm_apply_cols(const void* x, const float* colstats, const float* gamma, const float* beta,
                                       void* y, float* stats_out, int32_t NB, int32_t HW, int32_t C, int32_t G,
                                       float eps, int32_t silu_act, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && C % G == 0 && C / 8 <= 1024 && G <= 512,
                 "fd_groupnorm_apply_cols: bad C=%d G=%d", C, G);
    dim3 grid, block;
    int rpb;
    gn_geometry(HW, C, NB, grid, block, rpb);
    gn_apply_kernel<2><<<grid, block, 18 * G * sizeof(float), stream>>>((const bf16*)x, colstats, gamma, beta, (bf16*)y,
                                                                       HW, C, G, rpb, silu_act,
                                                                       1.0f / ((float)HW * (C / G)), eps, stats_out);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_groupnorm_bwd(const void* x, const float* stats, const float* gamma,
                                const float* beta, const void* dy, void* dx, float* scratch,
                                int32_t NB, int32_t HW, int32_t C, int32_t G, int32_t silu_act,
                                void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && C % G == 0 && C / 8 <= 1024, "fd_groupnorm_bwd: bad C=%d G=%d", C, G);
    FD_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * NB * G, stream));
    dim3 grid, block;
    int rpb;
    gn_geometry(HW, C, NB, grid, block, rpb);
    gn_reduce_kernel<1><<<grid, block, 2 * G * sizeof(float), stream>>>(
        (const bf16*)x, (const bf16*)dy, stats, gamma, beta, scratch, HW, C, G, rpb, silu_act);
    FD_CHECK_LAUNCH();
    gn_bwd_apply_kernel<<<grid, block, 0, stream>>>((const bf16*)x, (const bf16*)dy, stats, gamma, beta,
                                                    scratch, (bf16*)dx, HW, C, G, rpb, silu_act,
                                                    1.0f / ((float)HW * (C / G)));
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y,
                                float* stats, int32_t rows, int32_t C, float eps, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * LN_MAXV, "fd_layernorm_fwd: bad C=%d", C);
    const int warps_per_block = 8;
    const int blocks = (rows + warps_per_block - 1) / warps_per_block;
    ln_fwd_kernel<<<blocks, warps_per_block * 32, 0, stream>>>((const bf16*)x, gamma, beta, (bf16*)y,
                                                               stats, rows, C, eps);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_layernorm_modulate(const void* x, const float* scale, const float* shift, int64_t ld_mod, void* y,
                                     int32_t rows, int32_t C, int32_t rows_per_batch, float eps, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * LN_MAXV && ld_mod % 4 == 0 && rows_per_batch > 0,
                 "fd_layernorm_modulate: bad C=%d / ld_mod", C);
    const int warps_per_block = 8;
    const int blocks = (rows + warps_per_block - 1) / warps_per_block;
    ln_modulate_kernel<<<blocks, warps_per_block * 32, 0, stream>>>((const bf16*)x, scale, shift, ld_mod, (bf16*)y,
                                                                    rows, C, rows_per_batch, eps);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_layernorm_bwd(const void* x, const float* stats, const float* gamma,
                                const void* dy, void* dx, int32_t rows, int32_t C, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * LN_MAXV, "fd_layernorm_bwd: bad C=%d", C);
    const int warps_per_block = 8;
    const int blocks = (rows + warps_per_block - 1) / warps_per_block;
    ln_bwd_kernel<<<blocks, warps_per_block * 32, 0, stream>>>((const bf16*)x, stats, gamma,
                                                               (const bf16*)dy, (bf16*)dx, rows, C);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_layernorm_modulate_bwd(const void* x, const void* dy, const float* scale, int64_t ld_mod, void* dx,
                                         float* dscale, float* dshift, int32_t rows, int32_t C,
                                         int32_t rows_per_batch, float eps, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * LN_MAXV && ld_mod % 4 == 0 && rows_per_batch > 0 &&
                     rows % rows_per_batch == 0,
                 "fd_layernorm_modulate_bwd: bad C=%d / ld_mod / rows", C);
    const int B = rows / rows_per_batch;
    FD_CHECK_CUDA(cudaMemsetAsync(dscale, 0, sizeof(float) * (size_t)B * C, stream));
    FD_CHECK_CUDA(cudaMemsetAsync(dshift, 0, sizeof(float) * (size_t)B * C, stream));
    dim3 grid;
    int rpb;
    per_sample_geometry(B, rows_per_batch, 16, grid, rpb);
    ln_modulate_bwd_kernel<<<grid, 256, 2 * C * sizeof(float), stream>>>(
        (const bf16*)x, (const bf16*)dy, scale, ld_mod, (bf16*)dx, dscale, dshift, C, rows_per_batch, rpb, eps);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_gate_residual(const void* h, const float* gate, int64_t ld_gate, const void* res, void* out,
                                int32_t rows, int32_t C, int32_t rows_per_batch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && ld_gate % 4 == 0 && rows_per_batch > 0, "fd_gate_residual: bad C=%d / ld_gate", C);
    const long long total = (long long)rows * (C / 8);
    long long g = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (g > cap) g = cap;
    gate_residual_kernel<<<(int)g, 256, 0, stream>>>((const bf16*)h, gate, ld_gate, (const bf16*)res, (bf16*)out,
                                                     rows, C, rows_per_batch);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_gate_bwd(const void* dout, const void* h, const float* gate, int64_t ld_gate, void* dh,
                           float* dgate, int32_t rows, int32_t C, int32_t rows_per_batch, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    FD_CHECK_ARG(C % 8 == 0 && ld_gate % 4 == 0 && rows_per_batch > 0 && rows % rows_per_batch == 0,
                 "fd_gate_bwd: bad C=%d / ld_gate / rows", C);
    const int B = rows / rows_per_batch;
    FD_CHECK_CUDA(cudaMemsetAsync(dgate, 0, sizeof(float) * (size_t)B * C, stream));
    dim3 grid;
    int rpb;
    per_sample_geometry(B, rows_per_batch, 16, grid, rpb);
    gate_bwd_kernel<<<grid, 256, C * sizeof(float), stream>>>((const bf16*)dout, (const bf16*)h, gate, ld_gate,
                                                              (bf16*)dh, dgate, C, rows_per_batch, rpb);
    FD_CHECK_LAUNCH();
    return 0;
}
