// fd_elem.cu — layout conversion and elementwise kernels (HBM-bound), plus the fused fp32
// elementwise kernels of the distillation step (noising, CFG + DPM-Solver++ update, student output).
// Reference call sites: src/flash/models/flash/flash_diffusion_model.py:243-257,267-280,316-328.
#include "fd_common.cuh"
#include "fd_host.h"

namespace fd {

__device__ __forceinline__ void ld8(const bf16* p, float (&f)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 t;
    t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ void st8(bf16* p, const float (&f)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                              pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

#define FD_GRID_STRIDE(i, n) \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

static inline int grid_for(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    const long long cap = (long long)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// x [NB,C,H,W] fp32 -> y [NB,H,W,Cpad] bf16 (small C: conv_in)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ y, int NB, int C,
                                    int HW, int Cpad) {
    const long long total = (long long)NB * HW * Cpad;
    FD_GRID_STRIDE(i, total) {
        const int c = (int)(i % Cpad);
        const long long p = i / Cpad;
        const int n = (int)(p / HW);
        const int hw = (int)(p % HW);
        y[i] = c < C ? __float2bfloat16(x[((long long)n * C + c) * HW + hw]) : __float2bfloat16(0.f);
    }
}

template <bool FP32IN>
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ x, long long ld, float* __restrict__ y,
                                    int NB, int C, int HW) {
    const long long total = (long long)NB * C * HW;
    FD_GRID_STRIDE(i, total) {
        const int hw = (int)(i % HW);
        const long long t = i / HW;
        const int c = (int)(t % C);
        const int n = (int)(t / C);
        const long long src = ((long long)n * HW + hw) * ld + c;
        y[i] = FP32IN ? reinterpret_cast<const float*>(x)[src]
                      : __bfloat162float(reinterpret_cast<const bf16*>(x)[src]);
    }
}

// x [NB*h*w, p*p*Cout] fp32 (columns ordered (pi, qi, c)) -> y [NB, Ckeep, h*p, w*p] fp32
__global__ void unpatchify_kernel(const float* __restrict__ x, float* __restrict__ y, int NB, int h, int w, int p,
                                  int Cout, int Ckeep) {
    const int H = h * p, W = w * p;
    const long long total = (long long)NB * Ckeep * H * W;
    FD_GRID_STRIDE(i, total) {
        const int X = (int)(i % W);
        long long t = i / W;
        const int Y = (int)(t % H);
        t /= H;
        const int c = (int)(t % Ckeep);
        const int n = (int)(t / Ckeep);
        const int hy = Y / p, pi = Y % p, wx = X / p, qi = X % p;
        y[i] = x[(((long long)n * h + hy) * w + wx) * (p * p * Cout) + (pi * p + qi) * Cout + c];
    }
}

// nearest 2x upsample, vectors of 8 channels
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int NB, int H,
                                  int W, int CV) {
    const long long total = (long long)NB * (2 * H) * (2 * W) * CV;
    FD_GRID_STRIDE(i, total) {
        const int cv = (int)(i % CV);
        long long p = i / CV;
        const int wo = (int)(p % (2 * W));
        p /= (2 * W);
        const int ho = (int)(p % (2 * H));
        const int n = (int)(p / (2 * H));
        y[i] = x[(((long long)n * H + (ho >> 1)) * W + (wo >> 1)) * CV + cv];
    }
}

__global__ void upsample2x_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int NB, int H,
                                      int W, int C) {
    const int CV = C / 8;
    const long long total = (long long)NB * H * W * CV;
    FD_GRID_STRIDE(i, total) {
        const int cv = (int)(i % CV);
        long long p = i / CV;
        const int w = (int)(p % W);
        p /= W;
        const int h = (int)(p % H);
        const int n = (int)(p / H);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float v[8];
                ld8(dy + ((((long long)n * 2 * H + 2 * h + a) * 2 * W + 2 * w + b) * C + cv * 8), v);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += v[j];
            }
        st8(dx + i * 8, acc);
    }
}

// x [NB,H,W,C] -> y [4*NB, H/2, W/2, C], phase-major: y[(p*NB+n), h/2, w/2] = x[n,h,w], p=(h&1)*2+(w&1)
template <bool INVERSE>
__global__ void space_to_depth_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int NB, int H,
                                      int W, int CV) {
    const long long total = (long long)NB * H * W * CV;
    FD_GRID_STRIDE(i, total) {
        const int cv = (int)(i % CV);
        long long p = i / CV;
        const int w = (int)(p % W);
        p /= W;
        const int h = (int)(p % H);
        const int n = (int)(p / H);
        const int ph = (h & 1) * 2 + (w & 1);
        const long long j = ((((long long)ph * NB + n) * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1)) * CV + cv;
        if (INVERSE)
            y[i] = x[j];
        else
            y[j] = x[i];
    }
}

__global__ void concat_kernel(const uint4* __restrict__ a, int V1, const uint4* __restrict__ b, int V2,
                              uint4* __restrict__ y, long long rows) {
    const int V = V1 + V2;
    const long long total = rows * V;
    FD_GRID_STRIDE(i, total) {
        const int v = (int)(i % V);
        const long long r = i / V;
        y[i] = v < V1 ? a[r * V1 + v] : b[r * V2 + (v - V1)];
    }
}

__global__ void slice_kernel(const uint4* __restrict__ x, int Vtot, int v0, int V, uint4* __restrict__ y,
                             long long rows) {
    const long long total = rows * V;
    FD_GRID_STRIDE(i, total) {
        const int v = (int)(i % V);
        const long long r = i / V;
        y[i] = x[r * Vtot + v0 + v];
    }
}

__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y,
                           long long nvec) {
    FD_GRID_STRIDE(i, nvec) {
        float x[8], z[8];
        ld8(a + i * 8, x);
        ld8(b + i * 8, z);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += z[j];
        st8(y + i * 8, x);
    }
}

// bf16 [rows, cols] -> [cols, rows], 32x32 tiles through shared memory
__global__ void transpose_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int rows, int cols,
                                 long long ldy) {
    __shared__ bf16 tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = x[(long long)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) y[(long long)c * ldy + r] = tile[threadIdx.x][j];
    }
}

__global__ void cast_scale_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n, float s) {
    FD_GRID_STRIDE(i, n) y[i] = __float2bfloat16(x[i] * s);
}

__global__ void silu_cast_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
    FD_GRID_STRIDE(i, n) y[i] = __float2bfloat16(silu(x[i]));
}

// Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos(t f_i) | sin(t f_i)],
// f_i = exp(-ln(10000) * i / half)
__global__ void timestep_embedding_kernel(const float* __restrict__ t, bf16* __restrict__ y, int NB,
                                          int dim) {
    const int half = dim / 2;
    const long long total = (long long)NB * half;
    FD_GRID_STRIDE(i, total) {
        const int k = (int)(i % half);
        const int n = (int)(i / half);
        const float f = expf(-9.210340371976184f * (float)k / (float)half);
        const float a = t[n] * f;
        y[(long long)n * dim + k] = __float2bfloat16(cosf(a));
        y[(long long)n * dim + half + k] = __float2bfloat16(sinf(a));
    }
}

// acc interleaved: per 32-col block [16 value | 16 gate]; dout [M, N/2]
__global__ void geglu_bwd_kernel(const bf16* __restrict__ acc, const bf16* __restrict__ dout,
                                 bf16* __restrict__ dacc, long long M, int N) {
    const int blocks = N / 32;
    const long long total = M * blocks * 2;  // 8-wide vectors of the value half
    FD_GRID_STRIDE(i, total) {
        const int hv = (int)(i & 1);
        const long long t = i >> 1;
        const int b = (int)(t % blocks);
        const long long r = t / blocks;
        const bf16* pa = acc + r * N + b * 32 + hv * 8;
        float v[8], g[8], d[8];
        ld8(pa, v);
        ld8(pa + 16, g);
        ld8(dout + r * (N / 2) + b * 16 + hv * 8, d);
        float dv[8], dg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            dv[j] = d[j] * gelu_erf(g[j]);
            dg[j] = d[j] * v[j] * gelu_erf_grad(g[j]);
        }
        bf16* pd = dacc + r * N + b * 32 + hv * 8;
        st8(pd, dv);
        st8(pd + 16, dg);
    }
}

// dacc = dout * d/dx gelu_tanh(acc)   (the activation the GEMM epilogue applies with act = 1)
__global__ void gelu_tanh_bwd_kernel(const bf16* __restrict__ acc, const bf16* __restrict__ dout,
                                     bf16* __restrict__ dacc, long long nvec) {
    FD_GRID_STRIDE(i, nvec) {
        float a[8], d[8], o[8];
        ld8(acc + i * 8, a);
        ld8(dout + i * 8, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x = a[j];
            const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
            const float t = tanhf(u);
            const float du = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x * x);
            o[j] = d[j] * (0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du);
        }
        st8(dacc + i * 8, o);
    }
}

// inverse of unpatchify_kernel for the gradient: dy [NB, Ckeep, h*p, w*p] fp32 -> dx [NB*h*w, p*p*Cout] bf16
// (columns ordered (pi, qi, c); channels >= Ckeep get zero)
__global__ void patchify_kernel(const float* __restrict__ dy, bf16* __restrict__ dx, int NB, int h, int w, int p,
                                int Cout, int Ckeep) {
    const int H = h * p, W = w * p, cols = p * p * Cout;
    const long long total = (long long)NB * h * w * cols;
    FD_GRID_STRIDE(i, total) {
        const int col = (int)(i % cols);
        long long t = i / cols;
        const int wx = (int)(t % w);
        t /= w;
        const int hy = (int)(t % h);
        const int n = (int)(t / h);
        const int c = col % Cout, pq = col / Cout;
        const int pi = pq / p, qi = pq % p;
        float v = 0.f;
        if (c < Ckeep) v = dy[(((long long)n * Ckeep + c) * H + hy * p + pi) * W + wx * p + qi];
        dx[i] = __float2bfloat16(v);
    }
}

// ------------------------------------------------------------------ distillation-step kernels
__global__ void add_noise_kernel(const float* __restrict__ z, const float* __restrict__ noise,
                                 const float* __restrict__ sa, const float* __restrict__ sg,
                                 float* __restrict__ out, int B, long long n) {
    const long long total = (long long)B * n;
    FD_GRID_STRIDE(i, total) {
        const int b = (int)(i / n);
        out[i] = sa[b] * z[i] + sg[b] * noise[i];
    }
}

__global__ void cfg_dpm_kernel(const float* __restrict__ eps_c, const float* __restrict__ eps_u,
                               float* __restrict__ x, float* __restrict__ x0_prev, float w,
                               float alpha_t, float sigma_t, float c_x, float c_d0, float c_d1r,
                               long long n) {
    FD_GRID_STRIDE(i, n) {
        const float eps = w * eps_c[i] + (1.f - w) * eps_u[i];
        const float xi = x[i];
        const float x0 = (xi - sigma_t * eps) / alpha_t;
        const float d1 = x0 - x0_prev[i];
        x[i] = c_x * xi - c_d0 * x0 - c_d1r * d1;
        x0_prev[i] = x0;
    }
}

__global__ void student_output_kernel(const float* __restrict__ x_t, const float* __restrict__ eps,
                                      const float* __restrict__ sa, const float* __restrict__ sg,
                                      const float* __restrict__ c_skip, const float* __restrict__ c_out,
                                      float* __restrict__ out, int B, long long n) {
    const long long total = (long long)B * n;
    FD_GRID_STRIDE(i, total) {
        const int b = (int)(i / n);
        const float x0 = (x_t[i] - sg[b] * eps[i]) / sa[b];
        out[i] = c_skip[b] * x_t[i] + c_out[b] * x0;
    }
}


// ---------------------------------------------------------------------------- row softmax (large-head attention)
// The VAE mid-block attention has ONE head of 512 channels (UPSTREAM AutoencoderKL: Attention(heads=1, dim_head=512)):
// its O accumulator alone would fill the 512 TMEM columns, so it runs as two tcgen05 GEMMs (S = Q K^T, O = P V) around
// this row softmax instead of the fused FlashAttention kernels.  One block per row, fp32 math, 16-byte accesses; the
// row is read three times (max, sum, write) out of L1/L2.
__device__ __forceinline__ float block_reduce(float v, float* sm, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sm[w] = v;
    __syncthreads();
    float r = is_max ? -INFINITY : 0.f;
    for (int i = 0; i < nw; ++i) r = is_max ? fmaxf(r, sm[i]) : r + sm[i];
    return r;
}

__global__ void softmax_rows_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                    int L, float scale_log2) {
    __shared__ float sm[32];
    const bf16* xr = x + (long long)blockIdx.x * ldx;
    bf16* yr = y + (long long)blockIdx.x * ldy;
    const int nvec = L >> 3;
    float mx = -INFINITY;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        float f[8];
        ld8(xr + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) mx = fmaxf(mx, f[j]);
    }
    mx = block_reduce(mx, sm, true) * scale_log2;
    float sum = 0.f;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        float f[8];
        ld8(xr + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += exp2f(fmaf(f[j], scale_log2, -mx));
    }
    sum = block_reduce(sum, sm, false);
    const float inv = 1.f / sum;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        float f[8];
        ld8(xr + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = exp2f(fmaf(f[j], scale_log2, -mx)) * inv;
        st8(yr + v * 8, f);
    }
}

// ds = p * (dp - sum_j p_j dp_j) * scale   (backward of y = softmax(scale * x) given p = y)
__global__ void softmax_rows_bwd_kernel(const bf16* __restrict__ p, const bf16* __restrict__ dp, bf16* __restrict__ ds,
                                        long long ld, int L, float scale) {
    __shared__ float sm[32];
    const bf16* pr = p + (long long)blockIdx.x * ld;
    const bf16* dr = dp + (long long)blockIdx.x * ld;
    bf16* sr = ds + (long long)blockIdx.x * ld;
    const int nvec = L >> 3;
    float dot = 0.f;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        float a[8], b[8];
        ld8(pr + v * 8, a);
        ld8(dr + v * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) dot += a[j] * b[j];
    }
    dot = block_reduce(dot, sm, false);
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        float a[8], b[8];
        ld8(pr + v * 8, a);
        ld8(dr + v * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = a[j] * (b[j] - dot) * scale;
        st8(sr + v * 8, a);
    }
}


// ---------------------------------------------------------------------------- LPIPS pieces (VGG16 feature stack)
// 2x2 / stride-2 max pooling on NHWC bf16 (torchvision VGG16 `MaxPool2d(2, 2)`), 8 channels per thread.
__global__ void maxpool2x2_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int NB, int H, int W, int C) {
    const int Ho = H >> 1, Wo = W >> 1, nv = C >> 3;
    const long long total = (long long)NB * Ho * Wo * nv;
    FD_GRID_STRIDE(i, total) {
        const int v = (int)(i % nv);
        long long p = i / nv;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int n = (int)(p / Ho);
        const bf16* base = x + (((long long)n * H + 2 * ho) * W + 2 * wo) * C + v * 8;
        float a[8], b[8], c[8], d[8];
        ld8(base, a); ld8(base + C, b); ld8(base + (long long)W * C, c); ld8(base + (long long)W * C + C, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaxf(fmaxf(a[j], b[j]), fmaxf(c[j], d[j]));
        st8(y + (((long long)n * Ho + ho) * Wo + wo) * C + v * 8, a);
    }
}
// gradient of the above: dy goes to the FIRST maximum of each window in row-major order (torch's tie rule)
__global__ void maxpool2x2_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                                      int NB, int H, int W, int C) {
    const int Ho = H >> 1, Wo = W >> 1, nv = C >> 3;
    const long long total = (long long)NB * Ho * Wo * nv;
    FD_GRID_STRIDE(i, total) {
        const int v = (int)(i % nv);
        long long p = i / nv;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int n = (int)(p / Ho);
        const long long o00 = (((long long)n * H + 2 * ho) * W + 2 * wo) * C + v * 8;
        const long long offs[4] = {o00, o00 + C, o00 + (long long)W * C, o00 + (long long)W * C + C};
        float f[4][8], g[8], out[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k) ld8(x + offs[k], f[k]);
        ld8(dy + (((long long)n * Ho + ho) * Wo + wo) * C + v * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int best = 0;
            float m = f[0][j];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (f[k][j] > m) { m = f[k][j]; best = k; }
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k][j] = (k == best) ? g[j] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) st8(dx + offs[k], out[k]);
    }
}
// dx = dy where the ReLU output y is positive
__global__ void relu_bwd_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                                long long nvec) {
    FD_GRID_STRIDE(i, nvec) {
        float a[8], g[8];
        ld8(y + i * 8, a);
        ld8(dy + i * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = a[j] > 0.f ? g[j] : 0.f;
        st8(dx + i * 8, g);
    }
}

// One LPIPS layer: out[n] += (1 / HW) sum_pixels sum_c w_c (f0_c / (|f0| + eps) - f1_c / (|f1| + eps))^2.
// One warp per pixel (C <= 512: 16 channels per lane), block-level partial sums, one atomic per block and image.
constexpr int LP_MAXV = 2;   // 8-channel vectors per lane: C <= 512
__global__ void lpips_layer_kernel(const bf16* __restrict__ f0, const bf16* __restrict__ f1, const float* __restrict__ w,
                                   float* __restrict__ out, int HW, int C, float inv_hw) {
    __shared__ float sm[32];
    const int n = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nv = C >> 3;
    float acc = 0.f;
    for (int pix = blockIdx.x * nw + warp; pix < HW; pix += gridDim.x * nw) {
        const long long off = ((long long)n * HW + pix) * C;
        float a[LP_MAXV][8], b[LP_MAXV][8];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < LP_MAXV; ++k) {
            const int v = lane + k * 32;
            if (v < nv) {
                ld8(f0 + off + v * 8, a[k]);
                ld8(f1 + off + v * 8, b[k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) { s0 += a[k][j] * a[k][j]; s1 += b[k][j] * b[k][j]; }
            }
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        const float i0 = 1.f / (sqrtf(s0) + 1e-10f), i1 = 1.f / (sqrtf(s1) + 1e-10f);
#pragma unroll
        for (int k = 0; k < LP_MAXV; ++k) {
            const int v = lane + k * 32;
            if (v < nv) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = a[k][j] * i0 - b[k][j] * i1;
                    acc += __ldg(w + v * 8 + j) * d * d;
                }
            }
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < nw; ++i) t += sm[i];
        atomicAdd(out + n, t * inv_hw);
    }
}
// gradient with respect to f0: with u = f0 / s, s = |f0| + eps, a_c = 2 w_c (u_c - g_c) gout[n] / HW:
//   df0 = (a - u (a . u) s / |f0|) / s
__global__ void lpips_layer_bwd_kernel(const bf16* __restrict__ f0, const bf16* __restrict__ f1,
                                       const float* __restrict__ w, const float* __restrict__ gout,
                                       bf16* __restrict__ df0, int HW, int C, float inv_hw) {
    const int n = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nv = C >> 3;
    const float go = gout[n] * inv_hw;
    for (int pix = blockIdx.x * nw + warp; pix < HW; pix += gridDim.x * nw) {
        const long long off = ((long long)n * HW + pix) * C;
        float a[LP_MAXV][8], b[LP_MAXV][8];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < LP_MAXV; ++k) {
            const int v = lane + k * 32;
            if (v < nv) {
                ld8(f0 + off + v * 8, a[k]);
                ld8(f1 + off + v * 8, b[k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) { s0 += a[k][j] * a[k][j]; s1 += b[k][j] * b[k][j]; }
            }
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        const float n0 = sqrtf(s0);
        const float i0 = 1.f / (n0 + 1e-10f), i1 = 1.f / (sqrtf(s1) + 1e-10f);
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < LP_MAXV; ++k) {
            const int v = lane + k * 32;
            if (v < nv) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float u = a[k][j] * i0;
                    const float g = 2.f * __ldg(w + v * 8 + j) * (u - b[k][j] * i1) * go;
                    b[k][j] = g;              // a_c
                    a[k][j] = u;              // u_c
                    dot += g * u;
                }
            }
        }
        dot = warp_sum(dot);
        const float corr = n0 > 0.f ? dot / (n0 * i0) : 0.f;       // (a . u) s / |f0|
#pragma unroll
        for (int k = 0; k < LP_MAXV; ++k) {
            const int v = lane + k * 32;
            if (v < nv) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = (b[k][j] - a[k][j] * corr) * i0;
                st8(df0 + off + v * 8, o);
            }
        }
    }
}

}  // namespace fd

using namespace fd;

extern "C" int fd_nchw_to_nhwc(const float* x, void* y, int32_t NB, int32_t C, int32_t H, int32_t W,
                               int32_t Cpad, void* stream) {
    FD_CHECK_ARG(Cpad >= C, "fd_nchw_to_nhwc: Cpad < C");
    const long long total = (long long)NB * H * W * Cpad;
    nchw_to_nhwc_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, NB, C, H * W, Cpad);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_nhwc_to_nchw(const void* x, int32_t x_is_fp32, int64_t ld, float* y, int32_t NB,
                               int32_t C, int32_t H, int32_t W, void* stream) {
    const long long total = (long long)NB * C * H * W;
    if (x_is_fp32)
        nhwc_to_nchw_kernel<true><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, ld, y, NB, C, H * W);
    else
        nhwc_to_nchw_kernel<false><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, ld, y, NB, C, H * W);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_unpatchify(const float* x, float* y, int32_t NB, int32_t h, int32_t w, int32_t p, int32_t Cout,
                             int32_t Ckeep, void* stream) {
    FD_CHECK_ARG(Ckeep <= Cout && p > 0, "fd_unpatchify: bad channels");
    const long long total = (long long)NB * Ckeep * h * p * w * p;
    unpatchify_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, y, NB, h, w, p, Cout, Ckeep);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_upsample2x(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C,
                             void* stream) {
    FD_CHECK_ARG(C % 8 == 0, "fd_upsample2x: C %% 8");
    const long long total = (long long)NB * 4 * H * W * (C / 8);
    upsample2x_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, NB, H, W, C / 8);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_upsample2x_bwd(const void* dy, void* dx, int32_t NB, int32_t H, int32_t W, int32_t C,
                                 void* stream) {
    FD_CHECK_ARG(C % 8 == 0, "fd_upsample2x_bwd: C %% 8");
    const long long total = (long long)NB * H * W * (C / 8);
    upsample2x_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (bf16*)dx, NB, H, W, C);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_space_to_depth(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C,
                                 void* stream) {
    FD_CHECK_ARG(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "fd_space_to_depth: bad shape");
    const long long total = (long long)NB * H * W * (C / 8);
    space_to_depth_kernel<false><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, NB, H, W, C / 8);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_depth_to_space(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C,
                                 void* stream) {
    FD_CHECK_ARG(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "fd_depth_to_space: bad shape");
    const long long total = (long long)NB * H * W * (C / 8);
    space_to_depth_kernel<true><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, NB, H, W, C / 8);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_concat_channels(const void* a, int32_t C1, const void* b, int32_t C2, void* y,
                                  int64_t rows, void* stream) {
    FD_CHECK_ARG(C1 % 8 == 0 && C2 % 8 == 0, "fd_concat_channels: C %% 8");
    const long long total = rows * ((C1 + C2) / 8);
    concat_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const uint4*)a, C1 / 8, (const uint4*)b, C2 / 8, (uint4*)y, rows);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_slice_channels(const void* x, int32_t Ctot, int32_t c0, int32_t C, void* y, int64_t rows,
                                 void* stream) {
    FD_CHECK_ARG(Ctot % 8 == 0 && c0 % 8 == 0 && C % 8 == 0, "fd_slice_channels: C %% 8");
    slice_kernel<<<grid_for(rows * (C / 8)), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, Ctot / 8, c0 / 8, C / 8, (uint4*)y, rows);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_add(const void* a, const void* b, void* y, int64_t n, void* stream) {
    FD_CHECK_ARG(n % 8 == 0, "fd_add: n %% 8");
    add_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, (bf16*)y, n / 8);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_transpose(const void* x, void* y, int32_t rows, int32_t cols, int64_t ldy, void* stream) {
    FD_CHECK_ARG(ldy >= rows, "fd_transpose: ldy < rows");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, rows, cols, ldy);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_cast_scale(const float* x, void* y, int64_t n, float scale, void* stream) {
    cast_scale_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, n, scale);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_silu_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
    silu_cast_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, n);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_timestep_embedding(const float* t, void* y, int32_t NB, int32_t dim, void* stream) {
    FD_CHECK_ARG(dim % 2 == 0, "fd_timestep_embedding: odd dim");
    timestep_embedding_kernel<<<grid_for((long long)NB * dim / 2), 256, 0, (cudaStream_t)stream>>>(t, (bf16*)y, NB, dim);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_geglu_bwd(const void* acc, const void* dout, void* dacc, int64_t M, int32_t N,
                            void* stream) {
    FD_CHECK_ARG(N % 32 == 0, "fd_geglu_bwd: N %% 32");
    const long long total = M * (N / 32) * 2;
    geglu_bwd_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const bf16*)acc, (const bf16*)dout, (bf16*)dacc, M, N);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_step_add_noise(const float* z, const float* noise, const float* sa, const float* sg,
                                 float* out, int32_t B, int64_t n, void* stream) {
    add_noise_kernel<<<grid_for((long long)B * n), 256, 0, (cudaStream_t)stream>>>(z, noise, sa, sg, out, B, n);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_step_cfg_dpm(const float* eps_c, const float* eps_u, float* x, float* x0_prev,
                               const float* coef6, int64_t n, void* stream) {
    cfg_dpm_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(eps_c, eps_u, x, x0_prev, coef6[0], coef6[1],
                                                                  coef6[2], coef6[3], coef6[4], coef6[5], n);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_step_student_output(const float* x_t, const float* eps, const float* sa, const float* sg,
                                      const float* c_skip, const float* c_out, float* out, int32_t B,
                                      int64_t n, void* stream) {
    student_output_kernel<<<grid_for((long long)B * n), 256, 0, (cudaStream_t)stream>>>(x_t, eps, sa, sg, c_skip, c_out, out, B, n);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_gelu_tanh_bwd(const void* acc, const void* dout, void* dacc, int64_t n, void* stream) {
    FD_CHECK_ARG(n % 8 == 0, "fd_gelu_tanh_bwd: n %% 8");
    gelu_tanh_bwd_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)acc, (const bf16*)dout,
                                                                            (bf16*)dacc, n / 8);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_patchify(const float* dy, void* dx, int32_t NB, int32_t h, int32_t w, int32_t p, int32_t Cout,
                           int32_t Ckeep, void* stream) {
    FD_CHECK_ARG(Ckeep <= Cout && p > 0, "fd_patchify: bad channels");
    const long long total = (long long)NB * h * w * p * p * Cout;
    patchify_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(dy, (bf16*)dx, NB, h, w, p, Cout, Ckeep);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_softmax_rows(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t rows, int32_t L, float scale,
                               void* stream) {
    FD_CHECK_ARG(rows > 0 && L > 0 && L % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "fd_softmax_rows: L / strides must be multiples of 8");
    fd::softmax_rows_kernel<<<rows, L >= 2048 ? 256 : 128, 0, (cudaStream_t)stream>>>(
        (const fd::bf16*)x, ldx, (fd::bf16*)y, ldy, L, scale * 1.4426950408889634f);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_softmax_rows_bwd(const void* p, const void* dp, void* ds, int64_t ld, int32_t rows, int32_t L,
                                   float scale, void* stream) {
    FD_CHECK_ARG(rows > 0 && L > 0 && L % 8 == 0 && ld % 8 == 0, "fd_softmax_rows_bwd: L / stride must be multiples of 8");
    fd::softmax_rows_bwd_kernel<<<rows, L >= 2048 ? 256 : 128, 0, (cudaStream_t)stream>>>(
        (const fd::bf16*)p, (const fd::bf16*)dp, (fd::bf16*)ds, ld, L, scale);
    FD_CHECK_LAUNCH();
    return 0;
}

extern "C" int fd_maxpool2x2(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C, void* stream) {
    FD_CHECK_ARG(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "fd_maxpool2x2: C %% 8, even H / W");
    const long long total = (long long)NB * (H / 2) * (W / 2) * (C / 8);
    fd::maxpool2x2_kernel<<<fd::grid_for(total), 256, 0, (cudaStream_t)stream>>>((const fd::bf16*)x, (fd::bf16*)y, NB, H, W, C);
    FD_CHECK_LAUNCH();
    return 0;
}
extern "C" int fd_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int32_t NB, int32_t H, int32_t W, int32_t C,
                                 void* stream) {
    FD_CHECK_ARG(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "fd_maxpool2x2_bwd: C %% 8, even H / W");
    const long long total = (long long)NB * (H / 2) * (W / 2) * (C / 8);
    fd::maxpool2x2_bwd_kernel<<<fd::grid_for(total), 256, 0, (cudaStream_t)stream>>>(
        (const fd::bf16*)x, (const fd::bf16*)dy, (fd::bf16*)dx, NB, H, W, C);
    FD_CHECK_LAUNCH();
    return 0;
}
extern "C" int fd_relu_bwd(const void* y, const void* dy, void* dx, int64_t n, void* stream) {
    FD_CHECK_ARG(n % 8 == 0, "fd_relu_bwd: n %% 8");
    fd::relu_bwd_kernel<<<fd::grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>((const fd::bf16*)y, (const fd::bf16*)dy,
                                                                               (fd::bf16*)dx, n / 8);
    FD_CHECK_LAUNCH();
    return 0;
}
extern "C" int fd_lpips_layer(const void* f0, const void* f1, const float* w, float* out, int32_t NB, int32_t HW,
                              int32_t C, void* stream) {
    FD_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * fd::LP_MAXV, "fd_lpips_layer: C=%d must be a multiple of 8, <= 512", C);
    int gx = (HW + 7) / 8;
    const int cap = (fd::num_sms() * 8 + NB - 1) / NB;
    if (gx > cap) gx = cap;
    fd::lpips_layer_kernel<<<dim3(gx, NB), 256, 0, (cudaStream_t)stream>>>((const fd::bf16*)f0, (const fd::bf16*)f1, w, out,
                                                                           HW, C, 1.0f / (float)HW);
    FD_CHECK_LAUNCH();
    return 0;
}
extern "C" int fd_lpips_layer_bwd(const void* f0, const void* f1, const float* w, const float* gout, void* df0,
                                  int32_t NB, int32_t HW, int32_t C, void* stream) {
    FD_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * fd::LP_MAXV, "fd_lpips_layer_bwd: C=%d must be a multiple of 8, <= 512", C);
    int gx = (HW + 7) / 8;
    const int cap = (fd::num_sms() * 8 + NB - 1) / NB;
    if (gx > cap) gx = cap;
    fd::lpips_layer_bwd_kernel<<<dim3(gx, NB), 256, 0, (cudaStream_t)stream>>>(
        (const fd::bf16*)f0, (const fd::bf16*)f1, w, gout, (fd::bf16*)df0, HW, C, 1.0f / (float)HW);
    FD_CHECK_LAUNCH();
    return 0;
}
