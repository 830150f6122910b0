#include "common.cuh"
#include "nn_kernels.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace mimamo {

template <bool BF16>
__device__ __forceinline__ uint16_t to16(float v) {
  if (BF16) { __nv_bfloat16 h = __float2bfloat16(v); return *reinterpret_cast<uint16_t*>(&h); }
  __half h = __float2half(v);
  return *reinterpret_cast<uint16_t*>(&h);
}
template <bool BF16>
__device__ __forceinline__ float from16(uint16_t u) {
  if (BF16) return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
  return __half2float(*reinterpret_cast<__half*>(&u));
}

// ---- NCHW fp32 -> NHWC 16-bit (8 channels = one 16-byte store per thread) -------------------
template <bool BF16>
__global__ void nchw_to_nhwc16_kernel(const float* __restrict__ src, long long N, int C, int H, int W,
                                      uint16_t* __restrict__ dst, int ldc, int c_off, int groups) {
  const long long total = N * H * (long long)groups * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long long r = i / W;
    const int g = (int)(r % groups); r /= groups;
    const int h = (int)(r % H);
    const long long n = r / H;
    uint16_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      v[j] = c < C ? to16<BF16>(__ldg(src + ((n * C + c) * H + h) * (long long)W + w)) : (uint16_t)0;
    }
    uint4 o;
    o.x = v[0] | ((uint32_t)v[1] << 16); o.y = v[2] | ((uint32_t)v[3] << 16);
    o.z = v[4] | ((uint32_t)v[5] << 16); o.w = v[6] | ((uint32_t)v[7] << 16);
    *reinterpret_cast<uint4*>(dst + ((n * H + h) * (long long)W + w) * ldc + c_off + g * 8) = o;
  }
}

// Row-tiled variant for small channel counts (the 24-channel phase maps): one CTA transposes kRowsPerCta image rows
// through shared memory, so the fp32 planes are read as whole rows and every pixel's channel group goes out as one
// contiguous run (the generic kernel above writes 16 bytes per lane at a `ldc`-pixel stride: 607 us per 2048 windows
// against ~250 us of traffic).
constexpr int kRowsPerCta = 4;
template <bool BF16>
__global__ void __launch_bounds__(256)
nchw_to_nhwc16_rows_kernel(const float* __restrict__ src, int C, int H, int W, uint16_t* __restrict__ dst, int ldc, int c_off,
                           int groups) {
  extern __shared__ float tile[];                              // [kRowsPerCta * W][C + 1]
  const int n = blockIdx.y;
  const int h0 = blockIdx.x * kRowsPerCta;
  const int rows = min(kRowsPerCta, H - h0);
  const int pitch = C + 1;
  for (int i = threadIdx.x; i < C * rows * W; i += blockDim.x) {
    const int c = i / (rows * W), rem = i - c * (rows * W);    // rem = r * W + w: contiguous in memory for a fixed channel
    tile[rem * pitch + c] = __ldg(src + (((size_t)n * C + c) * H + h0) * W + rem);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows * W * groups; i += blockDim.x) {
    const int px = i / groups, g = i - px * groups;
    uint16_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      v[j] = c < C ? to16<BF16>(tile[px * pitch + c]) : (uint16_t)0;
    }
    uint4 o;
    o.x = v[0] | ((uint32_t)v[1] << 16); o.y = v[2] | ((uint32_t)v[3] << 16);
    o.z = v[4] | ((uint32_t)v[5] << 16); o.w = v[6] | ((uint32_t)v[7] << 16);
    *reinterpret_cast<uint4*>(dst + (((size_t)n * H + h0) * W + px) * ldc + c_off + g * 8) = o;
  }
}

int nchw_to_nhwc16(const float* src, int N, int C, int H, int W, void* dst, int ldc, int c_off, int c_fill,
                   ElemType elem, cudaStream_t s) {
  MM_REQUIRE(c_off % 8 == 0 && c_fill % 8 == 0 && ldc % 8 == 0 && c_fill >= C, MIMAMO_E_VALUE, "nchw_to_nhwc16: channel geometry must be 8-aligned");
  if (N == 0) return MIMAMO_OK;
  const int groups = c_fill / 8;
  const size_t smem = (size_t)kRowsPerCta * W * (C + 1) * sizeof(float);
  if (smem <= 48 * 1024 && N < 65536) {
    dim3 grid((H + kRowsPerCta - 1) / kRowsPerCta, N);
    if (elem == kBF16) nchw_to_nhwc16_rows_kernel<true><<<grid, 256, smem, s>>>(src, C, H, W, (uint16_t*)dst, ldc, c_off, groups);
    else nchw_to_nhwc16_rows_kernel<false><<<grid, 256, smem, s>>>(src, C, H, W, (uint16_t*)dst, ldc, c_off, groups);
    MM_LAUNCH_OK();
    return MIMAMO_OK;
  }
  const long long total = (long long)N * H * groups * W;
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  if (elem == kBF16) nchw_to_nhwc16_kernel<true><<<grid, 256, 0, s>>>(src, N, C, H, W, (uint16_t*)dst, ldc, c_off, groups);
  else nchw_to_nhwc16_kernel<false><<<grid, 256, 0, s>>>(src, N, C, H, W, (uint16_t*)dst, ldc, c_off, groups);
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}

// ---- conv1 im2col -------------------------------------------------------------------------
template <bool BF16>
__global__ void im2col_conv1_kernel(const float* __restrict__ x, long long B, uint16_t* __restrict__ a) {
  // one thread: 8 consecutive K entries (16 B) of one output pixel; K = c*49 + kh*7 + kw, 147 -> 192
  const long long total = B * 112 * 112 * 24;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % 24);
    const long long pix = i / 24;
    const int wo = (int)(pix % 112);
    const int ho = (int)((pix / 112) % 112);
    const long long b = pix / (112 * 112);
    uint16_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float val = 0.f;
      if (k < 147) {
        const int c = k / 49, r = k - c * 49, kh = r / 7, kw = r - kh * 7;
        const int hi = ho * 2 + kh - 3, wi = wo * 2 + kw - 3;
        if (hi >= 0 && hi < 224 && wi >= 0 && wi < 224) val = __ldg(x + ((b * 3 + c) * 224 + hi) * 224ll + wi);
      }
      v[j] = to16<BF16>(val);
    }
    uint4 o;
    o.x = v[0] | ((uint32_t)v[1] << 16); o.y = v[2] | ((uint32_t)v[3] << 16);
    o.z = v[4] | ((uint32_t)v[5] << 16); o.w = v[6] | ((uint32_t)v[7] << 16);
    *reinterpret_cast<uint4*>(a + pix * 192 + g * 8) = o;
  }
}

int im2col_conv1(const float* x, int B, void* a, ElemType elem, cudaStream_t s) {
  if (B == 0) return MIMAMO_OK;
  const int grid = 148 * 16;
  if (elem == kBF16) im2col_conv1_kernel<true><<<grid, 256, 0, s>>>(x, B, (uint16_t*)a);
  else im2col_conv1_kernel<false><<<grid, 256, 0, s>>>(x, B, (uint16_t*)a);
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}

// ---- conv1 space-to-depth ----------------------------------------------------------------------
template <bool BF16>
__global__ void conv1_s2d_kernel(const float* __restrict__ x, long long B, uint4* __restrict__ out) {
  const long long total = B * 115 * 115;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % 115);
    const int Y = (int)((i / 115) % 115);
    const long long b = i / (115 * 115);
    uint16_t v[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int hi = 2 * Y + (q >> 1) - 4, wi = 2 * X + (q & 1) - 4;
      const bool in = hi >= 0 && hi < 224 && wi >= 0 && wi < 224;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        v[q * 3 + c] = to16<BF16>(in ? __ldg(x + ((b * 3 + c) * 224 + hi) * 224ll + wi) : 0.f);
    }
    v[12] = v[13] = v[14] = v[15] = 0;
    uint4 o0, o1;
    o0.x = v[0] | ((uint32_t)v[1] << 16); o0.y = v[2] | ((uint32_t)v[3] << 16);
    o0.z = v[4] | ((uint32_t)v[5] << 16); o0.w = v[6] | ((uint32_t)v[7] << 16);
    o1.x = v[8] | ((uint32_t)v[9] << 16); o1.y = v[10] | ((uint32_t)v[11] << 16);
    o1.z = 0; o1.w = 0;
    const long long row = b * 115 + Y;                           // chunk-planar: [row][chunk][X]
    out[(row * 2) * 115 + X] = o0;
    out[(row * 2 + 1) * 115 + X] = o1;
  }
}

int conv1_space_to_depth(const float* x, int B, void* s2d, ElemType elem, cudaStream_t s) {
  if (B == 0) return MIMAMO_OK;
  const int grid = 148 * 16;
  if (elem == kBF16) conv1_s2d_kernel<true><<<grid, 256, 0, s>>>(x, B, (uint4*)s2d);
  else conv1_s2d_kernel<false><<<grid, 256, 0, s>>>(x, B, (uint4*)s2d);
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}

// ---- max pool 3x3 s2 ceil_mode --------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
  if (BF16) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

template <bool BF16>
__global__ void maxpool_kernel(const uint4* __restrict__ x, long long B, int H, int W, int C8, int Ho, int Wo,
                               uint4* __restrict__ out) {
  const long long total = B * Ho * Wo * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const long long b = r / Ho;
    uint4 m;
    bool first = true;
    for (int dh = 0; dh < 3; ++dh) {
      const int h = ho * 2 + dh;
      if (h >= H) break;
      for (int dw = 0; dw < 3; ++dw) {
        const int w = wo * 2 + dw;
        if (w >= W) break;
        const uint4 v = __ldg(x + ((b * H + h) * W + w) * C8 + c);
        if (first) { m = v; first = false; }
        else { m.x = max2<BF16>(m.x, v.x); m.y = max2<BF16>(m.y, v.y); m.z = max2<BF16>(m.z, v.z); m.w = max2<BF16>(m.w, v.w); }
      }
    }
    out[i] = m;
  }
}

int maxpool3x3s2_ceil(const void* x, int B, int H, int W, int C, void* out, ElemType elem, cudaStream_t s) {
  MM_REQUIRE(C % 8 == 0, MIMAMO_E_VALUE, "maxpool: C must be a multiple of 8");
  if (B == 0) return MIMAMO_OK;
  const int Ho = (H - 3 + 1) / 2 + 1, Wo = (W - 3 + 1) / 2 + 1;      // ceil((H-3)/2) + 1
  const int grid = 148 * 16;
  if (elem == kBF16) maxpool_kernel<true><<<grid, 256, 0, s>>>((const uint4*)x, B, H, W, C / 8, Ho, Wo, (uint4*)out);
  else maxpool_kernel<false><<<grid, 256, 0, s>>>((const uint4*)x, B, H, W, C / 8, Ho, Wo, (uint4*)out);
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}

// ---- global average pool ---------------------------------------------------------------------
template <bool BF16>
__global__ void avgpool_kernel(const uint16_t* __restrict__ x, long long N, int HW, int C, float* __restrict__ out,
                               int ldo, int relu) {
  const long long total = N * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long n = i / C;
    float acc = 0.f;
    for (int p = 0; p < HW; ++p) acc += from16<BF16>(__ldg(x + (n * HW + p) * C + c));
    acc /= (float)HW;
    if (relu) acc = fmaxf(acc, 0.f);
    out[n * ldo + c] = acc;
  }
}

int avgpool_to_f32(const void* x, int N, int HW, int C, float* out, int ldo, int relu, ElemType elem, cudaStream_t s) {
  if (N == 0) return MIMAMO_OK;
  const long long total = (long long)N * C;
  const int grid = (int)((total + 255) / 256);
  if (elem == kBF16) avgpool_kernel<true><<<grid, 256, 0, s>>>((const uint16_t*)x, N, HW, C, out, ldo, relu);
  else avgpool_kernel<false><<<grid, 256, 0, s>>>((const uint16_t*)x, N, HW, C, out, ldo, relu);
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}

// ---- per-channel mean of an NHWC 16-bit tensor (weight-rounding calibration, conv_layer_quantize) ----
// Deterministic: block b sums a fixed slice of rows per channel, a second kernel adds the slices in order.
template <bool BF16>
__global__ void channel_sum_kernel(const uint16_t* __restrict__ x, long long M, int C, int ld, int rows_per_block, double* __restrict__ partial) {
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double acc = 0.0;
    for (long long r = r0; r < r1; ++r) acc += (double)from16<BF16>(__ldg(x + r * ld + c));
    partial[(size_t)blockIdx.x * C + c] = acc;
  }
}
__global__ void channel_mean_finish_kernel(const double* __restrict__ partial, int blocks, int C, long long M, float* __restrict__ mean) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double acc = 0.0;
  for (int b = 0; b < blocks; ++b) acc += partial[(size_t)b * C + c];
  mean[c] = (float)(acc / (double)M);
}

int channel_means(const void* x, long long M, int C, int ld, float* mean_dev, double* scratch, size_t scratch_doubles, ElemType elem, cudaStream_t s) {
  if (M == 0 || C == 0) return MIMAMO_OK;
  int blocks = (int)(scratch_doubles / (size_t)C);
  if (blocks > 128) blocks = 128;
  MM_REQUIRE(blocks >= 1, MIMAMO_E_VALUE, "channel_means: scratch too small");
  const int rows_per_block = (int)((M + blocks - 1) / blocks);
  blocks = (int)((M + rows_per_block - 1) / rows_per_block);
  if (elem == kBF16) channel_sum_kernel<true><<<blocks, 256, 0, s>>>((const uint16_t*)x, M, C, ld, rows_per_block, scratch);
  else channel_sum_kernel<false><<<blocks, 256, 0, s>>>((const uint16_t*)x, M, C, ld, rows_per_block, scratch);
  MM_LAUNCH_OK();
  channel_mean_finish_kernel<<<(C + 255) / 256, 256, 0, s>>>(scratch, blocks, C, M, mean_dev);
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}

// ---- fp32 linear layer (head MLP / FC / GRU input projections; ~1.4 MMAC per window) ---------
constexpr int kLinTile = 64, kLinK = 16;

__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ a, int lda, int M, const float* __restrict__ w, int K, int N,
              const float* __restrict__ bias, const float* __restrict__ pre_s, const float* __restrict__ pre_t,
              const float* __restrict__ post_s, const float* __restrict__ post_t, int relu, float* __restrict__ out, int ldo) {
  __shared__ float As[kLinK][kLinTile + 4];
  __shared__ float Ws[kLinK][kLinTile + 4];
  const int m0 = blockIdx.y * kLinTile, n0 = blockIdx.x * kLinTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kLinK) {
    for (int i = threadIdx.x; i < kLinTile * kLinK; i += 256) {
      const int r = i / kLinK, kk = i - r * kLinK;
      const int m = m0 + r, n = n0 + r, k = k0 + kk;
      As[kk][r] = (m < M && k < K) ? __ldg(a + (size_t)m * lda + k) : 0.f;
      Ws[kk][r] = (n < N && k < K) ? __ldg(w + (size_t)n * K + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kLinK; ++kk) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; wv[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (pre_s) v = v * pre_s[n] + pre_t[n];
      if (relu) v = fmaxf(v, 0.f);
      if (post_s) v = v * post_s[n] + post_t[n];
      out[(size_t)m * ldo + n] = v;
    }
  }
}

// Same contract, K % 32 == 0 and 16-byte aligned rows: float4 global loads prefetched into registers one K block
// ahead, transposed into shared memory, float4 shared loads in the 4x4 register-blocked inner product.
constexpr int kLinK4 = 32;
__global__ void __launch_bounds__(256)
linear_kernel_v4(const float* __restrict__ a, int lda, int M, const float* __restrict__ w, int K, int N,
                 const float* __restrict__ bias, const float* __restrict__ pre_s, const float* __restrict__ pre_t,
                 const float* __restrict__ post_s, const float* __restrict__ post_t, int relu, float* __restrict__ out, int ldo) {
  __shared__ __align__(16) float As[kLinK4][kLinTile + 4];
  __shared__ __align__(16) float Ws[kLinK4][kLinTile + 4];
  const int m0 = blockIdx.y * kLinTile, n0 = blockIdx.x * kLinTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lr = threadIdx.x >> 3, lk = (threadIdx.x & 7) * 4;      // loader: rows lr, lr + 32; k offset lk
  float4 pa[2], pw[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int m = m0 + lr + 32 * u, n = n0 + lr + 32 * u;
      pa[u] = m < M ? __ldg(reinterpret_cast<const float4*>(a + (size_t)m * lda + k0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
      pw[u] = n < N ? __ldg(reinterpret_cast<const float4*>(w + (size_t)n * K + k0 + lk)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float acc[4][4] = {};
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kLinK4) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = lr + 32 * u;
      As[lk][r] = pa[u].x; As[lk + 1][r] = pa[u].y; As[lk + 2][r] = pa[u].z; As[lk + 3][r] = pa[u].w;
      Ws[lk][r] = pw[u].x; Ws[lk + 1][r] = pw[u].y; Ws[lk + 2][r] = pw[u].z; Ws[lk + 3][r] = pw[u].w;
    }
    __syncthreads();
    if (k0 + kLinK4 < K) fetch(k0 + kLinK4);
#pragma unroll
    for (int kk = 0; kk < kLinK4; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (pre_s) v = v * pre_s[n] + pre_t[n];
      if (relu) v = fmaxf(v, 0.f);
      if (post_s) v = v * post_s[n] + post_t[n];
      out[(size_t)m * ldo + n] = v;
    }
  }
}

static int up_opt(float** dst, const float* src, size_t n) {
  *dst = nullptr;
  if (!src) return MIMAMO_OK;
  return upload(dst, src, n);
}

int linear_init(LinearLayer& L, const float* w, const float* bias, int out_f, int in_f, int relu, const float* pre_s,
                const float* pre_t, const float* post_s, const float* post_t) {
  L.in_f = in_f; L.out_f = out_f; L.relu = relu;
  int rc = upload(&L.w, w, (size_t)out_f * in_f);
  if (!rc) rc = up_opt(&L.bias, bias, out_f);
  if (!rc) rc = up_opt(&L.pre_s, pre_s, out_f);
  if (!rc) rc = up_opt(&L.pre_t, pre_t, out_f);
  if (!rc) rc = up_opt(&L.post_s, post_s, out_f);
  if (!rc) rc = up_opt(&L.post_t, post_t, out_f);
  return rc;
}

void linear_free(LinearLayer& L) {
  cudaFree(L.w); cudaFree(L.bias); cudaFree(L.pre_s); cudaFree(L.pre_t); cudaFree(L.post_s); cudaFree(L.post_t);
  L = LinearLayer();
}

int linear_forward(const LinearLayer& L, const float* a, int lda, int M, float* out, int ldo, cudaStream_t s) {
  if (M == 0) return MIMAMO_OK;
  dim3 grid((L.out_f + kLinTile - 1) / kLinTile, (M + kLinTile - 1) / kLinTile);
  const bool vec = L.in_f % kLinK4 == 0 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0;
  if (vec) linear_kernel_v4<<<grid, 256, 0, s>>>(a, lda, M, L.w, L.in_f, L.out_f, L.bias, L.pre_s, L.pre_t, L.post_s, L.post_t,
                                                 L.relu, out, ldo);
  else linear_kernel<<<grid, 256, 0, s>>>(a, lda, M, L.w, L.in_f, L.out_f, L.bias, L.pre_s, L.pre_t, L.post_s, L.post_t,
                                          L.relu, out, ldo);
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}

// ---- GRU recurrence ---------------------------------------------------------------------------
// The reference builds nn.GRU without batch_first (api/mimamo_net.py:119) and feeds (bs, frames, 256),
// so the recurrence runs over dim 0 (snippets) and the 64 frames are independent batch rows
// (SURVEY.md section 0.2).  One CTA owns kGruRows batch rows of one direction for the whole
// sequence (persistent over time); thread j owns hidden unit j.  Gate order r, z, n as in torch.
constexpr int kGruRows = 1;       // one batch row per CTA: 64 rows x 2 directions = 128 CTAs fill the GPU and the per-step matvec is 4x shorter
                                  // than with 4 rows per CTA (the recurrence is a chain of 2 x S dependent steps: latency is what counts)
constexpr int kGruHd = 128, kGruG = 3 * kGruHd;
constexpr int kGruSmem = (kGruHd * kGruG + 2 * kGruRows * kGruHd + kGruRows * kGruG) * (int)sizeof(float);

// W_hh^T of the CTA's direction (128 x 384 fp32 = 192 KB) is loaded into shared memory once and stays
// there for the whole sequence; thread g owns gate column g (coalesced, conflict-free reads of W),
// the hidden state is broadcast from shared memory four k at a time.
__global__ void __launch_bounds__(kGruG)
gru_kernel(const float* __restrict__ xproj, const float* __restrict__ whhT, const float* __restrict__ bhh, int S, int Bt,
           float* __restrict__ y) {
  constexpr int Hd = kGruHd, G = kGruG, R = kGruRows;
  extern __shared__ __align__(16) float gsm[];
  float* W = gsm;                          // [Hd][G]
  float* h = W + Hd * G;                   // [2][R][Hd]
  float* gh = h + 2 * R * Hd;              // [R][G]  W_hh h + b_hh
  const int dir = blockIdx.y, g = threadIdx.x;
  const int row0 = blockIdx.x * R;
  {
    const float4* src = reinterpret_cast<const float4*>(whhT + (size_t)dir * Hd * G);
    float4* dst = reinterpret_cast<float4*>(W);
    for (int i = threadIdx.x; i < Hd * G / 4; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  const float bias = bhh[dir * G + g];
  for (int i = threadIdx.x; i < R * Hd; i += blockDim.x) h[i] = 0.f;
  __syncthreads();
  int cur = 0;
  for (int step = 0; step < S; ++step) {
    const int s = dir == 0 ? step : S - 1 - step;
    // this step's input projections are fetched before the matvec so that their latency hides behind it
    float xr[2], xz[2], xn[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = threadIdx.x + u * (int)blockDim.x;
      xr[u] = xz[u] = xn[u] = 0.f;
      if (i < R * Hd) {
        const int r = i / Hd, j = i - r * Hd;
        if (row0 + r < Bt) {
          const float* xp = xproj + (((size_t)s * Bt + row0 + r) * 2 + dir) * G;
          xr[u] = __ldg(xp + j); xz[u] = __ldg(xp + Hd + j); xn[u] = __ldg(xp + 2 * Hd + j);
        }
      }
    }
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = bias;
    const float* hc = h + cur * R * Hd;
#pragma unroll 2
    for (int k = 0; k < Hd; k += 4) {
      const float w0 = W[(k + 0) * G + g], w1 = W[(k + 1) * G + g], w2 = W[(k + 2) * G + g], w3 = W[(k + 3) * G + g];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(hc + r * Hd + k);
        acc[r] = fmaf(hv.x, w0, acc[r]); acc[r] = fmaf(hv.y, w1, acc[r]);
        acc[r] = fmaf(hv.z, w2, acc[r]); acc[r] = fmaf(hv.w, w3, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) gh[r * G + g] = acc[r];
    __syncthreads();
    float* hn_buf = h + (cur ^ 1) * R * Hd;
    static_assert(R * Hd <= 2 * G, "two gate items per thread at most");
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = threadIdx.x + u * (int)blockDim.x;
      if (i >= R * Hd) break;
      const int r = i / Hd, j = i - r * Hd;
      const int row = row0 + r;
      float hn = 0.f;
      if (row < Bt) {
        const float rg = 1.f / (1.f + expf(-(xr[u] + gh[r * G + j])));
        const float zg = 1.f / (1.f + expf(-(xz[u] + gh[r * G + Hd + j])));
        const float ng = tanhf(xn[u] + rg * gh[r * G + 2 * Hd + j]);
        hn = (1.f - zg) * ng + zg * hc[r * Hd + j];
        y[((size_t)s * Bt + row) * (2 * Hd) + dir * Hd + j] = hn;
      }
      hn_buf[i] = hn;
    }
    __syncthreads();
    cur ^= 1;
  }
}

int gru_layer(const float* xproj, const float* whhT, const float* bhh, int S, int Bt, int Hd, float* y, cudaStream_t s) {
  MM_REQUIRE(Hd == kGruHd, MIMAMO_E_RUNTIME, "GRU kernel is specialised for hidden size 128");
  if (S == 0 || Bt == 0) return MIMAMO_OK;
  static DeviceOnce attr;
  if (attr.need()) {
    MM_CUDA(cudaFuncSetAttribute(gru_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGruSmem));
    attr.mark();
  }
  dim3 grid((Bt + kGruRows - 1) / kGruRows, 2);
  gru_kernel<<<grid, kGruG, kGruSmem, s>>>(xproj, whhT, bhh, S, Bt, y);
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}

}  // namespace mimamo
