// On-device replacement of the reference's PIL / torchvision preprocessing of OpenFace face crops
// (SURVEY.md section 8(f).2):
//
//   gray stack : Image.open(bmp).convert('L') -> Resize(phase_size, LANCZOS) -> float / 255
//                (api/sampler/snippet_sampler.py:156-185, api/utils/data_utils.py:71-120)
//   RGB frame  : Resize(256) [PIL bilinear] -> CenterCrop(224) -> ToTensor -> x * 255 -> Normalize(mean, 1)
//                (api/utils/model_utils.py:26-40, api/sampler/image_sampler.py:118-119)
//
// Pillow's resampler is integer arithmetic (libImaging/Resample.c): per output index a window
// [xmin, xmin+n) of 22-bit fixed-point taps, a horizontal pass rounded and clipped to uint8, then a
// vertical pass rounded and clipped again.  The tap tables are data independent; the host builds
// them exactly like Pillow does (api/utils/pil_tables.py) and uploads them once, so the kernels only
// do the integer multiply-accumulate and are BIT-EXACT with the PIL pipeline (tests/golden/preproc_*).
// uint8 crops are 16x smaller than the fp32 tensors the reference ships to the GPU (37.6 KB per
// 112x112x3 frame instead of 602 KB RGB + window copies), which is what makes the host-facing path
// PCIe-cheap; the RGB kernel can write the space-to-depth'ed 16-bit conv1 operand directly.
#include "common.cuh"
#include "conv_engine.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <vector>

namespace mimamo {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kPreThreads = 512;

struct PreprocDev {
  int src;                 // crop edge (112)
  int gsize, gk;           // gray size (48) and its tap count
  int rsize, rk;           // RGB resize (256) and its tap count
  int crop, off;           // centre crop (224) and its offset in the resized image
  const int* g_bounds;     // [gsize][2] (first input index, tap count)
  const int* g_kk;         // [gsize][gk]
  const int* r_bounds;     // [rsize][2]
  const int* r_kk;         // [rsize][rk]
  float mean[3];
};

__device__ __forceinline__ uint32_t clip8(int acc) {
  const int v = acc >> kPrecisionBits;
  return (uint32_t)min(max(v, 0), 255);
}

// block-wide copy of `bytes` (multiple of 4, 4-byte aligned source) into shared memory
__device__ __forceinline__ void copy_to_smem(uint8_t* dst, const uint8_t* __restrict__ src, int bytes) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) d[i] = __ldg(s + i);
}

// ---- gray: one CTA per crop ---------------------------------------------------------------------
__global__ void __launch_bounds__(256)
crops_gray_kernel(const __grid_constant__ PreprocDev P, const uint8_t* __restrict__ crops, float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int S = P.src, G = P.gsize, K = P.gk;
  int* kk = reinterpret_cast<int*>(sm);                       // [G][K]
  int* bounds = kk + G * K;                                   // [G][2]
  uint8_t* rgb = reinterpret_cast<uint8_t*>(bounds + 2 * G);  // [S][S][3]
  uint8_t* lum = rgb + (size_t)S * S * 3;                     // [S][S]
  uint8_t* tmp = lum + (size_t)S * S;                         // [S][G]
  const size_t n = blockIdx.x;
  for (int i = threadIdx.x; i < G * K; i += blockDim.x) kk[i] = P.g_kk[i];
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) bounds[i] = P.g_bounds[i];
  copy_to_smem(rgb, crops + n * (size_t)S * S * 3, S * S * 3);
  __syncthreads();
  // Convert.c rgb2l
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
    const uint32_t r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    lum[i] = (uint8_t)((r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * G; i += blockDim.x) {     // horizontal pass
    const int y = i / G, xx = i - y * G;
    const int x0 = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    int acc = 1 << (kPrecisionBits - 1);
    for (int k = 0; k < cnt; ++k) acc += (int)lum[y * S + x0 + k] * kk[xx * K + k];
    tmp[i] = (uint8_t)clip8(acc);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * G; i += blockDim.x) {     // vertical pass, then float / 255
    const int yy = i / G, x = i - yy * G;
    const int y0 = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    int acc = 1 << (kPrecisionBits - 1);
    for (int k = 0; k < cnt; ++k) acc += (int)tmp[(y0 + k) * G + x] * kk[yy * K + k];
    out[n * (size_t)G * G + i] = __fdiv_rn((float)clip8(acc), 255.f);
  }
}

// ---- RGB: one CTA per crop ----------------------------------------------------------------------
// MODE 0: fp32 NCHW [n][3][crop][crop] (what the reference's Image_Sampler yields);
// MODE 1 / 2: bf16 / fp16 space-to-depth'ed, chunk-planar conv1 operand [n][115][2][115][8] (nn_kernels.cuh).
template <int MODE>
__global__ void __launch_bounds__(kPreThreads)
crops_rgb_kernel(const __grid_constant__ PreprocDev P, const uint8_t* __restrict__ crops, void* __restrict__ out_) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int S = P.src, R = P.rsize, K = P.rk, C = P.crop, off = P.off;
  int* kk = reinterpret_cast<int*>(sm);                       // [R][K]
  int* bounds = kk + R * K;                                   // [R][2]
  float* lut = reinterpret_cast<float*>(bounds + 2 * R);      // [3][256]: (u8 / 255) * 255 - mean
  uint8_t* rgb = reinterpret_cast<uint8_t*>(lut + 768);       // [S][S][3]
  uint8_t* tmp = rgb + (size_t)S * S * 3;                     // [S][C][3]: horizontally resized, cropped columns
  const size_t n = blockIdx.x;
  for (int i = threadIdx.x; i < R * K; i += blockDim.x) kk[i] = P.r_kk[i];
  for (int i = threadIdx.x; i < 2 * R; i += blockDim.x) bounds[i] = P.r_bounds[i];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    // ToTensor (/255), the x * 255.0 lambda, Normalize(mean, std = 1): three separately rounded fp32 operations
    const float v = __fmul_rn(__fdiv_rn((float)(i & 255), 255.f), 255.f);
    lut[i] = __fdiv_rn(__fsub_rn(v, P.mean[i >> 8]), 1.f);
  }
  copy_to_smem(rgb, crops + n * (size_t)S * S * 3, S * S * 3);
  __syncthreads();
  for (int i = threadIdx.x; i < S * C; i += blockDim.x) {     // horizontal pass over the kept columns
    const int y = i / C, xo = i - y * C;
    const int xx = xo + off;
    const int x0 = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    const uint8_t* row = rgb + ((size_t)y * S + x0) * 3;
    for (int k = 0; k < cnt; ++k) {
      const int w = kk[xx * K + k];
      a0 += (int)row[3 * k] * w; a1 += (int)row[3 * k + 1] * w; a2 += (int)row[3 * k + 2] * w;
    }
    uint8_t* t = tmp + (size_t)i * 3;
    t[0] = (uint8_t)clip8(a0); t[1] = (uint8_t)clip8(a1); t[2] = (uint8_t)clip8(a2);
  }
  __syncthreads();
  // vertical pass of output pixel (h, w) of the cropped image, all three channels
  auto pixel = [&](int h, int w, float (&v)[3]) {
    const int yy = h + off;
    const int y0 = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int k = 0; k < cnt; ++k) {
      const int wk = kk[yy * K + k];
      const uint8_t* t = tmp + ((size_t)(y0 + k) * C + w) * 3;
      a0 += (int)t[0] * wk; a1 += (int)t[1] * wk; a2 += (int)t[2] * wk;
    }
    v[0] = lut[clip8(a0)]; v[1] = lut[256 + clip8(a1)]; v[2] = lut[512 + clip8(a2)];
  };
  if (MODE == 0) {
    float* out = reinterpret_cast<float*>(out_) + n * (size_t)3 * C * C;
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
      float v[3];
      pixel(i / C, i % C, v);
      out[i] = v[0]; out[(size_t)C * C + i] = v[1]; out[(size_t)2 * C * C + i] = v[2];
    }
  } else {
    const int D = (C + 6) / 2;                                // 115 for crop 224
    uint4* out = reinterpret_cast<uint4*>(out_) + n * (size_t)D * D * 2;
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
      const int Y = i / D, X = i - Y * D;
      uint16_t h16[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int hi = 2 * Y + (q >> 1) - 4, wi = 2 * X + (q & 1) - 4;
        float v[3] = {0.f, 0.f, 0.f};
        if (hi >= 0 && hi < C && wi >= 0 && wi < C) pixel(hi, wi, v);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (MODE == 1) { __nv_bfloat16 b = __float2bfloat16(v[c]); h16[q * 3 + c] = *reinterpret_cast<uint16_t*>(&b); }
          else { __half b = __float2half(v[c]); h16[q * 3 + c] = *reinterpret_cast<uint16_t*>(&b); }
        }
      }
      uint4 o0, o1;
      o0.x = h16[0] | ((uint32_t)h16[1] << 16); o0.y = h16[2] | ((uint32_t)h16[3] << 16);
      o0.z = h16[4] | ((uint32_t)h16[5] << 16); o0.w = h16[6] | ((uint32_t)h16[7] << 16);
      o1.x = h16[8] | ((uint32_t)h16[9] << 16); o1.y = h16[10] | ((uint32_t)h16[11] << 16);
      o1.z = 0; o1.w = 0;
      out[((size_t)Y * 2) * D + X] = o0;                        // chunk-planar rows: [Y][chunk][X]
      out[((size_t)Y * 2 + 1) * D + X] = o1;
    }
  }
}


// ---- RGB, 3-tap (up-sampling) fast path ---------------------------------------------------------
// Pillow's bilinear filter has support 1 when enlarging, i.e. at most three taps per output index.
// Each output index then needs one int4 {first input, k0, k1, k2} (absent taps are zero), pixels are
// moved as aligned 32-bit words (four RGB pixels = three words) instead of single bytes, and a thread
// produces four horizontally adjacent pixels at a time: ~2.5x fewer shared-memory instructions than
// the generic kernel above, same integers.
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int b) { return (w >> (8 * b)) & 0xffu; }

template <int MODE>
__global__ void __launch_bounds__(kPreThreads)
crops_rgb3_kernel(const __grid_constant__ PreprocDev P, const uint8_t* __restrict__ crops, void* __restrict__ out_) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int S = P.src, R = P.rsize, C = P.crop, off = P.off;
  int4* tab = reinterpret_cast<int4*>(sm);                    // [R] {first, k0, k1, k2}
  float* lut = reinterpret_cast<float*>(tab + R);             // [3][256]
  uint8_t* rgb = reinterpret_cast<uint8_t*>(lut + 768);       // [S][S][3] + 16 B slack: zero-weight taps of the last pixels read (and ignore) up to 6 B past the image
  uint8_t* tmp = rgb + (((size_t)S * S * 3 + 15) & ~(size_t)15) + 16;   // [S][C][3]
  const size_t n = blockIdx.x;
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const int cnt = P.r_bounds[2 * i + 1];
    int4 t;
    t.x = P.r_bounds[2 * i];
    t.y = cnt > 0 ? P.r_kk[3 * i] : 0; t.z = cnt > 1 ? P.r_kk[3 * i + 1] : 0; t.w = cnt > 2 ? P.r_kk[3 * i + 2] : 0;
    tab[i] = t;
  }
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const float v = __fmul_rn(__fdiv_rn((float)(i & 255), 255.f), 255.f);
    lut[i] = __fdiv_rn(__fsub_rn(v, P.mean[i >> 8]), 1.f);
  }
  copy_to_smem(rgb, crops + n * (size_t)S * S * 3, S * S * 3);
  __syncthreads();
  const uint32_t* rgbw = reinterpret_cast<const uint32_t*>(rgb);
  uint32_t* tmpw = reinterpret_cast<uint32_t*>(tmp);
  const int C4 = C / 4;
  // horizontal pass: item = (input row y, four consecutive kept columns)
  for (int i = threadIdx.x; i < S * C4; i += blockDim.x) {
    const int y = i / C4, g = i - y * C4;
    uint32_t o[12];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int4 t = tab[4 * g + q + off];
      const int addr = (y * S + t.x) * 3;
      const int sh = (addr & 3) * 8;
      const uint32_t* w = rgbw + (addr >> 2);
      const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
      const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = w2 >> sh;   // bytes 0..3, 4..7, 8
      const int h = 1 << (kPrecisionBits - 1);
      const int a0 = h + (int)byte_of(v0, 0) * t.y + (int)byte_of(v0, 3) * t.z + (int)byte_of(v1, 2) * t.w;
      const int a1 = h + (int)byte_of(v0, 1) * t.y + (int)byte_of(v1, 0) * t.z + (int)byte_of(v1, 3) * t.w;
      const int a2 = h + (int)byte_of(v0, 2) * t.y + (int)byte_of(v1, 1) * t.z + (int)byte_of(v2, 0) * t.w;
      o[3 * q] = clip8(a0); o[3 * q + 1] = clip8(a1); o[3 * q + 2] = clip8(a2);
    }
    uint32_t* dst = tmpw + (size_t)i * 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) dst[j] = o[4 * j] | (o[4 * j + 1] << 8) | (o[4 * j + 2] << 16) | (o[4 * j + 3] << 24);
  }
  __syncthreads();
  // vertical pass of row h, columns 4*g .. 4*g+3: twelve interleaved RGB bytes
  auto quad = [&](int h, int g, uint32_t (&u)[12]) {
    const int4 t = tab[h + off];
    int acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 1 << (kPrecisionBits - 1);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int wk = k == 0 ? t.y : (k == 1 ? t.z : t.w);
      const uint32_t* w = tmpw + ((size_t)min(t.x + k, S - 1) * C4 + g) * 3;    // absent taps have zero weight; keep the address inside tmp
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint32_t v = w[j];
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[4 * j + b] += (int)byte_of(v, b) * wk;
      }
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) u[j] = clip8(acc[j]);
  };
  if (MODE == 0) {
    float* out = reinterpret_cast<float*>(out_) + n * (size_t)3 * C * C;
    for (int i = threadIdx.x; i < C * C4; i += blockDim.x) {
      const int h = i / C4, g = i - h * C4;
      uint32_t u[12];
      quad(h, g, u);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4 v = make_float4(lut[c * 256 + u[c]], lut[c * 256 + u[3 + c]], lut[c * 256 + u[6 + c]], lut[c * 256 + u[9 + c]]);
        *reinterpret_cast<float4*>(out + ((size_t)c * C + h) * C + 4 * g) = v;
      }
    }
  } else {
    // item = (s2d row Y, pair of s2d pixels X = 2*gx, 2*gx+1) = image rows 2Y-4, 2Y-3 x columns 4*gx-4 .. 4*gx-1
    const int D = (C + 6) / 2, DP = (D + 1) / 2;
    uint4* out = reinterpret_cast<uint4*>(out_) + n * (size_t)D * D * 2;
    for (int i = threadIdx.x; i < D * DP; i += blockDim.x) {
      const int Y = i / DP, gx = i - Y * DP;
      const int g = gx - 1;                                    // column group of the cropped image
      uint16_t h16[2][16];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int j = 0; j < 16; ++j) h16[a][j] = 0;
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        const int hi = 2 * Y + py - 4;
        if (hi < 0 || hi >= C || g < 0 || g >= C4) continue;
        uint32_t u[12];
        quad(hi, g, u);
#pragma unroll
        for (int px4 = 0; px4 < 4; ++px4)                       // column 4g + px4 -> s2d pixel (px4 >> 1), px = px4 & 1
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float v = lut[c * 256 + u[3 * px4 + c]];
            uint16_t bits;
            if (MODE == 1) { __nv_bfloat16 b = __float2bfloat16(v); bits = *reinterpret_cast<uint16_t*>(&b); }
            else { __half b = __float2half(v); bits = *reinterpret_cast<uint16_t*>(&b); }
            h16[px4 >> 1][(py * 2 + (px4 & 1)) * 3 + c] = bits;
          }
      }
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int X = 2 * gx + a;
        if (X >= D) continue;
        uint4 o0, o1;
        o0.x = h16[a][0] | ((uint32_t)h16[a][1] << 16); o0.y = h16[a][2] | ((uint32_t)h16[a][3] << 16);
        o0.z = h16[a][4] | ((uint32_t)h16[a][5] << 16); o0.w = h16[a][6] | ((uint32_t)h16[a][7] << 16);
        o1.x = h16[a][8] | ((uint32_t)h16[a][9] << 16); o1.y = h16[a][10] | ((uint32_t)h16[a][11] << 16);
        o1.z = 0; o1.w = 0;
        out[((size_t)Y * 2) * D + X] = o0;                      // chunk-planar rows: [Y][chunk][X]
        out[((size_t)Y * 2 + 1) * D + X] = o1;
      }
    }
  }
}

}  // namespace mimamo

using namespace mimamo;

struct mimamo_preproc {
  PreprocDev d;
  int* dev_blob = nullptr;
  size_t gray_smem = 0, rgb_smem = 0, rgb3_smem = 0;
  bool fast3 = false;      // 3-tap RGB fast path usable
};

extern "C" void mimamo_preproc_destroy(mimamo_preproc* p) {
  if (!p) return;
  cudaFree(p->dev_blob);
  delete p;
}

extern "C" int mimamo_preproc_create(int32_t src, int32_t gray_size, int32_t gray_taps, const int32_t* gray_bounds_host,
                                     const int32_t* gray_kk_host, int32_t resize, int32_t rgb_taps,
                                     const int32_t* rgb_bounds_host, const int32_t* rgb_kk_host, int32_t crop,
                                     int32_t crop_off, const float* mean_host, mimamo_preproc** out) {
  MM_REQUIRE(out && gray_bounds_host && gray_kk_host && rgb_bounds_host && rgb_kk_host && mean_host, MIMAMO_E_VALUE, "null argument");
  MM_REQUIRE(src >= 1 && (src * src * 3) % 4 == 0, MIMAMO_E_VALUE, "crop edge %d: src*src*3 must be a multiple of 4", src);
  MM_REQUIRE(gray_size >= 1 && gray_taps >= 1 && crop >= 1 && rgb_taps >= 1 && crop_off >= 0 && crop_off + crop <= resize, MIMAMO_E_VALUE, "bad preprocessing geometry");
  auto check = [&](const int32_t* b, int n, int taps) {
    for (int i = 0; i < n; ++i)
      if (b[2 * i] < 0 || b[2 * i + 1] < 0 || b[2 * i + 1] > taps || b[2 * i] + b[2 * i + 1] > src) return false;
    return true;
  };
  MM_REQUIRE(check(gray_bounds_host, gray_size, gray_taps) && check(rgb_bounds_host, resize, rgb_taps), MIMAMO_E_VALUE,
             "tap windows leave the source image");
  mimamo_preproc* p = new mimamo_preproc();
  PreprocDev& d = p->d;
  d.src = src; d.gsize = gray_size; d.gk = gray_taps; d.rsize = resize; d.rk = rgb_taps; d.crop = crop;
  d.off = crop_off;                                           // torchvision center_crop: int(round((h - th) / 2.)), computed by the caller
  for (int c = 0; c < 3; ++c) d.mean[c] = mean_host[c];
  const size_t n_int = (size_t)gray_size * (2 + gray_taps) + (size_t)resize * (2 + rgb_taps);
  if (cudaMalloc((void**)&p->dev_blob, n_int * sizeof(int)) != cudaSuccess) {
    set_error("cudaMalloc of the resampling tables failed");
    delete p;
    return MIMAMO_E_CUDA;
  }
  int* cur = p->dev_blob;
  auto put = [&](const int32_t* h, size_t n) { cudaMemcpy(cur, h, n * sizeof(int), cudaMemcpyHostToDevice); const int* at = cur; cur += n; return at; };
  d.g_bounds = put(gray_bounds_host, (size_t)gray_size * 2);
  d.g_kk = put(gray_kk_host, (size_t)gray_size * gray_taps);
  d.r_bounds = put(rgb_bounds_host, (size_t)resize * 2);
  d.r_kk = put(rgb_kk_host, (size_t)resize * rgb_taps);
  p->gray_smem = (size_t)gray_size * (2 + gray_taps) * sizeof(int) + (size_t)src * src * 4 + (size_t)src * gray_size + 16;
  p->rgb_smem = (size_t)resize * (2 + rgb_taps) * sizeof(int) + 768 * sizeof(float) + (size_t)src * src * 3 + (size_t)src * crop * 3 + 16;
  p->rgb3_smem = (size_t)resize * 16 + 768 * sizeof(float) + (((size_t)src * src * 3 + 15) & ~(size_t)15) + 16 + (size_t)src * crop * 3 + 16;
  {
    const char* e = getenv("MIMAMO_PREPROC_GENERIC");
    p->fast3 = rgb_taps == 3 && crop % 4 == 0 && !(e && e[0] == '1');
  }
  if (p->gray_smem > 200 * 1024 || p->rgb_smem > 200 * 1024 || p->rgb3_smem > 200 * 1024) {
    set_error("crop edge %d too large for the shared-memory resampler", src);
    mimamo_preproc_destroy(p);
    return MIMAMO_E_RUNTIME;
  }
  cudaError_t e = cudaFuncSetAttribute(crops_gray_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->gray_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(crops_rgb_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->rgb_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(crops_rgb_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->rgb_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(crops_rgb_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->rgb_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(crops_rgb3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->rgb3_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(crops_rgb3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->rgb3_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(crops_rgb3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->rgb3_smem);
  if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    set_error("preprocessing plan setup failed: %s", cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
    mimamo_preproc_destroy(p);
    return MIMAMO_E_CUDA;
  }
  *out = p;
  return MIMAMO_OK;
}

extern "C" int mimamo_preproc_geometry(const mimamo_preproc* p, int32_t* src, int32_t* gray_size, int32_t* crop) {
  MM_REQUIRE(p, MIMAMO_E_VALUE, "null plan");
  if (src) *src = p->d.src;
  if (gray_size) *gray_size = p->d.gsize;
  if (crop) *crop = p->d.crop;
  return MIMAMO_OK;
}

extern "C" int mimamo_crops_to_gray(const mimamo_preproc* p, const uint8_t* crops, int64_t n, float* out, void* stream) {
  MM_REQUIRE(p && (n == 0 || (crops && out)) && n >= 0 && n < (1ll << 31), MIMAMO_E_VALUE, "bad arguments");
  if (n == 0) return MIMAMO_OK;
  crops_gray_kernel<<<(unsigned)n, 256, p->gray_smem, (cudaStream_t)stream>>>(p->d, crops, out);
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}

namespace mimamo {
// mode 0: fp32 NCHW, 1: bf16 s2d, 2: fp16 s2d
int crops_rgb_launch(const mimamo_preproc* p, const uint8_t* crops, int64_t n, void* out, int mode, cudaStream_t s) {
  MM_REQUIRE(p && (n == 0 || (crops && out)) && n >= 0 && n < (1ll << 31), MIMAMO_E_VALUE, "bad arguments");
  MM_REQUIRE(mode == 0 || p->d.crop == 224, MIMAMO_E_RUNTIME, "the conv1 operand layout needs a 224x224 centre crop");
  if (n == 0) return MIMAMO_OK;
  if (p->fast3) {
    if (mode == 0) crops_rgb3_kernel<0><<<(unsigned)n, kPreThreads, p->rgb3_smem, s>>>(p->d, crops, out);
    else if (mode == 1) crops_rgb3_kernel<1><<<(unsigned)n, kPreThreads, p->rgb3_smem, s>>>(p->d, crops, out);
    else crops_rgb3_kernel<2><<<(unsigned)n, kPreThreads, p->rgb3_smem, s>>>(p->d, crops, out);
  } else if (mode == 0) crops_rgb_kernel<0><<<(unsigned)n, kPreThreads, p->rgb_smem, s>>>(p->d, crops, out);
  else if (mode == 1) crops_rgb_kernel<1><<<(unsigned)n, kPreThreads, p->rgb_smem, s>>>(p->d, crops, out);
  else crops_rgb_kernel<2><<<(unsigned)n, kPreThreads, p->rgb_smem, s>>>(p->d, crops, out);
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}
}  // namespace mimamo

extern "C" int mimamo_crops_to_rgb(const mimamo_preproc* p, const uint8_t* crops, int64_t n, float* out, void* stream) {
  return crops_rgb_launch(p, crops, n, out, 0, (cudaStream_t)stream);
}
