// Full complex steerable pyramid of arbitrary (un-mirrored) square images, and its inverse.
//
// Replaces SCFpyr_PyTorch.build / _build_levels (api/steerable/SCFpyr_PyTorch.py:70-208: hi0 residual, every oriented
// band at its full size, low residual) and SCFpyr_PyTorch.reconstruct / _reconstruct_levels (:214-314).  The inference
// hot path never calls these (Phase_Difference_Extractor only reads the oriented bands of mirror-extended frames, which
// pyramid.cu computes in the DCT domain); they complete the drop-in surface of the class and give the test suite the
// reference's own round-trip property reconstruct(build(x)) ~ x.
//
// Formulation.  Every output of the reference is  ifft2(ifftshift(crop(fftshift(fft2(x))) * mask))  for a real,
// data-independent mask; all shifts and crops are index permutations, so per output "unit" (hi0, one pyramid level,
// lo) the host (api/steerable/plan_tables.py::full_pyramid_tables) provides the level size s, the gather index
// src[u] of natural-order frequency u of the level into the natural-order spectrum of the image, and the cumulative
// mask in natural order.  On the device: one forward 2-D DFT per image, then per unit gather * mask * (-i)^twist and
// an inverse 2-D DFT of size s.  The inverse runs the same steps backwards with scatter-adds into one spectrum.
// A 1-D DFT pass is a tiled complex matrix product against an exact twiddle table (index j*k mod s, built in
// float64); sizes are arbitrary (the reference's 96 = 2^5*3, 224 = 2^5*7, odd sizes).
#include "common.cuh"
#include <math.h>
#include <vector>

namespace mimamo {

constexpr int kDftTile = 16;

// Y[b][j][r] = scale * sum_k X[b][r][k] * exp(sign * 2 pi i * j k / S)      (transposed store: two passes = 2-D DFT)
template <bool REAL_IN, bool REAL_OUT>
__global__ void __launch_bounds__(kDftTile * kDftTile)
dft_pass_kernel(const float* __restrict__ X, float* __restrict__ Y, int R, int S, float sign, float scale,
                const float2* __restrict__ tw) {
  extern __shared__ float2 s_tw[];                       // [S] exp(2 pi i j / S)
  __shared__ float2 Xs[kDftTile][kDftTile + 1];          // [r][k]
  for (int i = threadIdx.x; i < S; i += blockDim.x) s_tw[i] = tw[i];
  const int tr = threadIdx.x & (kDftTile - 1), tj = threadIdx.x / kDftTile;
  const int r0 = blockIdx.y * kDftTile, j0 = blockIdx.x * kDftTile;
  const long long b = blockIdx.z;
  const int j = j0 + tj, r = r0 + tr;
  const float* Xb = X + (size_t)b * R * S * (REAL_IN ? 1 : 2);
  const int lr = threadIdx.x / kDftTile, lk = threadIdx.x & (kDftTile - 1);     // load mapping: k fastest
  float acc_re = 0.f, acc_im = 0.f;
  for (int k0 = 0; k0 < S; k0 += kDftTile) {
    __syncthreads();
    float2 v = make_float2(0.f, 0.f);
    if (r0 + lr < R && k0 + lk < S) {
      if (REAL_IN) v.x = __ldg(Xb + (size_t)(r0 + lr) * S + k0 + lk);
      else v = __ldg(reinterpret_cast<const float2*>(Xb) + (size_t)(r0 + lr) * S + k0 + lk);
    }
    Xs[lr][lk] = v;
    __syncthreads();
    if (j < S) {
      int idx = (int)(((long long)j * k0) % S);
      float pr = 0.f, pi = 0.f;                              // partial sums per 16-term chunk keep the fp32 error growth small
#pragma unroll
      for (int kk = 0; kk < kDftTile; ++kk) {
        const float2 w = s_tw[idx];
        const float2 x = Xs[tr][kk];
        const float wi = sign * w.y;
        pr = fmaf(x.x, w.x, pr); pr = fmaf(-x.y, wi, pr);
        pi = fmaf(x.x, wi, pi);  pi = fmaf(x.y, w.x, pi);
        idx += j;
        if (idx >= S) idx -= S;
      }
      acc_re += pr; acc_im += pi;
    }
  }
  if (j < S && r < R) {
    if (REAL_OUT) Y[((size_t)b * S + j) * R + r] = acc_re * scale;
    else reinterpret_cast<float2*>(Y)[((size_t)b * S + j) * R + r] = make_float2(acc_re * scale, acc_im * scale);
  }
}

__device__ __forceinline__ float2 twist_mul(float2 v, int t) {      // v * (-i)^t
  switch (t & 3) {
    case 1: return make_float2(v.y, -v.x);
    case 2: return make_float2(-v.x, -v.y);
    case 3: return make_float2(-v.y, v.x);
    default: return v;
  }
}

// Z[p][n][u][v] = F[n][src[u]][src[v]] * M[p][u][v] * (-i)^twist
__global__ void spec_gather_kernel(const float2* __restrict__ F, int S, const int* __restrict__ src, int s, int planes, long long N,
                                   const float* __restrict__ M, int twist, float2* __restrict__ Z) {
  const long long total = (long long)planes * N * s * s;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % s);
    long long rest = i / s;
    const int u = (int)(rest % s);
    rest /= s;
    const long long n = rest % N;
    const int p = (int)(rest / N);
    const float2 f = __ldg(F + ((size_t)n * S + __ldg(src + u)) * S + __ldg(src + v));
    const float m = __ldg(M + ((size_t)p * s + u) * s + v);
    Z[i] = twist_mul(make_float2(f.x * m, f.y * m), twist);
  }
}

// acc[n][src[u]][src[v]] += sum_p B[p][n][u][v] * M[p][u][v] * (-i)^twist      (distinct (u,v) hit distinct cells)
__global__ void spec_scatter_kernel(const float2* __restrict__ B, int S, const int* __restrict__ src, int s, int planes, long long N,
                                    const float* __restrict__ M, int twist, float2* __restrict__ acc) {
  const long long total = N * s * s;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % s);
    long long rest = i / s;
    const int u = (int)(rest % s);
    const long long n = rest / s;
    float2 sum = make_float2(0.f, 0.f);
    for (int p = 0; p < planes; ++p) {
      const float2 b = __ldg(B + (((size_t)p * N + n) * s + u) * s + v);
      const float m = __ldg(M + ((size_t)p * s + u) * s + v);
      sum.x = fmaf(b.x, m, sum.x); sum.y = fmaf(b.y, m, sum.y);
    }
    sum = twist_mul(sum, twist);
    float2* dst = acc + ((size_t)n * S + __ldg(src + u)) * S + __ldg(src + v);
    float2 cur = *dst;
    cur.x += sum.x; cur.y += sum.y;
    *dst = cur;
  }
}

// Image means: the DC coefficient of an image with values in [0,1] is ~S^2/2 while every other coefficient is ~S/3, so
// summing the raw image in fp32 spends the mantissa on the mean (the low residual, which carries it times (S/s)^2, came
// out 5x less accurate than the reference's FFT).  The forward DFT therefore runs on x - mean and DC is restored exactly.
__global__ void image_mean_kernel(const float* __restrict__ x, long long plane, float* __restrict__ centred, float* __restrict__ mean) {
  __shared__ double red[32];
  const float* p = x + (size_t)blockIdx.x * plane;
  double acc = 0.0;
  for (long long i = threadIdx.x; i < plane; i += blockDim.x) acc += (double)__ldg(p + i);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    red[0] = tot / (double)plane;
  }
  __syncthreads();
  const float m = (float)red[0];
  if (threadIdx.x == 0) mean[blockIdx.x] = m;
  float* q = centred + (size_t)blockIdx.x * plane;
  for (long long i = threadIdx.x; i < plane; i += blockDim.x) q[i] = __ldg(p + i) - m;
}
__global__ void restore_dc_kernel(float2* __restrict__ F, long long plane, const float* __restrict__ mean, long long N) {
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (n < N) F[(size_t)n * plane].x += mean[n] * (float)plane;
}

struct ScfUnit {
  int s, planes, is_real, twist_build, twist_recon;
  int* src;             // [s]
  float* build_mask;    // [planes][s][s]
  float* recon_mask;
  float2* tw;           // [s] (shared between units of equal size)
};

}  // namespace mimamo

using namespace mimamo;

struct mimamo_scf_plan {
  int S = 0;
  std::vector<ScfUnit> units;
  std::vector<std::pair<int, float2*>> tw_tables;
  float2* tw_S = nullptr;
};

static float2* twiddles_for(mimamo_scf_plan* plan, int s) {
  for (auto& t : plan->tw_tables)
    if (t.first == s) return t.second;
  std::vector<float2> h((size_t)s);
  for (int j = 0; j < s; ++j) {
    const double a = 2.0 * M_PI * (double)j / (double)s;
    h[j] = make_float2((float)cos(a), (float)sin(a));
  }
  float2* d = nullptr;
  if (upload(&d, h.data(), h.size()) != MIMAMO_OK) return nullptr;
  plan->tw_tables.emplace_back(s, d);
  return d;
}

extern "C" void mimamo_scf_plan_destroy(mimamo_scf_plan* plan) {
  if (!plan) return;
  for (auto& u : plan->units) { cudaFree(u.src); cudaFree(u.build_mask); cudaFree(u.recon_mask); }
  for (auto& t : plan->tw_tables) cudaFree(t.second);
  delete plan;
}

extern "C" int mimamo_scf_plan_create(int32_t S, int32_t n_units, const mimamo_scf_unit_desc* units, mimamo_scf_plan** plan_out) {
  MM_REQUIRE(plan_out && units && n_units >= 2 && S >= 2 && S <= 4096, MIMAMO_E_VALUE, "bad pyramid plan geometry");
  mimamo_scf_plan* plan = new mimamo_scf_plan();
  plan->S = S;
  int rc = MIMAMO_OK;
  plan->tw_S = twiddles_for(plan, S);
  if (!plan->tw_S) rc = MIMAMO_E_CUDA;
  for (int i = 0; i < n_units && !rc; ++i) {
    const mimamo_scf_unit_desc& d = units[i];
    if (!(d.s >= 1 && d.s <= S && d.planes >= 1 && d.src_index_host && d.build_mask_host && d.recon_mask_host)) {
      set_error("bad pyramid unit %d", i);
      rc = MIMAMO_E_VALUE;
      break;
    }
    ScfUnit u = {};
    u.s = d.s; u.planes = d.planes; u.is_real = d.is_real; u.twist_build = d.twist_build; u.twist_recon = d.twist_recon;
    const size_t mcount = (size_t)d.planes * d.s * d.s;
    rc = upload(&u.src, d.src_index_host, (size_t)d.s);
    if (!rc) rc = upload(&u.build_mask, d.build_mask_host, mcount);
    if (!rc) rc = upload(&u.recon_mask, d.recon_mask_host, mcount);
    u.tw = twiddles_for(plan, d.s);
    if (!rc && !u.tw) rc = MIMAMO_E_CUDA;
    plan->units.push_back(u);
  }
  if (rc) { mimamo_scf_plan_destroy(plan); return rc; }
  *plan_out = plan;
  return MIMAMO_OK;
}

static size_t scf_plane_floats(const mimamo_scf_plan* plan, int64_t N) {
  size_t m = (size_t)N * plan->S * plan->S;
  for (auto& u : plan->units) {
    const size_t c = (size_t)u.planes * N * u.s * u.s;
    if (c > m) m = c;
  }
  return m * 2;                                          // complex
}

extern "C" int mimamo_scf_workspace_bytes(const mimamo_scf_plan* plan, int64_t N, size_t* bytes_out) {
  MM_REQUIRE(plan && bytes_out && N >= 0, MIMAMO_E_VALUE, "bad arguments");
  *bytes_out = 3 * align_up(scf_plane_floats(plan, N) * sizeof(float), 256) + align_up((size_t)(N > 0 ? N : 1) * sizeof(float), 256) + 256;
  return MIMAMO_OK;
}

template <bool REAL_IN, bool REAL_OUT>
static int dft_pass(const float* X, float* Y, long long batch, int R, int S, float sign, float scale, const float2* tw, cudaStream_t st) {
  for (long long b0 = 0; b0 < batch; b0 += 65535) {          // gridDim.z limit
    const long long nb = batch - b0 < 65535 ? batch - b0 : 65535;
    dim3 grid((unsigned)((S + kDftTile - 1) / kDftTile), (unsigned)((R + kDftTile - 1) / kDftTile), (unsigned)nb);
    dft_pass_kernel<REAL_IN, REAL_OUT><<<grid, kDftTile * kDftTile, (size_t)S * sizeof(float2), st>>>(
        X + (size_t)b0 * R * S * (REAL_IN ? 1 : 2), Y + (size_t)b0 * R * S * (REAL_OUT ? 1 : 2), R, S, sign, scale, tw);
    MM_LAUNCH_OK();
  }
  return MIMAMO_OK;
}

static int grid_for(long long total) {
  long long g = (total + 255) / 256;
  return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

extern "C" int mimamo_scf_build(const mimamo_scf_plan* plan, const float* images, int64_t N, float* const* unit_out,
                                void* workspace, size_t workspace_bytes, void* stream) {
  MM_REQUIRE(plan && unit_out && N >= 0, MIMAMO_E_VALUE, "bad arguments");
  if (N == 0) return MIMAMO_OK;
  size_t need = 0;
  mimamo_scf_workspace_bytes(plan, N, &need);
  MM_REQUIRE(images && workspace && workspace_bytes >= need, MIMAMO_E_VALUE, "workspace too small: need %zu bytes", need);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t plane = align_up(scf_plane_floats(plan, N) * sizeof(float), 256);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  float* F = reinterpret_cast<float*>(base);
  float* T1 = reinterpret_cast<float*>(base + plane);
  float* T2 = reinterpret_cast<float*>(base + 2 * plane);
  const int S = plan->S;
  // forward 2-D DFT of every image (torch.rfft(onesided=False), SCFpyr_PyTorch.py:110), natural frequency order
  float* means = reinterpret_cast<float*>(base + 3 * plane);
  image_mean_kernel<<<(unsigned)N, 256, 0, st>>>(images, (long long)S * S, T2, means);
  MM_LAUNCH_OK();
  int rc = dft_pass<true, false>(T2, T1, N, S, S, -1.f, 1.f, plan->tw_S, st);
  if (!rc) rc = dft_pass<false, false>(T1, F, N, S, S, -1.f, 1.f, plan->tw_S, st);
  if (!rc) {
    restore_dc_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(reinterpret_cast<float2*>(F), (long long)S * S, means, N);
    MM_LAUNCH_OK();
  }
  for (size_t i = 0; i < plan->units.size() && !rc; ++i) {
    const ScfUnit& u = plan->units[i];
    MM_REQUIRE(unit_out[i], MIMAMO_E_VALUE, "null output for pyramid unit %zu", i);
    const long long planes = (long long)u.planes * N;
    spec_gather_kernel<<<grid_for(planes * u.s * u.s), 256, 0, st>>>(reinterpret_cast<const float2*>(F), S, u.src, u.s, u.planes, N,
                                                                      u.build_mask, u.twist_build, reinterpret_cast<float2*>(T1));
    MM_LAUNCH_OK();
    const float inv = 1.0f / ((float)u.s * (float)u.s);      // torch.ifft normalisation
    rc = dft_pass<false, false>(T1, T2, planes, u.s, u.s, 1.f, 1.f, u.tw, st);
    if (rc) break;
    if (u.is_real) rc = dft_pass<false, true>(T2, unit_out[i], planes, u.s, u.s, 1.f, inv, u.tw, st);
    else rc = dft_pass<false, false>(T2, unit_out[i], planes, u.s, u.s, 1.f, inv, u.tw, st);
  }
  return rc;
}

extern "C" int mimamo_scf_reconstruct(const mimamo_scf_plan* plan, const float* const* unit_in, int64_t N, float* images_out,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  MM_REQUIRE(plan && unit_in && N >= 0, MIMAMO_E_VALUE, "bad arguments");
  if (N == 0) return MIMAMO_OK;
  size_t need = 0;
  mimamo_scf_workspace_bytes(plan, N, &need);
  MM_REQUIRE(images_out && workspace && workspace_bytes >= need, MIMAMO_E_VALUE, "workspace too small: need %zu bytes", need);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t plane = align_up(scf_plane_floats(plan, N) * sizeof(float), 256);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  float* ACC = reinterpret_cast<float*>(base);
  float* T1 = reinterpret_cast<float*>(base + plane);
  float* T2 = reinterpret_cast<float*>(base + 2 * plane);
  const int S = plan->S;
  MM_CUDA(cudaMemsetAsync(ACC, 0, (size_t)N * S * S * sizeof(float2), st));
  int rc = MIMAMO_OK;
  for (size_t i = 0; i < plan->units.size() && !rc; ++i) {
    const ScfUnit& u = plan->units[i];
    MM_REQUIRE(unit_in[i], MIMAMO_E_VALUE, "null input for pyramid unit %zu", i);
    const long long planes = (long long)u.planes * N;
    if (u.is_real) rc = dft_pass<true, false>(unit_in[i], T1, planes, u.s, u.s, -1.f, 1.f, u.tw, st);
    else rc = dft_pass<false, false>(unit_in[i], T1, planes, u.s, u.s, -1.f, 1.f, u.tw, st);
    if (!rc) rc = dft_pass<false, false>(T1, T2, planes, u.s, u.s, -1.f, 1.f, u.tw, st);
    if (rc) break;
    spec_scatter_kernel<<<grid_for((long long)N * u.s * u.s), 256, 0, st>>>(reinterpret_cast<const float2*>(T2), S, u.src, u.s, u.planes, N,
                                                                             u.recon_mask, u.twist_recon, reinterpret_cast<float2*>(ACC));
    MM_LAUNCH_OK();
  }
  if (!rc) rc = dft_pass<false, false>(ACC, T1, N, S, S, 1.f, 1.f, plan->tw_S, st);
  if (!rc) rc = dft_pass<false, true>(T1, images_out, N, S, S, 1.f, 1.0f / ((float)S * (float)S), plan->tw_S, st);
  return rc;
}
