// P0+P1: complex steerable pyramid of mirror-extended frames, one CTA per frame.
//
// Replaces symmetric_extension_batch (api/utils/phase_utils.py:116-129), SCFpyr_PyTorch.build /
// _build_levels (api/steerable/SCFpyr_PyTorch.py:70-208) and the stack/permute/crop of
// Phase_Difference_Extractor.build_pyramid (api/phase_difference_extractor.py:38-87).
//
// Formulation (derived in mimamo-net_b200/api/steerable/plan_tables.py): because the frame is
// mirror-extended before the FFT, its spectrum is a real 2-D DCT-II `C` times unit phases, and
// each oriented band collapses to four real matrix products against host-built tables:
//
//   Ct[l][k]      = sum_{m,n} X[m][n] dct[m][k] dct[n][l]                      (2 products)
//   U[half][k][x] = sum_l (Ct[l][k] * mask[b][ch][half][l][k]) * trig[sel][l][x]
//   out_ch[y][x]  = sum_{kk < 2hp} trig[kk][y] * U[kk][x]
//
// All operands are K-major so every product is  D[i][j] = sum_k A[k][i] B[k][j]  with float4
// loads along i and j; the frame's spectrum, the intermediates U and the frame itself never
// leave shared memory, and no mirrored copy / fftshift / crop is ever materialised.
#include "common.cuh"

namespace mimamo {

struct LevelDev {
  int c, h, hp, cp;
  const float* trig;    // [2][hp][cp]
  const float* masks;   // [nb][2][2][hp][hp]
  int inner_sel[2][2];
  int units_per_chunk;  // how many bands (4 * hp * cp floats each) fit the work region at once
};

struct PlanDev {
  int H, Hp, Kp, nb, n_levels;
  int work_floats;
  int large;            // 1: frame too big for shared memory -> spectrum / intermediates live in a per-CTA global scratch
  int use_tc;           // large frames: block products as 3xTF32 tcgen05 MMAs (default; MIMAMO_PYR_TC=0: the fp32-FMA block products)
  const float* dct_t;   // [Hp][Kp]
  LevelDev lv[MIMAMO_MAX_LEVELS];
};

struct OutPtrs { float* p[MIMAMO_MAX_LEVELS]; };

constexpr int kPyrThreads = 256;

__device__ __forceinline__ float4 ld4(const float* p, bool global) {
  return global ? __ldg(reinterpret_cast<const float4*>(p)) : *reinterpret_cast<const float4*>(p);
}

// Register-tiled product of K-major operands:  acc[r][c] += sum_k A[k*lda + r] * (mask[k*ldm + r]) * B[k*ldb + c],
// r in [0,TM), c in [0,TN).  The first version used 4x4 tiles: two or three 16-byte loads per 16 FMAs kept the kernel on
// the load/store pipe (~13 % of the fp32 peak at every frame size); 4x8 / 8x4 tiles halve the loads per FMA, and the
// outer product below shares its table operand between the real and the imaginary channel.
template <int TM, int TN, bool A_GLOBAL, bool B_GLOBAL, bool MASKED>
__device__ __forceinline__ void tile_mma(const float* __restrict__ A, int lda, const float* __restrict__ Mk, int ldm,
                                         const float* __restrict__ B, int ldb, int K, float (&acc)[TM][TN]) {
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    float av[TM], bv[TN];
#pragma unroll
    for (int q = 0; q < TM / 4; ++q) {
      float4 a = ld4(A + (size_t)k * lda + 4 * q, A_GLOBAL);
      if (MASKED) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(Mk + (size_t)k * ldm + 4 * q));
        a.x *= m.x; a.y *= m.y; a.z *= m.z; a.w *= m.w;
      }
      av[4 * q] = a.x; av[4 * q + 1] = a.y; av[4 * q + 2] = a.z; av[4 * q + 3] = a.w;
    }
#pragma unroll
    for (int q = 0; q < TN / 4; ++q) {
      const float4 b = ld4(B + (size_t)k * ldb + 4 * q, B_GLOBAL);
      bv[4 * q] = b.x; bv[4 * q + 1] = b.y; bv[4 * q + 2] = b.z; bv[4 * q + 3] = b.w;
    }
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
  }
}

template <int TM, int TN>
__device__ __forceinline__ void zero(float (&acc)[TM][TN]) {
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[r][c] = 0.f;
}

template <int TM, int TN>
__device__ __forceinline__ void store_rows(float* dst, int ld, const float (&acc)[TM][TN]) {
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int q = 0; q < TN / 4; ++q)
      *reinterpret_cast<float4*>(dst + (size_t)r * ld + 4 * q) = make_float4(acc[r][4 * q], acc[r][4 * q + 1], acc[r][4 * q + 2], acc[r][4 * q + 3]);
}

__device__ float block_sum(float v, float* scratch) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += scratch[i];
  __syncthreads();
  return tot;
}

// Forward half: frame -> Ct (kept in shared memory).  Xs / R1 live in the work region.  All leading dimensions are
// multiples of 8 (plan_tables pads), so 4x8 tiles need no edge handling.
__device__ void frame_spectrum(const PlanDev& P, const float* __restrict__ frame, float* Ct, float* W,
                               float* red) {
  const int H = P.H, Hp = P.Hp, Kp = P.Kp;
  float* Xs = W;                 // [Hp][Hp]  X[m][n], zero padded
  float* R1 = W + Hp * Hp;       // [Hp][Kp]  R1[n][k]
  float part = 0.f;
  for (int i = threadIdx.x; i < Hp * Hp; i += blockDim.x) {
    const int m = i / Hp, n = i - m * Hp;
    float v = 0.f;
    if (m < H && n < H) v = __ldg(frame + (size_t)m * H + n);
    Xs[i] = v;
    part += v;
  }
  // The frame mean only feeds C[0][0], which every band mask zeroes; removing it up front keeps
  // the fp32 accumulations small (better agreement with the fp64 truth).
  const float mean = block_sum(part, red) / (float)(H * H);
  for (int i = threadIdx.x; i < Hp * Hp; i += blockDim.x) {
    const int m = i / Hp, n = i - m * Hp;
    if (m < H && n < H) Xs[i] -= mean;
  }
  __syncthreads();
  // R1[n][k] = sum_m X[m][n] dct[m][k]
  {
    const int tn = Kp >> 3, tiles = (Hp >> 2) * tn;
    for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
      const int i0 = (t / tn) << 2, j0 = (t % tn) << 3;
      float acc[4][8];
      zero(acc);
      tile_mma<4, 8, false, true, false>(Xs + i0, Hp, nullptr, 0, P.dct_t + j0, Kp, Hp, acc);
      store_rows(R1 + (size_t)i0 * Kp + j0, Kp, acc);
    }
  }
  __syncthreads();
  // Ct[l][k] = sum_n dct[n][l] R1[n][k]
  {
    const int tn = Kp >> 3, tiles = (Kp >> 2) * tn;
    for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
      const int i0 = (t / tn) << 2, j0 = (t % tn) << 3;
      float acc[4][8];
      zero(acc);
      tile_mma<4, 8, true, false, false>(P.dct_t + i0, Kp, nullptr, 0, R1 + j0, Kp, Hp, acc);
      store_rows(Ct + (size_t)i0 * Kp + j0, Kp, acc);
    }
  }
  __syncthreads();
}

// Inner products of one chunk of bands: U[(b*2+ch)*2*hp + half*hp + k][x]  (both channels, both halves of every band).
__device__ void band_inner(const PlanDev& P, const LevelDev& L, const float* Ct, float* W, int band0, int n_bands) {
  const int hp = L.hp, cp = L.cp;
  const int tk = hp >> 2, tx = cp >> 3;
  const int per_job = tk * tx, tiles = n_bands * 4 * per_job;
  for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
    const int job = t / per_job, r = t - job * per_job;
    const int u = job >> 1, half = job & 1;                 // u = local (band, ch)
    const int unit = band0 * 2 + u, b = unit >> 1, ch = unit & 1;
    const int i0 = (r / tx) << 2, j0 = (r % tx) << 3;       // i: k, j: x
    const float* mask = L.masks + ((size_t)((b * 2 + ch) * 2 + half) * hp) * hp;
    const float* tab = L.trig + (size_t)L.inner_sel[ch][half] * hp * cp;
    float acc[4][8];
    zero(acc);
    tile_mma<4, 8, false, true, true>(Ct + i0, P.Kp, mask + i0, hp, tab + j0, cp, hp, acc);
    store_rows(W + ((size_t)u * 2 * hp + half * hp + i0) * cp + j0, cp, acc);
  }
}

// Outer products: out_ch[y][x] = sum_kk trig[kk][y] * U_ch[kk][x]; one thread owns an 8 (y) x 4 (x) tile of BOTH channels
// of a band (the table operand is shared), so the coefficient leaves as (re, im) -- or (phase, magnitude) -- pairs.
template <class Sink>
__device__ void band_outer(const LevelDev& L, const float* W, int band0, int n_bands, Sink sink) {
  const int hp = L.hp, cp = L.cp;
  const int ty = cp >> 3, tx = cp >> 2;
  const int per_job = ty * tx, tiles = n_bands * per_job;
  for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
    const int u = t / per_job, r = t - u * per_job;
    const int i0 = (r / tx) << 3, j0 = (r % tx) << 2;       // i: y, j: x
    float re[8][4], im[8][4];
    zero(re);
    zero(im);
    const float* A = L.trig + i0;
    const float* Bre = W + (size_t)(2 * u) * 2 * hp * cp + j0;
    const float* Bim = Bre + (size_t)2 * hp * cp;
#pragma unroll 2
    for (int k = 0; k < 2 * hp; ++k) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(A + (size_t)k * cp));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(A + (size_t)k * cp + 4));
      const float4 br = *reinterpret_cast<const float4*>(Bre + (size_t)k * cp);
      const float4 bi = *reinterpret_cast<const float4*>(Bim + (size_t)k * cp);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float rv[4] = {br.x, br.y, br.z, br.w}, iv[4] = {bi.x, bi.y, bi.z, bi.w};
#pragma unroll
      for (int y = 0; y < 8; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) { re[y][x] = fmaf(av[y], rv[x], re[y][x]); im[y][x] = fmaf(av[y], iv[x], im[y][x]); }
    }
    sink(band0 + u, i0, j0, re, im);
  }
}

// ---- large frames (operands in the per-CTA global scratch): block products staged through shared memory ----
// With the operands in global memory the register-tile loops above re-read every element of A and B from L1/L2 once per
// 4x8 tile (28-56 times at 224x224): the 224x224 configuration ran on the L2 -> SM path (16 TFLOP/s).  Here the CTA
// computes one (16*TM) x (16*TN) block at a time, K in chunks of 16 staged through shared memory (register prefetch of the
// next chunk), so each operand element crosses L2 -> SM once per block.  A thread's TN = 8 columns are two groups of four
// (j and j + 64), which keeps every float4 shared-memory read a contiguous 16-byte-per-lane wavefront.
constexpr int kBK = 16;
struct BlockSmem {
  float a[kBK][128];
  float b0[kBK][128];
  float b1[kBK][128];
};

// acc0[r][c] (+acc1) += sum_k A[k*lda + i(r)] * mask * B0[k*ldb + j(c)],   i(r) = i0 + lane_i(r), j(c) = j0 + lane_j(c)
// rows / columns beyond M / N read as zero.  BI = 16*TM, BJ = 16*TN; tiles of 8 are split as {4t..4t+3} u {64+4t..}.
template <int TM, int TN, bool MASKED, bool DUAL>
__device__ __forceinline__ void block_gemm(const float* __restrict__ A, int lda, const float* __restrict__ Mk, int ldm, int M,
                                           const float* __restrict__ B0, const float* __restrict__ B1, int ldb, int N, int K,
                                           int i0, int j0, BlockSmem& sm, float (&acc0)[TM][TN], float (&acc1)[TM][TN]) {
  constexpr int BI = 16 * TM, BJ = 16 * TN;
  constexpr int A4 = kBK * BI / 4 / kPyrThreads, B4 = kBK * BJ / 4 / kPyrThreads;      // float4 loads per thread per chunk
  static_assert(A4 >= 1 && B4 >= 1, "block too small for 256 threads");
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  float4 ra[A4], rb0[B4], rb1[DUAL ? B4 : 1];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < A4; ++q) {
      const int e = threadIdx.x + q * kPyrThreads, kk = e / (BI / 4), i = i0 + (e % (BI / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < K && i < M) {
        v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(k0 + kk) * lda + i));
        if (MASKED) {
          const float4 m = __ldg(reinterpret_cast<const float4*>(Mk + (size_t)(k0 + kk) * ldm + i));
          v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
        }
      }
      ra[q] = v;
    }
#pragma unroll
    for (int q = 0; q < B4; ++q) {
      const int e = threadIdx.x + q * kPyrThreads, kk = e / (BJ / 4), j = j0 + (e % (BJ / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f), w = v;
      if (k0 + kk < K && j < N) {
        v = __ldg(reinterpret_cast<const float4*>(B0 + (size_t)(k0 + kk) * ldb + j));
        if (DUAL) w = __ldg(reinterpret_cast<const float4*>(B1 + (size_t)(k0 + kk) * ldb + j));
      }
      rb0[q] = v;
      if (DUAL) rb1[q] = w;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kBK) {
#pragma unroll
    for (int q = 0; q < A4; ++q) {
      const int e = threadIdx.x + q * kPyrThreads;
      *reinterpret_cast<float4*>(&sm.a[e / (BI / 4)][(e % (BI / 4)) * 4]) = ra[q];
    }
#pragma unroll
    for (int q = 0; q < B4; ++q) {
      const int e = threadIdx.x + q * kPyrThreads;
      *reinterpret_cast<float4*>(&sm.b0[e / (BJ / 4)][(e % (BJ / 4)) * 4]) = rb0[q];
      if (DUAL) *reinterpret_cast<float4*>(&sm.b1[e / (BJ / 4)][(e % (BJ / 4)) * 4]) = rb1[q];
    }
    __syncthreads();
    if (k0 + kBK < K) fetch(k0 + kBK);                         // global loads of the next chunk fly during the FMAs
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float av[TM], bv[TN], cv[DUAL ? TN : 1];
#pragma unroll
      for (int q = 0; q < TM / 4; ++q) {
        const float4 a = *reinterpret_cast<const float4*>(&sm.a[kk][q * 64 + ti * 4]);
        av[4 * q] = a.x; av[4 * q + 1] = a.y; av[4 * q + 2] = a.z; av[4 * q + 3] = a.w;
      }
#pragma unroll
      for (int q = 0; q < TN / 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(&sm.b0[kk][q * 64 + tj * 4]);
        bv[4 * q] = b.x; bv[4 * q + 1] = b.y; bv[4 * q + 2] = b.z; bv[4 * q + 3] = b.w;
        if (DUAL) {
          const float4 c = *reinterpret_cast<const float4*>(&sm.b1[kk][q * 64 + tj * 4]);
          cv[4 * q] = c.x; cv[4 * q + 1] = c.y; cv[4 * q + 2] = c.z; cv[4 * q + 3] = c.w;
        }
      }
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) {
          acc0[r][c] = fmaf(av[r], bv[c], acc0[r][c]);
          if (DUAL) acc1[r][c] = fmaf(av[r], cv[c], acc1[r][c]);
        }
    }
    __syncthreads();
  }
}

// row / column of element (r, c) of a thread's tile inside its block (tiles of 8 are split in two groups of four)
__device__ __forceinline__ int tile_off(int t, int r) { return (r >> 2) * 64 + t * 4 + (r & 3); }

// C[i][j] = sum_k A[k][i] (* mask) B[k][j] for all i < M, j < N, written to C (leading dimension ldc); 64 x 128 blocks
template <bool MASKED>
__device__ void large_product(const float* A, int lda, const float* Mk, int ldm, int M, const float* B, int ldb, int N, int K,
                              float* C, int ldc, BlockSmem& sm) {
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  for (int i0 = 0; i0 < M; i0 += 64)
    for (int j0 = 0; j0 < N; j0 += 128) {
      float acc[4][8], unused[4][8];
      zero(acc);
      block_gemm<4, 8, MASKED, false>(A, lda, Mk, ldm, M, B, nullptr, ldb, N, K, i0, j0, sm, acc, unused);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = i0 + ti * 4 + r;
        if (i >= M) continue;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int j = j0 + q * 64 + tj * 4;
          if (j < N) *reinterpret_cast<float4*>(C + (size_t)i * ldc + j) = make_float4(acc[r][4 * q], acc[r][4 * q + 1], acc[r][4 * q + 2], acc[r][4 * q + 3]);
        }
      }
    }
}

template <bool LARGE>
__global__ void __launch_bounds__(kPyrThreads, 2)
pyr_build_kernel(const __grid_constant__ PlanDev P, const float* __restrict__ frames, int T,
                 const __grid_constant__ OutPtrs outs, const int* __restrict__ root, float* scratch, long long n_frames, int polar) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  __shared__ __align__(16) BlockSmem bsm_storage[LARGE ? 1 : 0 + !LARGE];     // only the large-frame path uses it
  BlockSmem& bsm = bsm_storage[0];
  // Small frames: one CTA per frame, everything in shared memory.  Large frames (H > ~128, e.g. the
  // 224x224 configuration): a persistent grid walks the frames and keeps the same buffers in a
  // private slice of a caller-provided global scratch (L2-resident), same code otherwise.
  float* Ct = LARGE ? scratch + (size_t)blockIdx.x * ((size_t)P.Kp * P.Kp + P.work_floats) : smem;
  float* W = Ct + (size_t)P.Kp * P.Kp;
  for (long long n = blockIdx.x; n < n_frames; n += gridDim.x) {
    if (root != nullptr && root[n] != (int)n) continue;    // duplicate of an earlier frame: its root's coefficients are reused
    const long long w = n / T;
    const int t = (int)(n - w * T);
    auto emit = [&](float2* out, int b, int c, int y, int x, float re, float im) {
      if (y >= c || x >= c) return;
      float2 v = make_float2(re, im);
      if (polar) {
        // (phase, magnitude) exactly as the phase tail computes them from (re, im): the fused paths store the polar form
        // once per distinct frame so that the 13 windows sharing it do not repeat the atan2 / sqrt
        const float ph = atan2f(v.y, v.x);
        const float mg = __fadd_rn(sqrtf(__fadd_rn(__fmul_rn(v.y, v.y), __fmul_rn(v.x, v.x))), 1e-10f);
        v = make_float2(ph, mg);
      }
      out[((((size_t)w * P.nb + b) * T + t) * c + y) * (size_t)c + x] = v;
    };
    if (LARGE) {
      // frame -> Xs (mean removed) in the scratch, then the same four products as block GEMMs
      const int H = P.H, Hp = P.Hp, Kp = P.Kp;
      float* Xs = W;
      float* R1 = W + (size_t)Hp * Hp;
      const float* frame = frames + (size_t)n * H * H;
      float part = 0.f;
      for (int i = threadIdx.x; i < Hp * Hp; i += blockDim.x) {
        const int m = i / Hp, nn = i - m * Hp;
        float v = 0.f;
        if (m < H && nn < H) v = __ldg(frame + (size_t)m * H + nn);
        Xs[i] = v;
        part += v;
      }
      const float mean = block_sum(part, red) / (float)(H * H);
      for (int i = threadIdx.x; i < Hp * Hp; i += blockDim.x) {
        const int m = i / Hp, nn = i - m * Hp;
        if (m < H && nn < H) Xs[i] -= mean;
      }
      __syncthreads();
      large_product<false>(Xs, Hp, nullptr, 0, Hp, P.dct_t, Kp, Kp, Hp, R1, Kp, bsm);          // R1[n][k] = sum_m X[m][n] dct[m][k]
      __syncthreads();
      large_product<false>(P.dct_t, Kp, nullptr, 0, Kp, R1, Kp, Kp, Hp, Ct, Kp, bsm);          // Ct[l][k] = sum_n dct[n][l] R1[n][k]
      __syncthreads();
      for (int li = 0; li < P.n_levels; ++li) {
        const LevelDev& L = P.lv[li];
        float2* out = reinterpret_cast<float2*>(outs.p[li]);
        const int c = L.c, hp = L.hp, cp = L.cp;
        for (int band0 = 0; band0 < P.nb; band0 += L.units_per_chunk) {
          const int n_bands = min(L.units_per_chunk, P.nb - band0);
          for (int job = 0; job < n_bands * 4; ++job) {          // U[(band, ch)][half*hp + k][x]
            const int u = job >> 1, half = job & 1, unit = band0 * 2 + u, b = unit >> 1, ch = unit & 1;
            const float* mask = L.masks + ((size_t)((b * 2 + ch) * 2 + half) * hp) * hp;
            const float* tab = L.trig + (size_t)L.inner_sel[ch][half] * hp * cp;
            large_product<true>(Ct, Kp, mask, hp, hp, tab, cp, cp, hp, W + ((size_t)u * 2 * hp + half * hp) * cp, cp, bsm);
          }
          __syncthreads();
          const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
          for (int u = 0; u < n_bands; ++u) {
            const float* Bre = W + (size_t)(2 * u) * 2 * hp * cp;
            const float* Bim = Bre + (size_t)2 * hp * cp;
            for (int i0 = 0; i0 < cp; i0 += 128)
              for (int j0 = 0; j0 < cp; j0 += 64) {
                float re[8][4], im[8][4];
                zero(re);
                zero(im);
                block_gemm<8, 4, false, true>(L.trig, cp, nullptr, 0, cp, Bre, Bim, cp, cp, 2 * hp, i0, j0, bsm, re, im);
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                  for (int q = 0; q < 4; ++q) emit(out, band0 + u, c, i0 + tile_off(ti, r), j0 + tj * 4 + q, re[r][q], im[r][q]);
              }
          }
          __syncthreads();
        }
      }
      continue;
    }
    frame_spectrum(P, frames + (size_t)n * P.H * P.H, Ct, W, red);
    for (int li = 0; li < P.n_levels; ++li) {
      const LevelDev& L = P.lv[li];
      float2* out = reinterpret_cast<float2*>(outs.p[li]);
      const int c = L.c;
      for (int band0 = 0; band0 < P.nb; band0 += L.units_per_chunk) {
        const int n_bands = min(L.units_per_chunk, P.nb - band0);
        band_inner(P, L, Ct, W, band0, n_bands);
        __syncthreads();
        band_outer(L, W, band0, n_bands, [&](int b, int y0, int x0, const float (&re)[8][4], const float (&im)[8][4]) {
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) emit(out, b, c, y0 + r, x0 + q, re[r][q], im[r][q]);
        });
        __syncthreads();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Large frames on the tensor cores: pyr_build_umma_kernel (tcgen05.mma kind::tf32, 3xTF32 split operands).
//
// At 224x224 the dense DCT-domain transform is 1.7 GFLOP per frame and the fp32-FMA block products above reach 29 TFLOP/s.
// The phase tolerance (1e-4 on an amplitude-weighted phase) needs fp32-class products, so every operand x is split into two
// TF32 values, hi = tf32(x) and lo = tf32(x - hi) (22 mantissa bits together), and each product is three tensor-core passes
// accumulated in fp32 in TMEM:  a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  (lo*lo is below fp32 resolution).
//   * operands are K-major in global memory in the sense of the FMA kernels (A[k][i], i contiguous), i.e. M/N-contiguous;
//     UMMA wants K-contiguous rows.  A K chunk of 32 is staged as [128 rows][32 k] SWIZZLE_128B planes (the descriptor format
//     the convolution engine uses; a 32-float row is one 128-byte swizzle atom row): a staging thread (q = lane & 7,
//     i4 = 4 * (warp & 7) + lane / 8) takes four float4 (k = 4q .. 4q+3, rows 4 i4 .. 4 i4 + 3), applies the band mask, splits,
//     and writes four 16-byte vectors (one per row) -- eight lanes with the same rows and q = 0..7 hit eight distinct 16-byte
//     bank groups of the swizzled row, so the transposing stores are conflict free;
//   * roles (19 warps): two loader warps bring the raw fp32 chunk global -> shared with cp.async, laid out as the staging
//     threads' private 16-byte pieces, two chunks deep, completion by cp.async.mbarrier.arrive; sixteen staging warps; one
//     MMA warp issues 12 UMMA 128 x N x 8 per B operand and chunk, completion by tcgen05.commit -> mbarrier.  Two plane
//     buffers (one for the two-B outer products); no CTA-wide barrier inside a block product;
//   * the accumulators (128 lanes x N columns; the outer products compute the real and the imaginary channel of a band
//     against the same trig rows into two column ranges) are read back with tcgen05.ld, thread = output row.
// One CTA per SM (224 KB of staging), persistent over the frames; Ct / U live in the per-CTA global scratch as before.
// Measured (configs[2], 3328 frames of 224x224): stage 216 -> 150 ms.  The block products are NOT tensor-bound (tensor pipe 14 %
// active): the split planes cost 28 bytes of shared-memory traffic per operand element (raw in/out, hi+lo planes written, each read
// 1.5x by the three passes), ~1.2 us of the 2.3 us a chunk takes; 63 ms of the 150 are outside the products (phase tail 24,
// polar conversion + coefficient stores 13).  Next: A operand through TMEM (tcgen05.st, no shared-memory traffic for it) and
// pre-split table planes fetched by cp.async.bulk.
// ---------------------------------------------------------------------------------------
namespace tcg {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {       // bounded spin: a pipeline bug traps instead of hanging the device
  uint32_t spins = 0, ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 27)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// K-major SWIZZLE_128B matrix descriptor: start address >> 4, LBO = 1 (unused), SBO = 1024 B (8 rows x 128 B), version 1, layout 2
__device__ __forceinline__ uint64_t desc128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// four K = 8 TF32 UMMAs over one 32-float (128-byte) K chunk: the start address advances by 32 bytes per step
__device__ __forceinline__ void umma_tf32_k32(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, uint32_t accumulate) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint64_t da = desc128(a_addr + ks * 32), db = desc128(b_addr + ks * 32);
    const uint32_t acc = ks == 0 ? accumulate : 1u;
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(da), "l"(db),
        "r"(idesc), "r"(acc) : "memory");
  }
}
// hi part of the TF32 split: x rounded to 10 mantissa bits (round half away from zero in the magnitude bits).  ptxas expands
// cvt.rna.tf32.f32 into ~10 instructions on sm_100a (the first versions of the staging loop spent most of their issue slots there);
// this is an add and a mask.  The lo part x - hi is rounded the same way: left as plain fp32 the tensor core would TRUNCATE it,
// a bias that accumulates linearly over K (measured: coefficient error 2.1e-6 -> 5.7e-6).
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
}  // namespace tcg

constexpr int kUmmaChunk = 32;                               // K per staged chunk: one 128-byte row of TF32
constexpr int kPlaneBytes = 128 * kUmmaChunk * 4;            // [128 rows][32 k] = 16 KB
constexpr int kRawSlotBytes = 3 * kPlaneBytes;               // raw fp32 chunk: A, mask | second B, B  (thread-private 16-byte pieces)
constexpr int kPlaneRegionBytes = 8 * kPlaneBytes;           // single B: 2 buffers x (a_hi, a_lo, b_hi, b_lo); two Bs: 1 buffer x 6 planes
constexpr int kUmmaSmemBytes = 2 * kRawSlotBytes + kPlaneRegionBytes + 1024 + 256;

struct UmmaCtx {
  uint8_t* raw;              // [2 slots][3 operands][4 kk][256 thread slots][16 B]: cp.async landing ring, two chunks deep
  uint8_t* planes;           // 1024-aligned staging planes (see kPlaneRegionBytes)
  uint64_t* free_bar;        // [2]: the MMAs that read staging buffer b have retired (tcgen05.commit)
  uint64_t* raw_full;        // [2]: every copy of the chunk in ring slot s has landed (cp.async.mbarrier.arrive by the 64 loader lanes)
  uint64_t* raw_empty;       // [2]: the 16 staging warps have read ring slot s
  uint64_t* planes_full;     // [2]: the 16 staging warps have written (and fenced) plane buffer b
  uint64_t* tmem_free;       // the 16 staging warps have read the previous block's accumulators out of TMEM
  uint32_t blocks;           // block products started so far (same value in every thread)
  uint32_t tmem;             // 256 columns: [0,128) first B operand, [128,256) second
  uint32_t uses[2];          // commits issued on free_bar[b] so far (same value in every thread)
  uint32_t ring;             // chunks pushed through the raw ring so far (same value in every thread); slot = ring & 1
  int debug;                 // timing experiments (MIMAMO_PYR_TC=3: no MMAs, 5: no output stores)
};

constexpr int kStagerThreads = 512;                          // warps 0-15: split / transpose / mask; thread 0 issues the MMAs
constexpr int kUmmaThreads = kStagerThreads + 96;            // warps 16, 17: cp.async loaders of the A-side and the B-side operands; warp 18: MMA issuer

// The accumulators of one 128 x nj block:  D0[i][j] (+ D1 with B1) = sum_k A[k][i0 + i] * mask * B[k][j0 + j], left in TMEM.
// Rows beyond M, columns beyond N and k beyond K are staged as zeros.  nj: UMMA N, a multiple of 16 up to 128.
// Roles: two LOADER warps copy global -> shared (cp.async, 16-byte pieces laid out per staging thread) into a two-chunk ring and
// signal an mbarrier when a chunk has landed; sixteen STAGING warps (warps 0-7 the A operand and its mask, 8-15 the B operand(s);
// thread (q = lane & 7, i4 = 4 * (warp & 7) + lane / 8) owns k = 4q .. 4q + 3 of rows 4 i4 .. 4 i4 + 3) turn their pieces into the
// swizzled K-major hi / lo planes and hand them to the MMA warp through an mbarrier; one elected thread of warp 18 issues the UMMAs.
// No CTA-wide barrier sits in the chunk loop: the first versions joined all staging threads with __syncthreads every chunk and let
// one of them issue the MMAs (several hundred serial instructions), which cost 2.8 us per chunk whatever the thread count, the
// prefetch depth or the way the loads were issued.
template <bool MASKED, bool DUAL>
__device__ __forceinline__ void umma_block(UmmaCtx& cx, const float* __restrict__ A, int lda, const float* __restrict__ Mk, int ldm, int M,
                                           const float* __restrict__ B0, const float* __restrict__ B1, int ldb, int N, int K, int i0, int j0, int nj) {
  static_assert(!(MASKED && DUAL), "the raw ring has three operand slabs");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (cx.debug & 8) return;                                  // timing experiment: everything but the block products
  const int nchunks = (K + kUmmaChunk - 1) / kUmmaChunk;
  const uint32_t block_no = cx.blocks++;
  if (warp == 18) {
    // ------------------------------- MMA issuer -------------------------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nj >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (block_no > 0) tcg::mbar_wait(cx.tmem_free, (block_no - 1u) & 1u);   // the previous block's accumulators have been read
    uint32_t buf = 1;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      buf = DUAL ? 0u : (buf ^ 1u);
      tcg::mbar_wait(&cx.planes_full[buf], cx.uses[buf] & 1u);
      tcg::fence_after();
      if (lane == 0 && (cx.debug & 2)) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tcg::smem_u32(&cx.free_bar[buf])) : "memory");
      } else if (lane == 0) {
        const uint32_t b = tcg::smem_u32(cx.planes + (size_t)buf * 4 * kPlaneBytes);
        const uint32_t ah = b, al = b + kPlaneBytes, bh = b + 2 * kPlaneBytes, bl = b + 3 * kPlaneBytes;
        tcg::umma_tf32_k32(cx.tmem, al, bh, idesc, c > 0 ? 1u : 0u);
        tcg::umma_tf32_k32(cx.tmem, ah, bl, idesc, 1u);
        tcg::umma_tf32_k32(cx.tmem, ah, bh, idesc, 1u);
        if (DUAL) {
          const uint32_t ch = b + 4 * kPlaneBytes, cl = b + 5 * kPlaneBytes;
          tcg::umma_tf32_k32(cx.tmem + 128, al, ch, idesc, c > 0 ? 1u : 0u);
          tcg::umma_tf32_k32(cx.tmem + 128, ah, cl, idesc, 1u);
          tcg::umma_tf32_k32(cx.tmem + 128, ah, ch, idesc, 1u);
        }
        tcg::commit(&cx.free_bar[buf]);
      }
      __syncwarp();
      ++cx.uses[buf];
    }
    return;
  }
  const bool is_b = warp >= 16 ? warp == 17 : warp >= 8;       // which operand side this thread works for
  const float* G = is_b ? B0 : A;
  const float* G2 = is_b ? B1 : Mk;
  const int ld = is_b ? ldb : lda, ld2 = is_b ? ldb : ldm;
  const int x0 = is_b ? j0 : i0, lim = is_b ? N : M;
  const bool two = is_b ? DUAL : MASKED;
  const uint32_t slab0 = tcg::smem_u32(cx.raw) + (is_b ? 2u * kPlaneBytes : 0u);     // slabs: [A][mask | B1][B0]
  const uint32_t slab2 = tcg::smem_u32(cx.raw) + kPlaneBytes;
  const int q = lane & 7;
  if (warp >= 16) {
    // ------------------------------- loader warp -------------------------------
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c, ++cx.ring) {
      const uint32_t slot = cx.ring & 1u, so = slot * kRawSlotBytes;
      if (cx.ring >= 2) tcg::mbar_wait(&cx.raw_empty[slot], ((cx.ring >> 1) - 1u) & 1u);
#pragma unroll 1
      for (int s = 0; s < 8; ++s) {                            // the 256 thread slots of this side, 32 per pass
        const int i4 = s * 4 + (lane >> 3);
        const int x = x0 + 4 * i4;
        const bool ok = x < lim && (!is_b || 4 * i4 < nj);
        const uint32_t dst = (uint32_t)(s * 32 + lane) * 16u + so;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = c * kUmmaChunk + 4 * q + kk;
          const bool v = ok && k < K;
          const size_t off = (size_t)(v ? k : 0) * ld + (v ? x : 0);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slab0 + dst + kk * 4096u), "l"(G + off), "r"(v ? 16u : 0u) : "memory");
          if (two) {
            const size_t off2 = (size_t)(v ? k : 0) * ld2 + (v ? x : 0);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slab2 + dst + kk * 4096u), "l"(G2 + off2), "r"(v ? 16u : 0u) : "memory");
          }
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tcg::smem_u32(&cx.raw_full[slot])) : "memory");
    }
    return;
  }
  // ------------------------------- staging warps -------------------------------
  const int i4 = (warp & 7) * 4 + (lane >> 3);
  const int tg = threadIdx.x & 255;
  const uint32_t raw0 = slab0 + tg * 16u, raw2 = slab2 + tg * 16u;
  auto lds4 = [&](uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
  };
  // rows 4 i4 + r (r = 0..3) of the [128][32] plane: the four k values of this thread form one 16-byte vector per row
  auto stage = [&](const float4 (&v)[4], uint8_t* hi_plane, uint8_t* lo_plane) {
    const float rows[4][4] = {{v[0].x, v[1].x, v[2].x, v[3].x}, {v[0].y, v[1].y, v[2].y, v[3].y},
                              {v[0].z, v[1].z, v[2].z, v[3].z}, {v[0].w, v[1].w, v[2].w, v[3].w}};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = 4 * i4 + r;
      const uint32_t off = (uint32_t)row * 128u + (uint32_t)((q ^ (row & 7)) << 4);
      uint4 h, l;
      h.x = tcg::tf32_hi(rows[r][0]); h.y = tcg::tf32_hi(rows[r][1]); h.z = tcg::tf32_hi(rows[r][2]); h.w = tcg::tf32_hi(rows[r][3]);
      l.x = tcg::tf32_hi(rows[r][0] - __uint_as_float(h.x)); l.y = tcg::tf32_hi(rows[r][1] - __uint_as_float(h.y));
      l.z = tcg::tf32_hi(rows[r][2] - __uint_as_float(h.z)); l.w = tcg::tf32_hi(rows[r][3] - __uint_as_float(h.w));
      *reinterpret_cast<uint4*>(hi_plane + off) = h;
      *reinterpret_cast<uint4*>(lo_plane + off) = l;
    }
  };
  uint32_t buf = 1;
  for (int c = 0; c < nchunks; ++c, ++cx.ring) {
    const uint32_t slot = cx.ring & 1u, so = slot * kRawSlotBytes;
    tcg::mbar_wait(&cx.raw_full[slot], (cx.ring >> 1) & 1u);  // the chunk has landed
    float4 r1[4], r2[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      r1[kk] = lds4(raw0 + so + kk * 4096u);
      if (two) r2[kk] = lds4(raw2 + so + kk * 4096u);
      if (MASKED && !is_b) { r1[kk].x *= r2[kk].x; r1[kk].y *= r2[kk].y; r1[kk].z *= r2[kk].z; r1[kk].w *= r2[kk].w; }
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tcg::smem_u32(&cx.raw_empty[slot])) : "memory");
    // single B: two plane buffers alternate; two Bs: one buffer.  Its previous MMAs must have retired before it is overwritten.
    buf = DUAL ? 0u : (buf ^ 1u);
    if (cx.uses[buf] > 0) tcg::mbar_wait(&cx.free_bar[buf], (cx.uses[buf] - 1u) & 1u);
    uint8_t* base = cx.planes + (size_t)buf * 4 * kPlaneBytes;
    if (!is_b) {
      stage(r1, base, base + kPlaneBytes);
    } else {
      stage(r1, base + 2 * kPlaneBytes, base + 3 * kPlaneBytes);
      if (DUAL) stage(r2, base + 4 * kPlaneBytes, base + 5 * kPlaneBytes);
    }
    tcg::fence_proxy_async_smem();                             // generic-proxy stores -> visible to the tensor core
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tcg::smem_u32(&cx.planes_full[buf])) : "memory");
    ++cx.uses[buf];
  }
  tcg::mbar_wait(&cx.free_bar[buf], (cx.uses[buf] - 1u) & 1u);   // every MMA of the block has retired: the accumulators are complete
  tcg::fence_after();
}

// block column width for an N-column product: the fewest blocks of at most 128 columns, equal widths rounded up to 16
__device__ __forceinline__ int umma_bj(int N) {
  const int nblk = (N + 127) / 128;
  return (((N + nblk - 1) / nblk) + 15) & ~15;
}

// C[i][j] = sum_k A[k][i] (* mask) B[k][j] for all i < M, j < N, written to C (leading dimension ldc)
// (columns [jb, je) only: the band loop below works one block column of the output at a time)
template <bool MASKED>
__device__ void umma_product(UmmaCtx& cx, const float* A, int lda, const float* Mk, int ldm, int M, const float* B, int ldb, int N, int K,
                             float* C, int ldc, int jb, int je) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bj = umma_bj(N);
  for (int i0 = 0; i0 < M; i0 += 128)
    for (int j0 = jb; j0 < je; j0 += bj) {
      umma_block<MASKED, false>(cx, A, lda, Mk, ldm, M, B, nullptr, ldb, N, K, i0, j0, bj);
      if (warp >= 16) continue;                                // loader warps run ahead into the next block
      const int i = i0 + 32 * (warp & 3) + lane;              // thread = accumulator row (TMEM lane); warp / 4 selects 32 of the columns
      const uint32_t taddr = cx.tmem + ((uint32_t)(32 * (warp & 3)) << 16);
      for (int t0 = 32 * (warp >> 2); t0 < bj && t0 < 32 * (warp >> 2) + 32; t0 += 16) {
        float v[16];
        tcg::tmem_ld16(taddr + t0, v);                         // warp-collective: every lane takes part, stores are predicated
        if (i < M) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int j = j0 + t0 + 4 * g;
            if (j < N) *reinterpret_cast<float4*>(C + (size_t)i * ldc + j) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          }
        }
      }
      tcg::fence_before();                                     // the next block's first MMA overwrites the accumulators: the MMA warp
      __syncwarp();                                            // waits for all sixteen warps
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tcg::smem_u32(cx.tmem_free)) : "memory");
    }
}

__global__ void __launch_bounds__(kUmmaThreads, 1)
pyr_build_umma_kernel(const __grid_constant__ PlanDev P, const float* __restrict__ frames, int T, const __grid_constant__ OutPtrs outs,
                      const int* __restrict__ root, float* scratch, long long n_frames, int polar) {
  extern __shared__ uint8_t umma_smem_raw[];
  __shared__ float red[32];
  __shared__ uint32_t tmem_slot;
  UmmaCtx cx;
  cx.planes = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(umma_smem_raw) + 1023) & ~(uintptr_t)1023);
  cx.raw = cx.planes + kPlaneRegionBytes;
  cx.free_bar = reinterpret_cast<uint64_t*>(cx.raw + 2 * kRawSlotBytes);
  cx.raw_full = cx.free_bar + 2;
  cx.raw_empty = cx.raw_full + 2;
  cx.planes_full = cx.raw_empty + 2;
  cx.tmem_free = cx.planes_full + 2;
  cx.uses[0] = cx.uses[1] = 0;
  cx.ring = 0;
  cx.blocks = 0;
  cx.debug = P.use_tc;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      tcg::mbar_init(&cx.free_bar[i], 1);
      tcg::mbar_init(&cx.raw_full[i], 64);                    // both loader warps' lanes (cp.async.mbarrier.arrive.noinc)
      tcg::mbar_init(&cx.raw_empty[i], 16);                   // one arrival per staging warp
      tcg::mbar_init(&cx.planes_full[i], 16);
    }
    tcg::mbar_init(cx.tmem_free, 16);
    tcg::fence_barrier_init();
  }
  if (threadIdx.x < 32) tcg::tmem_alloc(&tmem_slot, 256);
  tcg::fence_before();
  __syncthreads();
  tcg::fence_after();
  cx.tmem = tmem_slot;
  float* Ct = scratch + (size_t)blockIdx.x * ((size_t)P.Kp * P.Kp + P.work_floats);
  float* W = Ct + (size_t)P.Kp * P.Kp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long n = blockIdx.x; n < n_frames; n += gridDim.x) {
    if (root != nullptr && root[n] != (int)n) continue;    // duplicate of an earlier frame: its root's coefficients are reused
    const long long w = n / T;
    const int t = (int)(n - w * T);
    const int H = P.H, Hp = P.Hp, Kp = P.Kp;
    float* Xs = W;
    float* R1 = W + (size_t)Hp * Hp;
    const float* frame = frames + (size_t)n * H * H;
    float part = 0.f;
    for (int i = threadIdx.x; i < Hp * Hp; i += blockDim.x) {
      const int m = i / Hp, nn = i - m * Hp;
      float v = 0.f;
      if (m < H && nn < H) v = __ldg(frame + (size_t)m * H + nn);
      Xs[i] = v;
      part += v;
    }
    const float mean = block_sum(part, red) / (float)(H * H);
    for (int i = threadIdx.x; i < Hp * Hp; i += blockDim.x) {
      const int m = i / Hp, nn = i - m * Hp;
      if (m < H && nn < H) Xs[i] -= mean;
    }
    __syncthreads();
    umma_product<false>(cx, Xs, Hp, nullptr, 0, Hp, P.dct_t, Kp, Kp, Hp, R1, Kp, 0, Kp);       // R1[n][k] = sum_m X[m][n] dct[m][k]
    __syncthreads();
    umma_product<false>(cx, P.dct_t, Kp, nullptr, 0, Kp, R1, Kp, Kp, Hp, Ct, Kp, 0, Kp);       // Ct[l][k] = sum_n dct[n][l] R1[n][k]
    __syncthreads();
    for (int li = 0; li < P.n_levels; ++li) {
      const LevelDev& L = P.lv[li];
      float2* out = reinterpret_cast<float2*>(outs.p[li]);
      const int c = L.c, hp = L.hp, cp = L.cp;
      const int bj = umma_bj(cp);
      // One band and one block column [j0, j0 + bj) of its output at a time: the four inner products fill that column range of
      // U (both channels, both halves), the outer products consume it at once.  The live part of the per-CTA scratch is then
      // Ct + half a band of U (~600 KB at 224x224): with a whole band (1 MB x 148 CTAs) the scratch did not fit L2 and the first
      // version of this kernel moved 88 GB through DRAM per 3328 frames, every chunk waiting on a DRAM-latency load.
      for (int b = 0; b < P.nb; ++b)
        for (int j0 = 0; j0 < cp; j0 += bj) {
          const int je = min(j0 + bj, cp);
          for (int job = 0; job < 4; ++job) {                  // U[ch][half*hp + k][x], x in [j0, je)
            const int ch = job >> 1, half = job & 1;
            const float* mask = L.masks + ((size_t)((b * 2 + ch) * 2 + half) * hp) * hp;
            const float* tab = L.trig + (size_t)L.inner_sel[ch][half] * hp * cp;
            umma_product<true>(cx, Ct, Kp, mask, hp, hp, tab, cp, cp, hp, W + ((size_t)ch * 2 * hp + half * hp) * cp, cp, j0, je);
          }
          __syncthreads();
          const float* Bre = W;                                // out_ch[y][x] = sum_kk trig[kk][y] * U_ch[kk][x], both channels per block
          const float* Bim = Bre + (size_t)2 * hp * cp;
          for (int i0 = 0; i0 < cp; i0 += 128) {
            umma_block<false, true>(cx, L.trig, cp, nullptr, 0, cp, Bre, Bim, cp, cp, 2 * hp, i0, j0, bj);
            if (warp >= 16) continue;                          // loader warps run ahead into the next block
            const int y = i0 + 32 * (warp & 3) + lane;
            const uint32_t taddr = cx.tmem + ((uint32_t)(32 * (warp & 3)) << 16);
            for (int t0 = 32 * (warp >> 2); t0 < bj && t0 < 32 * (warp >> 2) + 32; t0 += 16) {
              float re[16], im[16];
              tcg::tmem_ld16(taddr + t0, re);
              tcg::tmem_ld16(taddr + 128 + t0, im);
              if (y < c) {
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                  const int x = j0 + t0 + g;
                  if (x >= c) continue;
                  float2 v = make_float2(re[g], im[g]);
                  if (polar) {   // (phase, magnitude) exactly as the phase tail computes them from (re, im), once per distinct frame
                    const float ph = atan2f(v.y, v.x);
                    const float mg = __fadd_rn(sqrtf(__fadd_rn(__fmul_rn(v.y, v.y), __fmul_rn(v.x, v.x))), 1e-10f);
                    v = make_float2(ph, mg);
                  }
                  if (!(cx.debug & 4)) out[((((size_t)w * P.nb + b) * T + t) * c + y) * (size_t)c + x] = v;
                }
              }
            }
            tcg::fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tcg::smem_u32(cx.tmem_free)) : "memory");
          }
          __syncthreads();
        }
    }
  }
  tcg::fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tcg::fence_after();
    tcg::tmem_dealloc(cx.tmem, 256);
  }
}

}  // namespace mimamo

using namespace mimamo;

struct mimamo_pyr_plan {
  PlanDev d;
  size_t smem_bytes;
  int large_grid;       // persistent grid size in large-frame mode
  float* dev_blob;      // one allocation holding every table
  int device = 0;       // the CUDA device the tables live on
};

extern "C" int mimamo_pyr_plan_create(int32_t H, int32_t Hp, int32_t Kp, int32_t nbands,
                                      const float* dct_t_host, int32_t n_levels,
                                      const mimamo_pyr_level_desc* levels, mimamo_pyr_plan** plan_out) {
  MM_REQUIRE(plan_out && dct_t_host && levels, MIMAMO_E_VALUE, "null argument");
  MM_REQUIRE(n_levels >= 1 && n_levels <= MIMAMO_MAX_LEVELS, MIMAMO_E_VALUE, "n_levels must be in [1,%d]", MIMAMO_MAX_LEVELS);
  MM_REQUIRE(H >= 1 && Hp % 8 == 0 && Kp % 8 == 0 && Hp >= H && nbands >= 2, MIMAMO_E_VALUE, "bad plan geometry (leading dimensions must be multiples of 8)");
  size_t total = (size_t)Hp * Kp;
  for (int i = 0; i < n_levels; ++i) {
    const mimamo_pyr_level_desc& L = levels[i];
    MM_REQUIRE(L.hp % 8 == 0 && L.cp % 8 == 0 && L.hp >= L.h && L.cp >= L.c && L.hp <= Kp, MIMAMO_E_VALUE, "bad level %d geometry", i);
    total += (size_t)2 * L.hp * L.cp + (size_t)nbands * 4 * L.hp * L.hp;
  }
  mimamo_pyr_plan* plan = new mimamo_pyr_plan();
  PlanDev& d = plan->d;
  d.H = H; d.Hp = Hp; d.Kp = Kp; d.nb = nbands; d.n_levels = n_levels;
  if (cudaMalloc((void**)&plan->dev_blob, total * sizeof(float)) != cudaSuccess) {
    set_error("cudaMalloc of %zu table bytes failed", total * sizeof(float));
    delete plan;
    return MIMAMO_E_CUDA;
  }
  float* cur = plan->dev_blob;
  auto put = [&](const float* host, size_t count) -> const float* {
    cudaMemcpy(cur, host, count * sizeof(float), cudaMemcpyHostToDevice);
    const float* at = cur;
    cur += count;
    return at;
  };
  d.dct_t = put(dct_t_host, (size_t)Hp * Kp);
  // shared-memory budget: Ct + a work region that must hold the forward scratch (X + R1) and at
  // least one band (both channels, both halves) of every level; larger regions batch more bands per barrier.
  int dev = 0, max_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t ct_floats = (size_t)Kp * Kp;
  size_t need = (size_t)Hp * Hp + (size_t)Hp * Kp;
  size_t want = need;
  for (int i = 0; i < n_levels; ++i) {
    const size_t unit = (size_t)4 * levels[i].hp * levels[i].cp;
    need = need > unit ? need : unit;
    const size_t all = unit * nbands;
    want = want > all ? want : all;
  }
  const size_t budget_floats = ((size_t)max_optin - 1024) / sizeof(float);
  size_t work;
  d.large = (ct_floats + need > budget_floats) ? 1 : 0;
  { const char* e = getenv("MIMAMO_PYR_TC"); d.use_tc = e ? atoi(e) : 1; }
  if (d.large) {
    work = need;                       // one band at a time, buffers in global scratch
  } else {
    // Work-region size: big enough to batch every band of a level between barriers if
    // that still leaves room for two CTAs per SM; otherwise as large as one CTA may have.
    const size_t half_budget = budget_floats / 2;
    if (ct_floats + want <= half_budget) work = want;
    else if (ct_floats + need <= half_budget) work = half_budget - ct_floats;
    else if (ct_floats + want <= budget_floats) work = want;
    else work = budget_floats - ct_floats;
    work &= ~(size_t)3;
  }
  d.work_floats = (int)work;
  for (int i = 0; i < n_levels; ++i) {
    const mimamo_pyr_level_desc& L = levels[i];
    LevelDev& o = d.lv[i];
    o.c = L.c; o.h = L.h; o.hp = L.hp; o.cp = L.cp;
    o.trig = put(L.trig_host, (size_t)2 * L.hp * L.cp);
    o.masks = put(L.masks_host, (size_t)nbands * 4 * L.hp * L.hp);
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) o.inner_sel[a][b] = L.inner_sel_host[a * 2 + b];
    const size_t unit = (size_t)4 * L.hp * L.cp;
    int fit = (int)(work / unit);
    if (fit > nbands) fit = nbands;
    o.units_per_chunk = fit < 1 ? 1 : fit;
  }
  plan->smem_bytes = d.large ? 0 : (ct_floats + work) * sizeof(float);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  plan->large_grid = 2 * sms;
  static std::atomic<size_t> max_smem_set[64];     // per device; several plans may coexist: only ever raise the limit
  cudaError_t e = cudaSuccess;
  plan->device = dev;
  if (plan->smem_bytes > max_smem_set[dev & 63].load()) {
    e = cudaFuncSetAttribute(pyr_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->smem_bytes);
    if (e == cudaSuccess) max_smem_set[dev & 63].store(plan->smem_bytes);
  }
  if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    set_error("pyramid plan setup failed: %s", cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
    cudaFree(plan->dev_blob);
    delete plan;
    return MIMAMO_E_CUDA;
  }
  *plan_out = plan;
  return MIMAMO_OK;
}

extern "C" void mimamo_pyr_plan_destroy(mimamo_pyr_plan* plan) {
  if (!plan) return;
  cudaFree(plan->dev_blob);
  delete plan;
}

static size_t pyr_scratch_bytes(const mimamo_pyr_plan* plan) {
  if (!plan->d.large) return 0;
  return (size_t)plan->large_grid * ((size_t)plan->d.Kp * plan->d.Kp + plan->d.work_floats) * sizeof(float);
}

int pyr_build_launch(const mimamo_pyr_plan* plan, const float* frames, int64_t n_windows, int32_t T,
                     float* const* coeff_out, const int* root, void* workspace, size_t workspace_bytes,
                     cudaStream_t stream, int polar) {
  MM_REQUIRE(plan && frames && coeff_out, MIMAMO_E_VALUE, "null argument");
  MM_REQUIRE(n_windows >= 0 && T >= 1, MIMAMO_E_VALUE, "bad batch geometry");
  MM_CHECK_DEVICE(plan->device);
  if (n_windows == 0) return MIMAMO_OK;
  MM_REQUIRE(n_windows * T < (1ll << 31), MIMAMO_E_VALUE, "too many frames for one launch");
  OutPtrs outs;
  for (int i = 0; i < plan->d.n_levels; ++i) {
    MM_REQUIRE(coeff_out[i], MIMAMO_E_VALUE, "null output for level %d", i);
    outs.p[i] = coeff_out[i];
  }
  const long long n_frames = n_windows * T;
  if (plan->d.large) {
    const size_t need = pyr_scratch_bytes(plan);
    MM_REQUIRE(workspace && workspace_bytes >= need, MIMAMO_E_VALUE, "workspace too small: need %zu bytes", need);
    if (plan->d.use_tc) {
      static DeviceOnce attr_set;
      if (attr_set.need()) {
        MM_CUDA(cudaFuncSetAttribute(pyr_build_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kUmmaSmemBytes));
        attr_set.mark();
      }
      const int sms = plan->large_grid / 2;                    // one CTA per SM (192 KB of staging planes)
      const unsigned grid = (unsigned)(n_frames < sms ? n_frames : sms);
      pyr_build_umma_kernel<<<grid, kUmmaThreads, kUmmaSmemBytes, stream>>>(plan->d, frames, T, outs, root, (float*)workspace, n_frames, polar);
    } else {
      const unsigned grid = (unsigned)(n_frames < plan->large_grid ? n_frames : plan->large_grid);
      pyr_build_kernel<true><<<grid, kPyrThreads, 0, stream>>>(plan->d, frames, T, outs, root, (float*)workspace, n_frames, polar);
    }
  } else {
    pyr_build_kernel<false><<<(unsigned)n_frames, kPyrThreads, plan->smem_bytes, stream>>>(plan->d, frames, T, outs, root, nullptr, n_frames, polar);
  }
  MM_LAUNCH_OK();
  return MIMAMO_OK;
}

extern "C" int mimamo_pyr_build_workspace_bytes(const mimamo_pyr_plan* plan, int64_t n_windows, int32_t T, size_t* bytes_out) {
  MM_REQUIRE(plan && bytes_out, MIMAMO_E_VALUE, "null argument");
  *bytes_out = pyr_scratch_bytes(plan);
  return MIMAMO_OK;
}

extern "C" int mimamo_pyr_build(const mimamo_pyr_plan* plan, const float* frames, int64_t n_windows,
                                int32_t T, float* const* coeff_out, void* workspace, size_t workspace_bytes, void* stream) {
  return pyr_build_launch(plan, frames, n_windows, T, coeff_out, nullptr, workspace, workspace_bytes, (cudaStream_t)stream, 0);
}

// accessors used by the fused path in phase_tail.cu
extern "C" int mimamo_pyr_plan_levels(const mimamo_pyr_plan* plan, int32_t* n_levels, int32_t* nbands, int32_t* crops) {
  if (crops == nullptr) { *n_levels = plan->d.H; return MIMAMO_OK; }     // frame size query
  *n_levels = plan->d.n_levels;
  *nbands = plan->d.nb;
  for (int i = 0; i < plan->d.n_levels; ++i) crops[i] = plan->d.lv[i].c;
  return MIMAMO_OK;
}
