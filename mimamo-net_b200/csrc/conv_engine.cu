// tcgen05 implicit-GEMM convolution engine for sm_100a.
//
// Persistent, warp-specialised kernels serve every convolution on the hot path (ResNet50's 1x1 / 3x3 /
// strided layers, api/resnet50_extractor.py:81, and PhaseNet's 3x3 layers, api/mimamo_net.py:68-90):
//
//   conv_gemm_kernel     1x1 layers (flat rows) and box-per-tap 3x3 / strided layers; optional "pair" mode: 2-CTA
//                        clusters whose weight boxes are fetched half each and TMA-multicast into both CTAs
//   conv_gemm2_kernel    the same GEMM with tcgen05.mma.cta_group::2: a CTA pair computes a 256 x 256 tile, each CTA
//                        staging its 128 activation rows and half of the weight rows (K-heavy 256-wide layers)
//   conv3x3_halo_kernel  stride-1 3x3 with Cout <= 128: the input patch is loaded once, nine shifted UMMA views
//   conv1_line_kernel    conv1_7x7_s2 over the space-to-depth'ed input, pool1_3x3_s2 fused into the epilogue
//
//   warp 0      TMA producer: activation + weight boxes into a ring of 128B-swizzled shared-memory stages
//               (mbarrier full/empty pipeline).  A 3x3 tap is just a shifted 4-D box over the NHWC activation
//               tensor; TMA's out-of-bounds zero fill IS the convolution padding, and its element strides
//               implement stride-2 layers, so no im2col buffer exists for any layer.
//   warp 1      single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16, kind::f16, fp32 accumulators in TMEM,
//               double buffered so tile i+1 overlaps tile i's epilogue).
//   warps 2-17  epilogue: tcgen05.ld TMEM -> registers (thread = output row), folded-BatchNorm scale/shift,
//               optional residual add + ReLU, 16-bit pack, swizzled staging tile, TMA bulk-tensor store.
//
// D[M = output pixels][N = Cout] = A[M][K] * B[N][K]^T,  K = taps * Cin_p (K-major, 16-bit).
#include "common.cuh"
#include "conv_engine.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <math.h>
#include <thread>
#include <vector>

namespace mimamo {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                   // 64 x 2 B = one 128-byte swizzle atom
constexpr int kAStageBytes = kBlockM * kBlockK * 2;
constexpr int kEpiWarps = 16;                 // four per TMEM lane quarter, each owning a quarter of the tile's columns
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;   // 1 TMA warp + 1 MMA warp + epilogue warps
constexpr int kResDepth = 2;                  // residual prefetch ring: chunks (32 columns) in flight per warp
constexpr int kUmmaK = 16;                  // elements per tcgen05.mma K step (umma_kblock issues kBlockK / kUmmaK = 4 of them)
static_assert(kBlockK / kUmmaK == 4, "umma_kblock is written for four K steps");

struct ConvParams {
  int mode;                 // 0: flat rows, 1: spatial boxes
  int M_total;              // flat: output pixels
  int num_k_blocks, cin_blocks, taps_w;
  int Wo, Ho, Nimg;
  int bw, bh, bn, tiles_w, tiles_h;
  int stride, pad;
  int a_rows;               // rows one activation box delivers (<= 128)
  int m_tiles, n_tiles;
  uint32_t idesc;
  const float* scale;
  const float* shift;
  const void* residual;
  const void* a_ptr;         // conv1 line kernel: the space-to-depth'ed input (plain bulk copies, no tensor map)
  void* out;
  int ldc, ld_res;
  int relu;
  int debug;                 // timing experiments only (MIMAMO_DEBUG): 1 = epilogue skips math+store (layers without residual), 2 = line kernel issues 4 of 16 UMMAs, 4 / 8 = halo kernel (see there)
  int pair;                  // conv_gemm_kernel launched as 2-CTA clusters: the two CTAs work on adjacent M tiles of the same N tile and
                             // each fetches half of every weight box, multicast into both shared memories (halves the L2->SM weight traffic)
  int store_mode;            // epilogue TMA store granularity: 0 = per warp (32 rows, flat layers), 1 = per column group, 2 = whole tile
  int resident_w;            // halo kernel: the whole 3x3 weight set stays in shared memory (Cin_p == 64, 9 taps <= kBStages boxes)
  int res_tma;               // residual fetched by TMA straight into the output staging tile (flat 256-wide layers, per-warp stores)
  alignas(64) CUtensorMap res_map;   // ... through this map: the residual tensor with the geometry of the output map
  int sub, sub_pad;          // strided k x k layers: tap (kh, kw) reads a DENSE box of one (row parity, column parity) sub-lattice of
                             // the input through sub_map[2 * row parity + column parity] (conv_forward); sub_pad = the layer's padding
  alignas(64) CUtensorMap sub_map[4];
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a pipeline bug (wrong expect_tx byte count, bad descriptor) traps instead of
// hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 27)) __trap();
  }
}
// one lane of a converged warp (warp-uniform control flow keeps loop state in uniform registers,
// which is what UTMALDG / UTCHMMA take their operands from)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; single thread issues on behalf of the CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (= 1, unused for swizzled K-major),
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B -> 64), [46,48) version = 1 (sm_100),
//   [61,64) layout type = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// ---- lean issue-path helpers: 32-bit shared addresses, no address conversion in the loops ----
__device__ __forceinline__ void bar_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0, ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 27)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void bar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma2d_u32(uint32_t dst, uint64_t map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma3d_u32(uint32_t dst, uint64_t map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma4d_u32(uint32_t dst, uint64_t map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void commit_u32(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Four K=16 UMMAs of one 64-wide K block.  Descriptors are built from their 32-bit halves inside
// PTX: lo = (address >> 4) field + constant LBO, hi = SBO/version/swizzle constant; +2 per K step.
constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);    // bits [32,64): SBO = 64, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_kblock(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      ".reg .b32 al, bl;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
      "add.u32 al, %1, 2;\n add.u32 bl, %2, 2;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n"
      "add.u32 al, %1, 4;\n add.u32 bl, %2, 4;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n"
      "add.u32 al, %1, 6;\n add.u32 bl, %2, 6;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  if (BF16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

// RES: 0 = no residual, 1 = residual through the per-lane cp.async ring (or, p.res_tma, by TMA with the ring's memory as
// the second staging buffer), 2 = residual by TMA into a SINGLE staging buffer, added in place -- no ring, no second
// buffer, and the 64 KB they would take go to the operand pipeline (K >= 256 layers: 2 -> 3 stages, cta_group::2: 3 -> 5)
template <int BLOCK_N, int RES>
struct GemmCfg {
  static constexpr int kBStageBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kEpiBufs = BLOCK_N <= 128 ? 2 : 1;                          // double-buffered staging where shared memory allows
  static constexpr int kEpiBytes = kEpiBufs * kBlockM * BLOCK_N * 2;             // 16-bit [128][BLOCK_N] TMA-store staging
  static constexpr int kResBytes = RES == 1 ? kEpiWarps * kResDepth * 2048 : 0;   // cp.async residual ring
  static constexpr int kBudget = 232448 - 1024 - 512 - kEpiBytes - kResBytes;
  static constexpr int kMaxStages = kBudget / (kAStageBytes + kBStageBytes);
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kTmemCols = 2 * BLOCK_N;          // double-buffered accumulator (128/256/512)
  static constexpr int kSmemBytes = kStages * (kAStageBytes + kBStageBytes) + kEpiBytes + kResBytes + 512 + 1024;
  static_assert(kStages >= 2, "pipeline needs at least two stages");
};

// output pixel of accumulator row `row` of M tile `m_tile` (-1: padding row, nothing to store)
__device__ __forceinline__ int row_pixel(const ConvParams& p, int m_tile, int row) {
  if (p.mode == 0) {
    const long long q = (long long)m_tile * kBlockM + row;
    return q < p.M_total ? (int)q : -1;
  }
  if (p.mode == 2) {                 // halo mode: rows run over the (W+2)-wide padded line, p.bw = W + 2
    const int n = m_tile / p.tiles_h, th = m_tile - n * p.tiles_h;
    const int dh = row / p.bw, dw = row - dh * p.bw;
    const int h = th * p.bh + dh;
    return (dh < p.bh && dw < p.Wo && h < p.Ho) ? (n * p.Ho + h) * p.Wo + dw : -1;
  }
  const int tw = m_tile % p.tiles_w, rest = m_tile / p.tiles_w;
  const int th = rest % p.tiles_h, tn = rest / p.tiles_h;
  const int per_img = p.bw * p.bh;
  const int dn = row / per_img, rem = row - dn * per_img;
  const int dh = rem / p.bw, dw = rem - dh * p.bw;
  const int n = tn * p.bn + dn, h = th * p.bh + dh, w = tw * p.bw + dw;
  return (row < p.a_rows && n < p.Nimg && h < p.Ho && w < p.Wo) ? (n * p.Ho + h) * p.Wo + w : -1;
}

// ---- TMA store / async-proxy helpers ----
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void tma_store_2d(uint64_t map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(uint64_t map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// ---- 2-CTA cluster helpers (pair mode of conv_gemm_kernel) ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2d_multicast_u32(uint32_t dst, uint64_t map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void commit_multicast_u32(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// ---- cta_group::2 (2-SM UMMA) helpers ----
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {       // same offset in the peer's shared memory
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster_u32(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// TMA loads of a CTA pair: data lands in the issuing CTA, the bytes are credited to the LEADER's mbarrier
__device__ __forceinline__ void tma2d_2sm_u32(uint32_t dst, uint64_t map, uint32_t leader_bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma4d_2sm_u32(uint32_t dst, uint64_t map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void commit2_multicast_u32(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// Four K=16 UMMAs of one 64-wide K block issued for BOTH CTAs of the pair (M = 256: 128 rows of A from each CTA's shared
// memory, N = 256: 128 rows of B from each), accumulating into the same TMEM columns of both CTAs.
__device__ __forceinline__ void umma2_kblock(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      ".reg .b32 al, bl;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n"
      "add.u32 al, %1, 2;\n add.u32 bl, %2, 2;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n"
      "add.u32 al, %1, 4;\n add.u32 bl, %2, 4;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n"
      "add.u32 al, %1, 6;\n add.u32 bl, %2, 6;\n mov.b64 da, {al, %5};\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Epilogue warps (shared by all three kernels).  ncu on the first versions (profiles/gemm_r1_ncu_summary_v2.txt)
// showed the 1x1 layers bound by the epilogue's shared-memory instruction queue (mio_throttle): TMEM -> fp32
// staging -> coalesced re-read -> 16-byte global stores moved every output through the LSU three times.
// Now each thread keeps the accumulator row tcgen05.ld hands it (thread = output row, 32 columns per chunk),
// applies folded BatchNorm / residual / ReLU in registers, packs to 16 bits and writes its 64 bytes ONCE into a
// swizzled [128 rows][COLS] staging tile; the four warps of a column group (one per TMEM lane quarter) then hand
// the whole tile to the TMA engine (cp.async.bulk.tensor store), which clips rows/columns outside the tensor --
// so padded accumulator rows (partial tiles, the two extra columns of a halo line) need no predicate at all.
// The residual is prefetched by per-lane cp.async into a ring kResDepth chunks deep that runs ahead ACROSS tiles,
// written in the same swizzled row layout so that thread = row reads it back conflict-free.  (Pulling the residual
// of later tiles into L2 first -- TMA prefetch or prefetch.global.L2 -- was measured 10-15 % SLOWER: these layers
// are bound by bytes through L2, and a prefetch moves every residual byte through it twice.)
template <int BLOCK_N, bool BF16, int RES, int EPI_BUFS>
__device__ __forceinline__ void epilogue_warps(const ConvParams& p, const CUtensorMap* tmOut, int warp, int lane, uint8_t* sEpi,
                                               uint8_t* sRes, uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base,
                                               int num_tiles, int tile0, int tile_stride, int m_shift = 0, int m_rank = 0,
                                               uint32_t tmem_empty_leader = 0 /* cta_group::2: shared::cluster address of the leader's tmem_empty[0] */,
                                               uint64_t* res_bar = nullptr /* [16 warps][2] mbarriers of the TMA residual path */) {
    constexpr bool HAS_RES = RES != 0;
    constexpr bool RES_SINGLE = RES == 2;                     // one staging buffer, residual added in place
    const int ew = warp - 2;
    const int quarter = warp & 3;                             // TMEM lanes [32q, 32q+32) belong to warp%4 == q
    constexpr int PARTS = BLOCK_N / 32 < 4 ? BLOCK_N / 32 : 4;   // column groups per tile (64-wide tiles: 2)
    const int half = ew >> 2;                                 // which column group this warp owns
    constexpr int COLS = BLOCK_N / PARTS;                     // columns per column group: 32 or 64
    constexpr int CPW = COLS / 32;                            // 32-column chunks per warp per tile
    constexpr int ROW_BYTES = COLS * 2;                       // staging row: 64 B (SWIZZLE_64B) or 128 B (SWIZZLE_128B)
    if (half < PARTS) {
    const int row = quarter * 32 + lane;                      // accumulator row of this thread
    constexpr int BUF_BYTES = kBlockM * BLOCK_N * 2;         // one staging tile (all column groups)
    const uint32_t stage0_u32 = smem_u32(sEpi + half * (128 * ROW_BYTES));
    const uint32_t swz = ROW_BYTES == 128 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);   // XOR on the 16-byte chunk index
    const uint32_t res_u32 = smem_u32(sRes + ew * (kResDepth * 2048));
    const int sub_row = lane >> 2, pair = lane & 3;           // residual fetch: row (it*8 + sub_row), 16-byte piece `pair`
    const uint16_t* res_base = reinterpret_cast<const uint16_t*>(p.residual);
    // Who hands finished rows to the TMA engine (measured per layer type, profiles/launches_r1_*.csv): the epilogue-
    // bound flat 1x1 layers want the warps fully decoupled (each warp stores its own 32 rows); the tensor-bound 3x3
    // layers run faster when the whole tile goes out behind one barrier; strided 1x1 boxes go per column group.
    const int store_mode = p.store_mode;
    const bool issuer = store_mode == 0 ? lane == 0 : (store_mode == 1 ? (quarter == 0 && lane == 0) : (ew == 2 && lane == 0));
    const int bar_id = 1 + half;
    const uint64_t map_out = reinterpret_cast<uint64_t>(tmOut);
    auto sync_store_group = [&]() {
      if (store_mode == 0) __syncwarp();
      else if (store_mode == 1) named_bar_sync(bar_id, 128);
      else named_bar_sync(1, PARTS * 128);
    };

    auto issue_residual = [&](int q) {                        // chunk q of this warp's flattened (tile, chunk) list
      if (HAS_RES) {
        const int tile = tile0 + (q / CPW) * tile_stride;
        if (tile < num_tiles) {
          const int m_lin = tile / p.n_tiles, n_tile = tile - m_lin * p.n_tiles;
          const int m_tile = (m_lin << m_shift) + m_rank;
          const int ch = n_tile * BLOCK_N + half * COLS + (q % CPW) * 32 + pair * 8;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + sub_row;
            const int rp = row_pixel(p, m_tile, quarter * 32 + r);
            if (rp >= 0) {
              const uint32_t dst = res_u32 + (q % kResDepth) * 2048 + r * 64 + ((pair ^ ((r >> 1) & 3)) << 4);
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(res_base + (size_t)rp * p.ld_res + ch) : "memory");
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    };
    // Residual by TMA (p.res_tma: flat 256-wide layers with per-warp stores): a warp's 32 x 64 slice of the residual lands,
    // through a box of the same geometry as its output box, in the very staging slice the warp will store from; each
    // thread adds its accumulator row to its residual row IN PLACE.  The staging tile and the former cp.async ring form
    // two such buffers, so the residual of tile i+1 is in flight while tile i is processed.  This removes the per-lane
    // cp.async address arithmetic from the epilogue (ncu, round 1: these layers issue at 66 % with DRAM at 65-70 %).
    const bool res_tma = RES_SINGLE || (HAS_RES && p.res_tma != 0);
    const uint64_t map_res = reinterpret_cast<uint64_t>(&p.res_map);
    auto issue_res_tma = [&](int tile, int buf) {              // lane 0 only
      const int m_lin = tile / p.n_tiles, n_tile = tile - m_lin * p.n_tiles;
      const int m_tile = (m_lin << m_shift) + m_rank;
      const uint32_t bar = smem_u32(&res_bar[ew * 2 + buf]);
      bar_expect_tx_u32(bar, 32u * ROW_BYTES);
      tma2d_u32(stage0_u32 + (uint32_t)buf * BUF_BYTES + quarter * (32 * ROW_BYTES), map_res, bar,
                n_tile * BLOCK_N + half * COLS, m_tile * kBlockM + quarter * 32);
    };
    int qc = 0;                                               // chunks consumed so far
    if (res_tma) { if (lane == 0 && tile0 < num_tiles) issue_res_tma(tile0, 0); }
    else for (int q = 0; q < kResDepth; ++q) issue_residual(q);

    int local = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_stride, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int m_lin = tile / p.n_tiles, n_tile = tile - m_lin * p.n_tiles;
      const int m_tile = (m_lin << m_shift) + m_rank;
      const int n0 = n_tile * BLOCK_N + half * COLS;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      // staging tiles alternate (EPI_BUFS == 2): this one is free once the store issued two tiles ago has read it
      const uint32_t buf_off = (EPI_BUFS == 2 || (res_tma && !RES_SINGLE)) ? (uint32_t)(local & 1) * BUF_BYTES : 0u;
      const uint32_t stage_u32 = stage0_u32 + buf_off;
      const uint32_t my_row_u32 = stage_u32 + row * ROW_BYTES;
      if (RES_SINGLE) {
        mbar_wait(&res_bar[ew * 2], local & 1);                  // this tile's residual slice has landed (it was requested
                                                                 // once the previous tile's store had read the buffer)
      } else if (res_tma) {
        if (lane == 0) {
          tma_store_wait_read();                                 // every store of this warp has read its staging slice
          if (tile + tile_stride < num_tiles) issue_res_tma(tile + tile_stride, (local & 1) ^ 1);
        }
        __syncwarp();
        mbar_wait(&res_bar[ew * 2 + (local & 1)], (local >> 1) & 1);
      } else {
        if (issuer) { if (EPI_BUFS == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
        sync_store_group();
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + half * COLS;
      if ((p.debug & 1) && !HAS_RES && !tmem_empty_leader) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&tmem_empty[acc]); continue; }
#pragma unroll 1
      for (int c0 = 0; c0 < COLS; c0 += 32, ++qc) {
        uint32_t v[32];
        tmem_ld16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_ld16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        uint32_t rw[16];
        if (HAS_RES && res_tma) {                              // this thread's residual row sits where its output row will go
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rw[4 * j]), "=r"(rw[4 * j + 1]), "=r"(rw[4 * j + 2]), "=r"(rw[4 * j + 3])
                         : "r"(my_row_u32 + (((uint32_t)(c0 / 8 + j) ^ swz) << 4)));
        } else if (HAS_RES) {
          asm volatile("cp.async.wait_group %0;" ::"n"(kResDepth - 1) : "memory");
          __syncwarp();                                        // every lane's pieces of the chunk have landed
          const uint32_t src = res_u32 + (qc % kResDepth) * 2048 + lane * 64;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rw[4 * j]), "=r"(rw[4 * j + 1]), "=r"(rw[4 * j + 2]), "=r"(rw[4 * j + 3])
                         : "r"(src + ((j ^ ((lane >> 1) & 3)) << 4)));
        }
        tmem_ld_wait();
        const float4* sc = reinterpret_cast<const float4*>(p.scale + n0 + c0);      // warp-uniform addresses: broadcast loads
        const float4* sh = reinterpret_cast<const float4*>(p.shift + n0 + c0);
#pragma unroll
        for (int g = 0; g < 4; ++g) {                          // 8 columns = one 16-byte chunk of the 16-bit row
          const float4 s0 = __ldg(sc + 2 * g), s1 = __ldg(sc + 2 * g + 1), t0 = __ldg(sh + 2 * g), t1 = __ldg(sh + 2 * g + 1);
          float o[8] = {__uint_as_float(v[8 * g]) * s0.x + t0.x, __uint_as_float(v[8 * g + 1]) * s0.y + t0.y,
                        __uint_as_float(v[8 * g + 2]) * s0.z + t0.z, __uint_as_float(v[8 * g + 3]) * s0.w + t0.w,
                        __uint_as_float(v[8 * g + 4]) * s1.x + t1.x, __uint_as_float(v[8 * g + 5]) * s1.y + t1.y,
                        __uint_as_float(v[8 * g + 6]) * s1.z + t1.z, __uint_as_float(v[8 * g + 7]) * s1.w + t1.w};
          if (HAS_RES) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack2<BF16>(rw[4 * g + j]);
              o[2 * j] += f.x; o[2 * j + 1] += f.y;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
          }
          const uint32_t chunk = (uint32_t)(c0 / 8 + g);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row_u32 + ((chunk ^ swz) << 4)), "r"(pack2<BF16>(o[0], o[1])),
                       "r"(pack2<BF16>(o[2], o[3])), "r"(pack2<BF16>(o[4], o[5])), "r"(pack2<BF16>(o[6], o[7])) : "memory");
        }
        if (HAS_RES && !res_tma) {
          __syncwarp();                                        // all lanes have read the ring slot
          issue_residual(qc + kResDepth);                      // refill it
        }
      }
      // accumulator fully read: hand the TMEM buffer back before the store is even issued
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (tmem_empty_leader) mbar_arrive_cluster_u32(tmem_empty_leader + acc * 8);   // the pair's MMAs are issued by the leader CTA
        else mbar_arrive(&tmem_empty[acc]);
      }
      fence_proxy_async_smem();                                // generic-proxy writes -> visible to the TMA engine
      sync_store_group();
      if (issuer) {
        if (store_mode == 0) {                                 // flat layer: this warp's 32 rows
          tma_store_2d(map_out, stage_u32 + quarter * (32 * ROW_BYTES), n0, m_tile * kBlockM + quarter * 32);
        } else {
          const int g0 = store_mode == 1 ? half : 0, g1 = store_mode == 1 ? half + 1 : PARTS;
          for (int g = g0; g < g1; ++g) {
            const uint32_t src = smem_u32(sEpi) + buf_off + g * (128 * ROW_BYTES);
            const int nc = n_tile * BLOCK_N + g * COLS;
            if (p.mode == 0) {
              tma_store_2d(map_out, src, nc, m_tile * kBlockM);
            } else if (p.mode == 2) {
              const int n_img = m_tile / p.tiles_h, th = m_tile - n_img * p.tiles_h;
              tma_store_4d(map_out, src, nc, 0, th * p.bh, n_img);
            } else {
              const int tw = m_tile % p.tiles_w, rest = m_tile / p.tiles_w;
              const int th = rest % p.tiles_h, tn = rest / p.tiles_h;
              tma_store_4d(map_out, src, nc, tw * p.bw, th * p.bh, tn * p.bn);
            }
          }
        }
        tma_store_commit();
        if (RES_SINGLE) {                                        // single buffer: the next residual may land once this store has read it
          tma_store_wait_read();
          if (tile + tile_stride < num_tiles) issue_res_tma(tile + tile_stride, 0);
        }
      }
    }
    if (HAS_RES && !res_tma) asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (issuer) tma_store_wait_all();
    }
}

template <int BLOCK_N, bool BF16, int RES>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ConvParams p) {
  using Cfg = GemmCfg<BLOCK_N, RES>;
  constexpr bool HAS_RES = RES != 0;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kAStageBytes;
  uint8_t* sEpi = sB + STAGES * Cfg::kBStageBytes;
  uint8_t* sRes = sEpi + Cfg::kEpiBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sRes + Cfg::kResBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);          // [16 warps][2]: TMA residual path

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (p.sub) for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.sub_map[i]);
    if (RES == 2 || (HAS_RES && p.res_tma)) {
      tma_prefetch_desc(&p.res_map);
      for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    }
    // pair mode: a stage is reusable once BOTH CTAs' MMAs have read it (each CTA's producer also writes the peer's stage)
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], p.pair ? 2 : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4 * (BLOCK_N / 32 < 4 ? BLOCK_N / 32 : 4)); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  if (p.pair) cluster_sync_all();                            // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tile list of this CTA: plain = tiles blockIdx.x, +gridDim.x, ...; pair = tiles of (M-tile pair, N tile), this CTA taking
  // M tile 2 * pair + rank (an M tile past the end is computed on zero-filled rows and clipped by the TMA store)
  const uint32_t rank = p.pair ? cluster_ctarank() : 0u;
  const int m_shift = p.pair ? 1 : 0;
  const int num_tiles = p.pair ? ((p.m_tiles + 1) >> 1) * p.n_tiles : p.m_tiles * p.n_tiles;
  const int tile0 = p.pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_stride = p.pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  // The issue warps execute a strictly serial instruction stream: at ~5 cycles per dependent
  // instruction their loop length, not TMA or the tensor pipe, bounded the first versions of this
  // kernel (~600 cycles per K block, ncu source view).  Both loops are therefore kept minimal:
  // election hoisted, addresses advanced incrementally, descriptors built from 32-bit halves.
  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    const bool leader = elect_one();
    const uint32_t tx_bytes = (uint32_t)p.a_rows * (kBlockK * 2) + Cfg::kBStageBytes;
    const int n_tiles = p.n_tiles, mode = p.mode, cin_blocks = p.cin_blocks, taps_w = p.taps_w;
    const bool pair = p.pair != 0;
    const int taps_h = p.num_k_blocks / (cin_blocks * taps_w);
    const int tiles_w = p.tiles_w, tiles_h = p.tiles_h;
    const int step_w = p.bw * p.stride, step_h = p.bh * p.stride, pad = p.pad, bn = p.bn;
    const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const uint64_t mapA = reinterpret_cast<uint64_t>(&tmA), mapB = reinterpret_cast<uint64_t>(&tmB);
    const uint64_t mapS = reinterpret_cast<uint64_t>(&p.sub_map[0]);
    const bool sub = p.sub != 0;                               // stride-2 k x k: per-tap sub-lattice maps (see conv_forward)
    const int sub_pad = p.sub_pad;
    int stage = 0; uint32_t parity = 1;                        // producer waits on empty with parity phase ^ 1
    uint32_t dA = sA0, dB = sB0, fb = full0, eb = empty0;
#pragma unroll 1
    for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
      const int m_lin = tile / n_tiles, n_tile = tile - m_lin * n_tiles;
      const int m_tile = (m_lin << m_shift) + (int)rank;
      const int n0 = n_tile * BLOCK_N;
      int c1 = 0, c2 = 0, c3 = 0;
      if (mode == 0) {
        c1 = m_tile * kBlockM;
      } else {
        const int tw = m_tile % tiles_w, rest = m_tile / tiles_w;
        const int th = rest % tiles_h, tn = rest / tiles_h;
        c1 = tw * step_w - pad;
        c2 = th * step_h - pad;
        c3 = tn * bn;
      }
      int kcol = 0;                                            // K offset into the weight matrix
#pragma unroll 1
      for (int kh = 0; kh < taps_h; ++kh) {
#pragma unroll 1
        for (int kw = 0; kw < taps_w; ++kw) {
#pragma unroll 1
          for (int cb = 0; cb < cin_blocks; ++cb, kcol += kBlockK) {
            bar_wait_u32(eb, parity);
            if (leader) {
              bar_expect_tx_u32(fb, tx_bytes);
              if (mode == 0) tma2d_u32(dA, mapA, fb, cb * kBlockK, c1);
              else if (sub) tma4d_u32(dA, mapS + (uint64_t)((((kh - sub_pad) & 1) << 1) | ((kw - sub_pad) & 1)) * sizeof(CUtensorMap), fb,
                                      cb * kBlockK, c1 + ((kw - sub_pad) >> 1), c2 + ((kh - sub_pad) >> 1), c3);
              else tma4d_u32(dA, mapA, fb, cb * kBlockK, c1 + kw, c2 + kh, c3);
              if (pair) tma2d_multicast_u32(dB + rank * (Cfg::kBStageBytes / 2), mapB, fb, kcol, n0 + (int)rank * (BLOCK_N / 2), (uint16_t)3);
              else tma2d_u32(dB, mapB, fb, kcol, n0);
            }
            if (++stage == STAGES) { stage = 0; parity ^= 1; dA = sA0; dB = sB0; fb = full0; eb = empty0; }
            else { dA += kAStageBytes; dB += Cfg::kBStageBytes; fb += 8; eb += 8; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    const bool leader = elect_one();
    const int nkb = p.num_k_blocks;
    const uint32_t idesc = p.idesc;
    const uint32_t a_lo0 = desc_lo(smem_u32(sA)), b_lo0 = desc_lo(smem_u32(sB));
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const uint32_t tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
    int stage = 0; uint32_t parity = 0;
    uint32_t a_lo = a_lo0, b_lo = b_lo0, fb = full0, eb = empty0;
    int local = 0;
    const bool pair = p.pair != 0;
#pragma unroll 1
    for (int tile = tile0; tile < num_tiles; tile += tile_stride, ++local) {
      const uint32_t acc = local & 1;
      bar_wait_u32(tempty0 + acc * 8, ((local >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        bar_wait_u32(fb, parity);
        tc_fence_after();
        if (leader) {
          umma_kblock(d_tmem, a_lo, b_lo, idesc, kb != 0 ? 1u : 0u);
          if (pair) commit_multicast_u32(eb, (uint16_t)3);     // frees the stage in both CTAs when these MMAs retire
          else commit_u32(eb);                                 // frees the smem stage when the MMAs retire
          if (kb == nkb - 1) commit_u32(tfull0 + acc * 8);
        }
        if (++stage == STAGES) { stage = 0; parity ^= 1; a_lo = a_lo0; b_lo = b_lo0; fb = full0; eb = empty0; }
        else { a_lo += kAStageBytes >> 4; b_lo += Cfg::kBStageBytes >> 4; fb += 8; eb += 8; }
      }
    }
  } else {
    epilogue_warps<BLOCK_N, BF16, RES, Cfg::kEpiBufs>(p, &tmOut, warp, lane, sEpi, sRes, tmem_full, tmem_empty, tmem_base, num_tiles, tile0, tile_stride,
                                                      m_shift, (int)rank, 0u, res_bar);
  }
  tc_fence_before();
  __syncthreads();
  if (p.pair) cluster_sync_all();                            // the peer may still signal this CTA's barriers until it is done too
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------
// conv_gemm2_kernel: the same implicit GEMM on CTA PAIRS with tcgen05.mma.cta_group::2.
//
// K-heavy 256-wide layers (stage 4/5 of ResNet50) pull 48 KB per 512 MMA cycles into every SM (16 KB of activations +
// 32 KB of weights per K block) and stall on that ingress (tensor pipe 50-60 % active, DRAM < 50 %).  With the 2-SM UMMA
// a pair of CTAs on adjacent M tiles computes a 256 x 256 tile: each CTA stages its own 128 activation rows and only
// HALF of the weight rows (16 KB); the tensor cores of both SMs read both halves.  Per-SM ingress drops to 32 KB per
// K block and the freed shared memory deepens the ring (5 stages instead of 3; 3 instead of 2 with a residual).
// Roles: both CTAs run a producer warp (their TMA bytes are credited to the LEADER's full barrier) and the 16
// epilogue warps (each CTA reads its own TMEM lanes); only the leader's MMA warp issues, and its tcgen05.commit is
// multicast to both CTAs' empty / tmem_full barriers; both epilogues arrive at the leader's tmem_empty barrier.
// ---------------------------------------------------------------------------------------
template <int BLOCK_N, int RES>
struct Gemm2Cfg {
  static constexpr int kBStageBytes = (BLOCK_N / 2) * kBlockK * 2;
  static constexpr int kEpiBufs = 1;
  static constexpr int kEpiBytes = kBlockM * BLOCK_N * 2;
  static constexpr int kResBytes = RES == 1 ? kEpiWarps * kResDepth * 2048 : 0;
  static constexpr int kBudget = 232448 - 1024 - 512 - kEpiBytes - kResBytes;
  static constexpr int kMaxStages = kBudget / (kAStageBytes + kBStageBytes);
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes = kStages * (kAStageBytes + kBStageBytes) + kEpiBytes + kResBytes + 512 + 1024;
  static_assert(kStages >= 2, "pipeline needs at least two stages");
};

template <int BLOCK_N, bool BF16, int RES>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ConvParams p) {
  using Cfg = Gemm2Cfg<BLOCK_N, RES>;
  constexpr bool HAS_RES = RES != 0;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kAStageBytes;
  uint8_t* sEpi = sB + STAGES * Cfg::kBStageBytes;
  uint8_t* sRes = sEpi + Cfg::kEpiBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sRes + Cfg::kResBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* res_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);          // [16 warps][2]: TMA residual path
  constexpr int kEpiArrivals = 4 * (BLOCK_N / 32 < 4 ? BLOCK_N / 32 : 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (p.sub) for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.sub_map[i]);
    if (RES == 2 || (HAS_RES && p.res_tma)) {
      tma_prefetch_desc(&p.res_map);
      for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    }
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * kEpiArrivals); }   // both CTAs' epilogue warps
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t rank = cluster_ctarank();
  const int num_tiles = ((p.m_tiles + 1) >> 1) * p.n_tiles;  // (M-tile pair, N tile); this CTA owns M tile 2 * pair + rank
  const int tile0 = (int)(blockIdx.x >> 1), tile_stride = (int)(gridDim.x >> 1);

  if (warp == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------------------
    const bool leader = elect_one();
    const uint32_t pair_tx = 2u * ((uint32_t)p.a_rows * (kBlockK * 2) + Cfg::kBStageBytes);
    const int n_tiles = p.n_tiles, mode = p.mode, cin_blocks = p.cin_blocks, taps_w = p.taps_w;
    const int taps_h = p.num_k_blocks / (cin_blocks * taps_w);
    const int tiles_w = p.tiles_w, tiles_h = p.tiles_h;
    const int step_w = p.bw * p.stride, step_h = p.bh * p.stride, pad = p.pad, bn = p.bn;
    const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const uint32_t lfull0 = mapa_u32(full0, 0);              // the leader's full barriers (own ones when rank == 0)
    const uint64_t mapA = reinterpret_cast<uint64_t>(&tmA), mapB = reinterpret_cast<uint64_t>(&tmB);
    const uint64_t mapS = reinterpret_cast<uint64_t>(&p.sub_map[0]);
    const bool sub = p.sub != 0;
    const int sub_pad = p.sub_pad;
    int stage = 0; uint32_t parity = 1;
    uint32_t dA = sA0, dB = sB0, fb = full0, lfb = lfull0, eb = empty0;
#pragma unroll 1
    for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
      const int m_lin = tile / n_tiles, n_tile = tile - m_lin * n_tiles;
      const int m_tile = (m_lin << 1) + (int)rank;
      const int n0 = n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2);     // this CTA's half of the weight rows
      int c1 = 0, c2 = 0, c3 = 0;
      if (mode == 0) {
        c1 = m_tile * kBlockM;
      } else {
        const int tw = m_tile % tiles_w, rest = m_tile / tiles_w;
        const int th = rest % tiles_h, tn = rest / tiles_h;
        c1 = tw * step_w - pad;
        c2 = th * step_h - pad;
        c3 = tn * bn;
      }
      int kcol = 0;
#pragma unroll 1
      for (int kh = 0; kh < taps_h; ++kh) {
#pragma unroll 1
        for (int kw = 0; kw < taps_w; ++kw) {
#pragma unroll 1
          for (int cb = 0; cb < cin_blocks; ++cb, kcol += kBlockK) {
            bar_wait_u32(eb, parity);
            if (leader) {
              if (rank == 0) bar_expect_tx_u32(fb, pair_tx);   // bytes of BOTH CTAs' boxes
              if (mode == 0) tma2d_2sm_u32(dA, mapA, lfb, cb * kBlockK, c1);
              else if (sub) tma4d_2sm_u32(dA, mapS + (uint64_t)((((kh - sub_pad) & 1) << 1) | ((kw - sub_pad) & 1)) * sizeof(CUtensorMap), lfb,
                                          cb * kBlockK, c1 + ((kw - sub_pad) >> 1), c2 + ((kh - sub_pad) >> 1), c3);
              else tma4d_2sm_u32(dA, mapA, lfb, cb * kBlockK, c1 + kw, c2 + kh, c3);
              tma2d_2sm_u32(dB, mapB, lfb, kcol, n0);
            }
            if (++stage == STAGES) { stage = 0; parity ^= 1; dA = sA0; dB = sB0; fb = full0; lfb = lfull0; eb = empty0; }
            else { dA += kAStageBytes; dB += Cfg::kBStageBytes; fb += 8; lfb += 8; eb += 8; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) ---------------------------------
    if (rank == 0) {
      const bool leader = elect_one();
      const int nkb = p.num_k_blocks;
      const uint32_t idesc = p.idesc;
      const uint32_t a_lo0 = desc_lo(smem_u32(sA)), b_lo0 = desc_lo(smem_u32(sB));
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      const uint32_t tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
      int stage = 0; uint32_t parity = 0;
      uint32_t a_lo = a_lo0, b_lo = b_lo0, fb = full0, eb = empty0;
      int local = 0;
#pragma unroll 1
      for (int tile = tile0; tile < num_tiles; tile += tile_stride, ++local) {
        const uint32_t acc = local & 1;
        bar_wait_u32(tempty0 + acc * 8, ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) {
          bar_wait_u32(fb, parity);
          tc_fence_after();
          if (leader) {
            umma2_kblock(d_tmem, a_lo, b_lo, idesc, kb != 0 ? 1u : 0u);
            commit2_multicast_u32(eb, (uint16_t)3);            // frees the stage in both CTAs
            if (kb == nkb - 1) commit2_multicast_u32(tfull0 + acc * 8, (uint16_t)3);
          }
          if (++stage == STAGES) { stage = 0; parity ^= 1; a_lo = a_lo0; b_lo = b_lo0; fb = full0; eb = empty0; }
          else { a_lo += kAStageBytes >> 4; b_lo += Cfg::kBStageBytes >> 4; fb += 8; eb += 8; }
        }
      }
    }
  } else {
    epilogue_warps<BLOCK_N, BF16, RES, Cfg::kEpiBufs>(p, &tmOut, warp, lane, sEpi, sRes, tmem_full, tmem_empty, tmem_base, num_tiles, tile0,
                                                      tile_stride, 1, (int)rank, mapa_u32(smem_u32(tmem_empty), 0), res_bar);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // neither CTA may retire while the other still signals / reads it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------
// Halo-resident 3x3 (stride 1, pad 1) convolution.
//
// The box-per-tap kernel above moves every activation nine times from L2 to shared memory (ncu:
// 727 MB through l1tex__m_xbar2l1tex for a 51 MB tensor) and is bound by that fabric (~42 B/clk/SM)
// for Cout <= 128.  Here the (bh+2) x (W+2) x 64-channel input patch of a tile of bh full output rows
// is loaded ONCE per channel block; output rows are indexed along the padded line (m = dh*(W+2)+dw,
// the two extra columns per line are discarded), so tap (kh,kw) is the SAME shared-memory tile
// read through a UMMA descriptor whose start address is shifted by (kh*(W+2)+kw) rows (the 128B
// swizzle is keyed on absolute shared-memory address bits, so a shifted start needs nothing else --
// verified on B200: base_offset = 0 is correct, (addr >> 7) & 7 is not).  Activations
// and weights use separate rings (one patch feeds nine weight boxes).
// ---------------------------------------------------------------------------------------
constexpr int kHaloABytes = 32768;            // (bh+2)*(W+2) <= 256 patch pixels x 128 B

template <int BLOCK_N>
struct HaloCfg {
  // Patch ring depth: with two stages the next tile's patch load only overlapped one tile's MMAs (~0.6 us for
  // 64 -> 64), shorter than the load latency, so the 64-wide layer sat at 1.6 us per tile; four stages (the
  // resident weights leave the room) keep three patches in flight.
  static constexpr int kAStages = BLOCK_N <= 64 ? 4 : 3;
  static constexpr int kBStageBytes = 3 * BLOCK_N * kBlockK * 2;          // the three taps of one kernel row
  static constexpr int kEpiBufs = 1;                                      // MMA-bound layers: the patch / weight rings need the room more
  static constexpr int kEpiBytes = kEpiBufs * kBlockM * BLOCK_N * 2;
  static constexpr int kRoom = (232448 - 1024 - 512 - kEpiBytes - kAStages * kHaloABytes) / kBStageBytes;
  static constexpr int kBStages = kRoom > 6 ? 6 : kRoom;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes = kAStages * kHaloABytes + kBStages * kBStageBytes + kEpiBytes + 512 + 1024;
  static_assert(kBStages >= 2, "weight ring needs at least two stages");
  static_assert(BLOCK_N > 64 || kBStages >= 3, "resident 3x3 weights need three boxes");
};

template <int BLOCK_N, bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ConvParams p) {
  using Cfg = HaloCfg<BLOCK_N>;
  constexpr int SA = Cfg::kAStages, SB = Cfg::kBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + SA * kHaloABytes;
  uint8_t* sEpi = sB + SB * Cfg::kBStageBytes;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(sEpi + Cfg::kEpiBytes);
  uint64_t* empty_a = full_a + SA;
  uint64_t* full_b = empty_a + SA;
  uint64_t* empty_b = full_b + SB;
  uint64_t* tmem_full = empty_b + SB;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < SA; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < SB; ++i) { mbar_init(&full_b[i], 1); mbar_init(&empty_b[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4 * (BLOCK_N / 32 < 4 ? BLOCK_N / 32 : 4)); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int cin_blocks = p.cin_blocks, n_tiles = p.n_tiles, tiles_h = p.tiles_h, bh = p.bh, line = p.bw;   // line = W + 2
  // Cin_p == 64 and one N tile: the nine weight taps (3 boxes) are loaded once and stay in stages 0..2 for
  // the life of the CTA.  Re-streaming them per tile (73 KB of weights against a 30 KB patch for 64 -> 64)
  // made those layers L2->SM-fill-bound: 207 us measured against an 87 us tensor floor per 512 images.
  const bool resident = p.resident_w != 0;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    const bool leader = elect_one();
    const uint32_t a_bytes = (uint32_t)p.a_rows * 128u;      // (bh+2)*(W+2) patch pixels
    const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB);
    const uint32_t fa0 = smem_u32(full_a), ea0 = smem_u32(empty_a), fb0 = smem_u32(full_b), eb0 = smem_u32(empty_b);
    const uint64_t mapA = reinterpret_cast<uint64_t>(&tmA), mapB = reinterpret_cast<uint64_t>(&tmB);
    int sa = 0, sb = 0; uint32_t pa = 1, pb = 1;
    // patches are issued one (tile, channel block) ahead of the weights that consume them
    int nt = blockIdx.x, ncb = 0;                            // next patch to issue
    auto issue_patch = [&]() {
      if (nt >= num_tiles) return;
      const int m_tile = nt / n_tiles;
      const int n_img = m_tile / tiles_h, th = m_tile - n_img * tiles_h;
      bar_wait_u32(ea0 + sa * 8, pa);
      if (leader) {
        bar_expect_tx_u32(fa0 + sa * 8, a_bytes);
        tma4d_u32(sA0 + sa * kHaloABytes, mapA, fa0 + sa * 8, ncb * kBlockK, -1, th * bh - 1, n_img);
      }
      if (++sa == SA) { sa = 0; pa ^= 1; }
      if (++ncb == cin_blocks) { ncb = 0; nt += gridDim.x; }
    };
    issue_patch();
#pragma unroll 1
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n0 = (tile % n_tiles) * BLOCK_N;
#pragma unroll 1
      for (int cb = 0; cb < cin_blocks; ++cb) {
        issue_patch();                                         // next (tile, cb) patch, if any
        if (resident && tile != (int)blockIdx.x) continue;     // resident weights: loaded with the first tile only
#pragma unroll 1
        for (int kh = 0; kh < 3; ++kh) {                       // one box = the three taps of a kernel row
          bar_wait_u32(eb0 + sb * 8, pb);
          if (leader) {
            bar_expect_tx_u32(fb0 + sb * 8, (uint32_t)Cfg::kBStageBytes);
            tma3d_u32(sB0 + sb * Cfg::kBStageBytes, mapB, fb0 + sb * 8, cb * kBlockK, n0, kh * 3);
          }
          if (++sb == SB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    const bool leader = elect_one();
    const uint32_t idesc = p.idesc;
    const uint32_t a_lo0 = desc_lo(smem_u32(sA)), b_lo0 = desc_lo(smem_u32(sB));
    const uint32_t fa0 = smem_u32(full_a), ea0 = smem_u32(empty_a), fb0 = smem_u32(full_b), eb0 = smem_u32(empty_b);
    const uint32_t tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
    // timing experiments (MIMAMO_DEBUG): 4 = every tap reads the un-shifted patch (is the shifted start address slower?),
    // 8 = one tap per kernel row (12 of 36 UMMAs per channel block: does the tile time follow the UMMA count?)
    const bool dbg_noshift = (p.debug & 4) != 0, dbg_third = (p.debug & 8) != 0;
    const uint32_t row16 = dbg_noshift ? 0u : (uint32_t)line * 8u;   // one padded line in (address >> 4) units
    const uint32_t px16 = dbg_noshift ? 0u : 8u;                     // one pixel (128 B)
    int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
    int local = 0;
#pragma unroll 1
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const uint32_t acc = local & 1;
      bar_wait_u32(tempty0 + acc * 8, ((local >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
#pragma unroll 1
      for (int cb = 0; cb < cin_blocks; ++cb) {
        bar_wait_u32(fa0 + sa * 8, pa);
        uint32_t a_row = a_lo0 + sa * (kHaloABytes >> 4);    // shifted views of the same patch: +kw rows, +kh lines
#pragma unroll 1
        for (int kh = 0; kh < 3; ++kh, a_row += row16) {
          if (resident) sb = kh;
          if (!resident || local == 0) bar_wait_u32(fb0 + sb * 8, pb);
          tc_fence_after();
          if (leader) {
            const uint32_t b_lo = b_lo0 + sb * (Cfg::kBStageBytes >> 4);
            umma_kblock(d_tmem, a_row, b_lo, idesc, (cb | kh) != 0 ? 1u : 0u);
            if (!dbg_third) {
              umma_kblock(d_tmem, a_row + px16, b_lo + (BLOCK_N * 128 >> 4), idesc, 1u);
              umma_kblock(d_tmem, a_row + 2 * px16, b_lo + 2 * (BLOCK_N * 128 >> 4), idesc, 1u);
            }
            if (!resident) commit_u32(eb0 + sb * 8);
            if (kh == 2) {
              commit_u32(ea0 + sa * 8);
              if (cb == cin_blocks - 1) commit_u32(tfull0 + acc * 8);
            }
          }
          if (!resident && ++sb == SB) { sb = 0; pb ^= 1; }
        }
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
    }
  } else {
    epilogue_warps<BLOCK_N, BF16, 0, Cfg::kEpiBufs>(p, &tmOut, warp, lane, sEpi, nullptr, tmem_full, tmem_empty, tmem_base, num_tiles, blockIdx.x, gridDim.x);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------
// conv_chain_kernel: two flat 1x1 layers in ONE persistent launch -- a layer with residual add (ResNet50's
// `_increase`) and the stride-1 layer that consumes its output (`_reduce` of the next bottleneck block).
//
// Executed layer by layer, stages 2-3 of ResNet50 are HBM-bound and the next block's `_reduce` re-reads from HBM the
// very tensor `_increase` has just written (a quarter of all bytes a stage-2 block moves).  Here every CTA, after the
// epilogue has stored the [128 pixels][N1] rows of an M tile, runs the second GEMM on exactly those rows: the producer
// warp waits until the tile's TMA stores have completed (an mbarrier the storing lanes arrive at once
// cp.async.bulk.wait_group -- completion, not .read -- has covered the store's bulk group), then streams the tile
// back through the same operand ring.  The rows were written microseconds earlier by this SM, so the loads hit in L2
// (ncu: DRAM reads of the launch = first layer's activations + residual only) and the second layer's activation read
// never reaches HBM; no shared memory is needed beyond a staging tile for the second layer's (narrow) output.
// The sequence is software pipelined by kChainLag M tiles --
//     main(m0), main(m1), main(m2), chain(m0), main(m3), chain(m1), ...
// -- so that the stores of a tile have two tile times to land before they are read back, and all three roles (producer,
// MMA issuer, epilogue) walk the same sub-tile sequence; accumulators alternate between two TMEM buffers per sub-tile.
// BN = 128-wide sub-tiles with THREE staging tiles: the residual slices of the next two sub-tiles land (by TMA) in the
// other two while the current one is processed -- with one-deep prefetch every warp's sub-tile period contained a whole
// HBM load latency (first version: 5.5 us per 128-pixel tile of stage 2 against a 3.6 us HBM floor).
// ---------------------------------------------------------------------------------------
constexpr int kChainLag = 2;                 // M tiles between a tile's main sub-tiles and its chained sub-tile
constexpr int kChainBufs = 3;                // staging / residual landing tiles

struct ChainParams {
  int m_tiles;
  int n1_tiles, nkb1;          // main layer: N1 / BN column tiles, K1 / 64 blocks
  int n2_cols, nkb2;           // chained layer: N2 <= BN output columns, N1 / 64 blocks
  uint32_t idesc1, idesc2;
  int relu1, relu2;
  int debug;                   // timing experiments only (MIMAMO_CHAIN_DEBUG bits): 1 no residual loads, 2 no main stores, 4 no chained stores
  const float* scale1; const float* shift1;
  const float* scale2; const float* shift2;
  alignas(64) CUtensorMap a1;  // main activations  [M][K1], box 64 x 128
  alignas(64) CUtensorMap b1;  // main weights      [N1][K1], box 64 x BN
  alignas(64) CUtensorMap out1;// main output       [M][N1], store box (BN/4) x 32
  alignas(64) CUtensorMap res; // residual          [M][N1], same box
  alignas(64) CUtensorMap a2;  // main output as the chained layer's activations, box 64 x 128
  alignas(64) CUtensorMap b2;  // chained weights   [N2][N1], box 64 x N2
  alignas(64) CUtensorMap out2;// chained output    [M][N2], store box (BN/4) x 32
};

template <int BN>
struct ChainCfg {
  static constexpr int kBStageBytes = BN * kBlockK * 2;
  static constexpr int kEpiBytes = kChainBufs * kBlockM * BN * 2;        // staging / residual landing tiles
  static constexpr int kEpi2Bytes = kBlockM * BN * 2;                    // chained layer's output staging (N2 <= BN)
  static constexpr int kBudget = 232448 - 1024 - 1024 - kEpiBytes - kEpi2Bytes;
  static constexpr int kMaxStages = kBudget / (kAStageBytes + kBStageBytes);
  static constexpr int kStages = kMaxStages > 8 ? 8 : kMaxStages;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * (kAStageBytes + kBStageBytes) + kEpiBytes + kEpi2Bytes + 1024 + 1024;
  static_assert(kStages >= 2, "pipeline needs at least two stages");
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_done2() { asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); }

template <int BN, bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_chain_kernel(const __grid_constant__ ChainParams p) {
  using Cfg = ChainCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kAStageBytes;
  uint8_t* sEpi = sB + STAGES * Cfg::kBStageBytes;
  uint8_t* sEpi2 = sEpi + Cfg::kEpiBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi2 + Cfg::kEpi2Bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* stored = tmem_empty + 2;                         // [4]: the main output rows of M tile ordinal s (s & 3) are in global memory
  uint64_t* res_bar = stored + 4;                            // [16 warps][kChainBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kChainBufs * kEpiWarps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n1 = p.n1_tiles, nkb1 = p.nkb1, nkb2 = p.nkb2;

  const int J = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // M tiles of this CTA (grid <= m_tiles)
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.a1); tma_prefetch_desc(&p.b1); tma_prefetch_desc(&p.out1); tma_prefetch_desc(&p.res);
    tma_prefetch_desc(&p.a2); tma_prefetch_desc(&p.b2); tma_prefetch_desc(&p.out2);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps); }
    for (int i = 0; i < 4; ++i) mbar_init(&stored[i], (uint32_t)(kEpiWarps * n1));   // every epilogue warp stores a slice of every main sub-tile
    for (int i = 0; i < kChainBufs * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    const bool leader = elect_one();
    const uint32_t tx1 = (uint32_t)kAStageBytes + (uint32_t)Cfg::kBStageBytes;
    const uint32_t tx2 = (uint32_t)kAStageBytes + (uint32_t)p.n2_cols * (kBlockK * 2);
    const uint32_t sA0 = smem_u32(sA), sB0 = smem_u32(sB), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const uint32_t stored0 = smem_u32(stored);
    const uint64_t mapA1 = reinterpret_cast<uint64_t>(&p.a1), mapB1 = reinterpret_cast<uint64_t>(&p.b1);
    const uint64_t mapA2 = reinterpret_cast<uint64_t>(&p.a2), mapB2 = reinterpret_cast<uint64_t>(&p.b2);
    int stage = 0; uint32_t parity = 1;
    uint32_t dA = sA0, dB = sB0, fb = full0, eb = empty0;
    auto load = [&](uint64_t mapA, uint64_t mapB, int kcol, int arow, int brow, uint32_t tx) {
      bar_wait_u32(eb, parity);
      if (leader) {
        bar_expect_tx_u32(fb, tx);
        tma2d_u32(dA, mapA, fb, kcol, arow);
        tma2d_u32(dB, mapB, fb, kcol, brow);
      }
      if (++stage == STAGES) { stage = 0; parity ^= 1; dA = sA0; dB = sB0; fb = full0; eb = empty0; }
      else { dA += kAStageBytes; dB += Cfg::kBStageBytes; fb += 8; eb += 8; }
    };
#pragma unroll 1
    for (int s = 0; s < J + kChainLag; ++s) {
      if (s < J) {
        const int arow = ((int)blockIdx.x + s * (int)gridDim.x) * kBlockM;
#pragma unroll 1
        for (int n = 0; n < n1; ++n)
#pragma unroll 1
          for (int kb = 0; kb < nkb1; ++kb) load(mapA1, mapB1, kb * kBlockK, arow, n * BN, tx1);
      }
      if (s >= kChainLag) {
        const int c = s - kChainLag;
        const int arow = ((int)blockIdx.x + c * (int)gridDim.x) * kBlockM;
        bar_wait_u32(stored0 + (c & 3) * 8, (c >> 2) & 1);     // the tile's rows are in global memory (L2)
        fence_proxy_async_all();                               // generic-proxy acquire -> async-proxy (TMA) reads
#pragma unroll 1
        for (int kb = 0; kb < nkb2; ++kb) load(mapA2, mapB2, kb * kBlockK, arow, 0, tx2);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    const bool leader = elect_one();
    const uint32_t a_lo0 = desc_lo(smem_u32(sA)), b_lo0 = desc_lo(smem_u32(sB));
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const uint32_t tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
    int stage = 0; uint32_t parity = 0;
    uint32_t a_lo = a_lo0, b_lo = b_lo0, fb = full0, eb = empty0;
    int t = 0;                                               // sub-tile counter (TMEM buffer = t & 1)
    auto subtile = [&](int nkb, uint32_t idesc) {
      const uint32_t acc = t & 1;
      bar_wait_u32(tempty0 + acc * 8, ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        bar_wait_u32(fb, parity);
        tc_fence_after();
        if (leader) {
          umma_kblock(d_tmem, a_lo, b_lo, idesc, kb != 0 ? 1u : 0u);
          commit_u32(eb);
          if (kb == nkb - 1) commit_u32(tfull0 + acc * 8);
        }
        if (++stage == STAGES) { stage = 0; parity ^= 1; a_lo = a_lo0; b_lo = b_lo0; fb = full0; eb = empty0; }
        else { a_lo += kAStageBytes >> 4; b_lo += Cfg::kBStageBytes >> 4; fb += 8; eb += 8; }
      }
      ++t;
    };
#pragma unroll 1
    for (int s = 0; s < J + kChainLag; ++s) {
      if (s < J) {
#pragma unroll 1
        for (int n = 0; n < n1; ++n) subtile(nkb1, p.idesc1);
      }
      if (s >= kChainLag) subtile(nkb2, p.idesc2);
    }
  } else {
    // ------------------------------- epilogue warps -------------------------------
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    constexpr int COLS = BN / 4;                              // columns per column group: 32 (SWIZZLE_64B rows) or 64 (SWIZZLE_128B)
    constexpr int ROW_BYTES = COLS * 2;
    constexpr int BUF_BYTES = kBlockM * BN * 2;
    const int row = quarter * 32 + lane;
    const uint32_t stage0_u32 = smem_u32(sEpi + half * (128 * ROW_BYTES));
    const uint32_t stage2_u32 = smem_u32(sEpi2 + half * (128 * ROW_BYTES));
    const uint32_t swz = ROW_BYTES == 128 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
    const uint64_t map_out1 = reinterpret_cast<uint64_t>(&p.out1), map_out2 = reinterpret_cast<uint64_t>(&p.out2);
    const uint64_t map_res = reinterpret_cast<uint64_t>(&p.res);
    const int total_main = J * n1;
    int t = 0, q = 0;                                        // sub-tile counter, main sub-tile counter
    int buf = 0;                                             // q % kChainBufs
    uint32_t res_parity = 0;                                 // (q / kChainBufs) & 1
    // lane 0: M tile ordinals (-1: not a main store) of the last two committed store groups, oldest first.  A main store is
    // signalled two commits later, when cp.async.bulk.wait_group 2 covers it without stalling.
    int pend0 = -1, pend1 = -1;
    auto issue_res = [&](int qq, int b) {                    // lane 0: this warp's 32 x COLS residual slice of main sub-tile qq
      const int s = qq / n1, n = qq - s * n1;
      const int m_tile = (int)blockIdx.x + s * (int)gridDim.x;
      const uint32_t bar = smem_u32(&res_bar[ew * kChainBufs + b]);
      bar_expect_tx_u32(bar, 32u * ROW_BYTES);
      tma2d_u32(stage0_u32 + (uint32_t)b * BUF_BYTES + quarter * (32 * ROW_BYTES), map_res, bar, n * BN + half * COLS,
                m_tile * kBlockM + quarter * 32);
    };
    auto committed = [&](int ordinal) {                      // lane 0, after tma_store_commit() of a group (ordinal >= 0: a main store)
      if (pend0 >= 0) {
        tma_store_wait_done2();                              // complete (not just read): everything but the two newest groups
        mbar_arrive(&stored[pend0 & 3]);
      }
      pend0 = pend1; pend1 = ordinal;
    };
    auto flush = [&]() {                                     // lane 0: no later commit will signal the pending main stores
      if (pend0 >= 0 || pend1 >= 0) {
        tma_store_wait_all();
        if (pend0 >= 0) mbar_arrive(&stored[pend0 & 3]);
        if (pend1 >= 0) mbar_arrive(&stored[pend1 & 3]);
        pend0 = pend1 = -1;
      }
    };
    // accumulator row -> scale/shift (+ residual from the staging row) -> ReLU -> 16-bit row in the staging tile
    auto drain = [&](uint32_t taddr, uint32_t my_row_u32, const float* scale, const float* shift, int relu, bool has_res) {
#pragma unroll 1
      for (int c0 = 0; c0 < COLS; c0 += 32) {
        uint32_t v[32];
        tmem_ld16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_ld16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        uint32_t rw[16];
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rw[4 * j]), "=r"(rw[4 * j + 1]), "=r"(rw[4 * j + 2]), "=r"(rw[4 * j + 3])
                         : "r"(my_row_u32 + (((uint32_t)(c0 / 8 + j) ^ swz) << 4)));
        }
        tmem_ld_wait();
        const float4* sc = reinterpret_cast<const float4*>(scale + c0);
        const float4* sh = reinterpret_cast<const float4*>(shift + c0);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 s0 = __ldg(sc + 2 * g), s1 = __ldg(sc + 2 * g + 1), t0 = __ldg(sh + 2 * g), t1 = __ldg(sh + 2 * g + 1);
          float o[8] = {__uint_as_float(v[8 * g]) * s0.x + t0.x, __uint_as_float(v[8 * g + 1]) * s0.y + t0.y,
                        __uint_as_float(v[8 * g + 2]) * s0.z + t0.z, __uint_as_float(v[8 * g + 3]) * s0.w + t0.w,
                        __uint_as_float(v[8 * g + 4]) * s1.x + t1.x, __uint_as_float(v[8 * g + 5]) * s1.y + t1.y,
                        __uint_as_float(v[8 * g + 6]) * s1.z + t1.z, __uint_as_float(v[8 * g + 7]) * s1.w + t1.w};
          if (has_res) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack2<BF16>(rw[4 * g + j]);
              o[2 * j] += f.x; o[2 * j + 1] += f.y;
            }
          }
          if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
          }
          const uint32_t chunk = (uint32_t)(c0 / 8 + g);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row_u32 + ((chunk ^ swz) << 4)), "r"(pack2<BF16>(o[0], o[1])),
                       "r"(pack2<BF16>(o[2], o[3])), "r"(pack2<BF16>(o[4], o[5])), "r"(pack2<BF16>(o[6], o[7])) : "memory");
        }
      }
    };
    const bool dbg_nores = (p.debug & 1) != 0, dbg_nostore = (p.debug & 2) != 0, dbg_nostore2 = (p.debug & 4) != 0;
    if (lane == 0 && !dbg_nores) {
      for (int i = 0; i < kChainBufs - 1 && i < total_main; ++i) issue_res(i, i);
    }
#pragma unroll 1
    for (int s = 0; s < J + kChainLag; ++s) {
      if (s < J) {
        const int m_tile = (int)blockIdx.x + s * (int)gridDim.x;
#pragma unroll 1
        for (int n = 0; n < n1; ++n, ++t, ++q) {
          const int acc = t & 1;
          mbar_wait(&tmem_full[acc], (uint32_t)((t >> 1) & 1));
          tc_fence_after();
          if (!dbg_nores) mbar_wait(&res_bar[ew * kChainBufs + buf], res_parity);
          const uint32_t stage_u32 = stage0_u32 + (uint32_t)buf * BUF_BYTES;
          const int n0 = n * BN + half * COLS;
          drain(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + half * COLS, stage_u32 + row * ROW_BYTES,
                p.scale1 + n0, p.shift1 + n0, p.relu1, true);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // the staging tile of sub-tile q - 1 (its store was committed a whole sub-tile ago) takes the residual of q + 2
            const int nb = buf == 0 ? kChainBufs - 1 : buf - 1;
            if (q + kChainBufs - 1 < total_main && !dbg_nores) {
              tma_store_wait_read();
              issue_res(q + kChainBufs - 1, nb);
            }
            if (!dbg_nostore) tma_store_2d(map_out1, stage_u32 + quarter * (32 * ROW_BYTES), n0, m_tile * kBlockM + quarter * 32);
            tma_store_commit();
            committed(s);
            if (q + 1 == total_main) flush();                // nothing follows that is guaranteed to signal them
          }
          if (++buf == kChainBufs) { buf = 0; res_parity ^= 1; }
        }
      }
      if (s >= kChainLag) {
        const int m_tile = (int)blockIdx.x + (s - kChainLag) * (int)gridDim.x;
        const int acc = t & 1;
        mbar_wait(&tmem_full[acc], (uint32_t)((t >> 1) & 1));
        tc_fence_after();
        if (half * COLS < p.n2_cols) {
          if (lane == 0) tma_store_wait_read();              // the previous chained store has read this slice
          __syncwarp();
          const int n0 = half * COLS;
          drain(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + half * COLS, stage2_u32 + row * ROW_BYTES,
                p.scale2 + n0, p.shift2 + n0, p.relu2, false);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (!dbg_nostore2) tma_store_2d(map_out2, stage2_u32 + quarter * (32 * ROW_BYTES), n0, m_tile * kBlockM + quarter * 32);
            tma_store_commit();
            committed(-1);
          }
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        ++t;
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------
// conv_chain_smem_kernel: the same two layers with the hand-over ON CHIP (ResNet50 stage 2: K1 = 64, N1 = 256, N2 = 64).
//
// Measured on the L2 hand-over above: with the activation tile of the second layer (64 KB per 128 pixels) going back
// through a 3-stage operand ring behind the first layer's loads, every M tile exposed two load latencies in sequence
// (3.4 us per tile with all residual loads and stores switched off, against a 3.6 us HBM floor with them on).  Both weight
// matrices of a stage-2 pair are 32 KB, so here they stay RESIDENT in shared memory, and the epilogue's 16-bit staging
// tile of a 128-column sub-tile -- two [128 pixels][64 channels] SWIZZLE_128B blocks, which is exactly the K-major
// operand layout UMMA reads -- is the second layer's A operand: once the epilogue warps have written and fenced a
// sub-tile, the MMA warp multiplies it by the matching two K blocks of the second layer's weights into a separate TMEM
// accumulator, while the TMA store of the same tile is in flight.  The ring carries one 16 KB activation box per M tile.
//
// Epilogue: the fixed latencies of one sub-tile round (barrier wake-up, tcgen05.ld, the proxy fence, TMA issue) cost a
// warp ~1.5 us whatever the tile size (measured with every load and store switched off), so the 16 warps form TWO
// groups of eight that work on alternate sub-tiles -- group = sub-tile parity = TMEM accumulator buffer; a warp owns
// 32 rows x 64 columns, i.e. whole 128-byte staging rows, and its own residual / store boxes -- and two sub-tiles are
// always in flight.  Staging tile q & 3 serves sub-tile q: each group rotates through two tiles of its own, the
// residual of sub-tile q + 2 is requested when the group starts sub-tile q (after the store of q - 2 has read the tile
// and the chained MMAs on it have retired: tcgen05.commit -> buf_free).  The 64 chained columns of an M tile are drained
// by the group that finished the tile (32 columns per warp, warp pairs share a staging row and one store).
// TMEM: 2 x 128 columns (main accumulators, alternating per sub-tile) + 2 x 64 (chained accumulators, per M tile).
// ---------------------------------------------------------------------------------------
constexpr int kChainSmemBufs = 4;

struct ChainSmemCfg {
  static constexpr int BN = 128;
  static constexpr int N1 = 256;
  static constexpr int N2 = 64;
  static constexpr int kWBytes = 65536;                                   // [256][64] + [64][256]
  static constexpr int kEpiBytes = kChainSmemBufs * kBlockM * BN * 2;
  static constexpr int kEpi2Bytes = kBlockM * N2 * 2;
  static constexpr int kTmemCols = 512;                                   // 256 + 128 used; allocations are powers of two
  static constexpr int kSmemBytes = kAStageBytes + kWBytes + kEpiBytes + kEpi2Bytes + 1024 + 1024;   // ONE activation stage (16 KB per 3.6 us)
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_chain_smem_kernel(const __grid_constant__ ChainParams p) {
  using Cfg = ChainSmemCfg;
  constexpr int BN = Cfg::BN, NB = kChainSmemBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sW = sA + kAStageBytes;                           // two boxes [128][64] of the main weights, then four boxes [64][64] of the chained layer
  uint8_t* sEpi = sW + Cfg::kWBytes;                         // NB x [2 column halves][128 rows][128 B]
  uint8_t* sEpi2 = sEpi + Cfg::kEpiBytes;                    // [128 rows][128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sEpi2 + Cfg::kEpi2Bytes);
  uint64_t* empty_bar = full_bar + 1;
  uint64_t* tmem_full = empty_bar + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* chain_full = tmem_empty + 2;
  uint64_t* chain_empty = chain_full + 2;
  uint64_t* staged = chain_empty + 2;                        // [NB]: the eight warps of the group have written (and fenced) the sub-tile in staging tile b
  uint64_t* buf_free = staged + NB;                          // [NB]: the chained MMAs reading staging tile b have retired
  uint64_t* res_bar = buf_free + NB;                         // [16 warps][2]
  uint64_t* wfull = res_bar + 2 * kEpiWarps;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_main = 2 * J;                              // two 128-column sub-tiles per M tile
  constexpr uint32_t kW1Bytes = Cfg::N1 * kBlockK * 2;       // 32 KB
  constexpr uint32_t kW2Box = Cfg::N2 * kBlockK * 2;         // 8 KB
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.a1); tma_prefetch_desc(&p.b1); tma_prefetch_desc(&p.out1); tma_prefetch_desc(&p.res);
    tma_prefetch_desc(&p.b2); tma_prefetch_desc(&p.out2);
    mbar_init(full_bar, 1); mbar_init(empty_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8);
      mbar_init(&chain_full[i], 1); mbar_init(&chain_empty[i], 8);
    }
    for (int i = 0; i < NB; ++i) { mbar_init(&staged[i], 8); mbar_init(&buf_free[i], 1); }
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_chain = tmem_base + 2 * BN;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (elect_one()) {
      const uint64_t mapA1 = reinterpret_cast<uint64_t>(&p.a1), mapB1 = reinterpret_cast<uint64_t>(&p.b1), mapB2 = reinterpret_cast<uint64_t>(&p.b2);
      const uint32_t wb = smem_u32(wfull), sW0 = smem_u32(sW);
      bar_expect_tx_u32(wb, kW1Bytes + 4u * kW2Box);
      for (int n = 0; n < 2; ++n) tma2d_u32(sW0 + (uint32_t)n * BN * (kBlockK * 2), mapB1, wb, 0, n * BN);
      for (int kb = 0; kb < 4; ++kb) tma2d_u32(sW0 + kW1Bytes + (uint32_t)kb * kW2Box, mapB2, wb, kb * kBlockK, 0);
      const uint32_t fb = smem_u32(full_bar), eb = smem_u32(empty_bar);
#pragma unroll 1
      for (int c = 0; c < J; ++c) {
        bar_wait_u32(eb, (uint32_t)(c & 1) ^ 1u);
        bar_expect_tx_u32(fb, (uint32_t)kAStageBytes);
        tma2d_u32(smem_u32(sA), mapA1, fb, 0, ((int)blockIdx.x + c * (int)gridDim.x) * kBlockM);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    const bool leader = elect_one();
    const uint32_t a_lo = desc_lo(smem_u32(sA)), w1_lo = desc_lo(smem_u32(sW)), w2_lo = desc_lo(smem_u32(sW) + kW1Bytes);
    const uint32_t e_lo0 = desc_lo(smem_u32(sEpi));
    const uint32_t tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
    const uint32_t cfull0 = smem_u32(chain_full), cempty0 = smem_u32(chain_empty);
    const uint32_t staged0 = smem_u32(staged), bfree0 = smem_u32(buf_free);
    // the chained layer's K blocks 2n, 2n + 1 for main sub-tile qq = 2c + n, read from staging tile qq & 3
    auto chain_part = [&](int qq) {
      const int c = qq >> 1, n = qq & 1, b = qq & (NB - 1);
      bar_wait_u32(staged0 + b * 8, (uint32_t)(qq >> 2) & 1u);
      if (n == 0) bar_wait_u32(cempty0 + (c & 1) * 8, (uint32_t)((c >> 1) & 1) ^ 1u);
      tc_fence_after();
      if (leader) {
        const uint32_t d = tmem_chain + (c & 1) * Cfg::N2;
        const uint32_t e_lo = e_lo0 + (uint32_t)b * (kBlockM * BN * 2 >> 4);
        umma_kblock(d, e_lo, w2_lo + (uint32_t)(2 * n) * (kW2Box >> 4), p.idesc2, n != 0 ? 1u : 0u);
        umma_kblock(d, e_lo + (kBlockM * 128 >> 4), w2_lo + (uint32_t)(2 * n + 1) * (kW2Box >> 4), p.idesc2, 1u);
        commit_u32(bfree0 + b * 8);
        if (n == 1) commit_u32(cfull0 + (c & 1) * 8);
      }
    };
    bar_wait_u32(smem_u32(wfull), 0);
#pragma unroll 1
    for (int q = 0; q < total_main; ++q) {
      const int c = q >> 1, n = q & 1;
      if (n == 0) bar_wait_u32(smem_u32(full_bar), (uint32_t)(c & 1));
      bar_wait_u32(tempty0 + n * 8, (uint32_t)(c & 1) ^ 1u);   // accumulator buffer = q & 1 = n, its use count = c
      tc_fence_after();
      if (leader) {
        umma_kblock(tmem_base + n * BN, a_lo, w1_lo + (uint32_t)n * (BN * (kBlockK * 2) >> 4), p.idesc1, 0u);
        commit_u32(tfull0 + n * 8);
        if (n == 1) commit_u32(smem_u32(empty_bar));
      }
      if (q >= 1) chain_part(q - 1);
    }
    if (total_main > 0) chain_part(total_main - 1);
  } else {
    // ------------------------------- epilogue warps -------------------------------
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int g = half & 1;                                  // 64-column half of the sub-tile = one [128 rows][128 B] staging block
    const int par = half >> 1;                               // group: handles sub-tiles q with q & 1 == par (column sub-tile n = par of every M tile)
    constexpr int BUF_BYTES = kBlockM * BN * 2;
    const int row = quarter * 32 + lane;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t slice0_u32 = smem_u32(sEpi) + g * (kBlockM * 128) + quarter * (32 * 128);   // this warp's 32 x 128 B slice of staging tile 0
    const uint32_t row0_u32 = smem_u32(sEpi) + g * (kBlockM * 128) + row * 128;
    const uint64_t map_out1 = reinterpret_cast<uint64_t>(&p.out1), map_out2 = reinterpret_cast<uint64_t>(&p.out2);
    const uint64_t map_res = reinterpret_cast<uint64_t>(&p.res);
    const bool dbg_nores = (p.debug & 1) != 0, dbg_nostore = (p.debug & 2) != 0, dbg_nostore2 = (p.debug & 4) != 0;
    auto issue_res = [&](int qq) {                           // lane 0: this warp's 32 x 64 residual slice of main sub-tile qq (same group)
      const int b = qq & (NB - 1);
      const uint32_t bar = smem_u32(&res_bar[ew * 2 + (b >> 1)]);
      bar_expect_tx_u32(bar, 32u * 128u);
      tma2d_u32(slice0_u32 + (uint32_t)b * BUF_BYTES, map_res, bar, (qq & 1) * BN + g * 64, ((int)blockIdx.x + (qq >> 1) * (int)gridDim.x) * kBlockM + quarter * 32);
    };
    // accumulator row, 32 columns -> scale/shift (+ residual from the staging row) -> ReLU -> 16-bit, into 16-byte chunks chunk0 .. chunk0 + 3 of the row
    auto drain32 = [&](uint32_t taddr, uint32_t my_row_u32, uint32_t chunk0, const float* scale, const float* shift, int relu, bool has_res) {
      uint32_t v[32];
      tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      tmem_ld16(taddr + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
      uint32_t rw[16];
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rw[4 * j]), "=r"(rw[4 * j + 1]), "=r"(rw[4 * j + 2]), "=r"(rw[4 * j + 3])
                       : "r"(my_row_u32 + (((chunk0 + (uint32_t)j) ^ swz) << 4)));
      }
      tmem_ld_wait();
      const float4* sc = reinterpret_cast<const float4*>(scale);
      const float4* sh = reinterpret_cast<const float4*>(shift);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 s0 = __ldg(sc + 2 * k), s1 = __ldg(sc + 2 * k + 1), t0 = __ldg(sh + 2 * k), t1 = __ldg(sh + 2 * k + 1);
        float o[8] = {__uint_as_float(v[8 * k]) * s0.x + t0.x, __uint_as_float(v[8 * k + 1]) * s0.y + t0.y,
                      __uint_as_float(v[8 * k + 2]) * s0.z + t0.z, __uint_as_float(v[8 * k + 3]) * s0.w + t0.w,
                      __uint_as_float(v[8 * k + 4]) * s1.x + t1.x, __uint_as_float(v[8 * k + 5]) * s1.y + t1.y,
                      __uint_as_float(v[8 * k + 6]) * s1.z + t1.z, __uint_as_float(v[8 * k + 7]) * s1.w + t1.w};
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack2<BF16>(rw[4 * k + j]);
            o[2 * j] += f.x; o[2 * j + 1] += f.y;
          }
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row_u32 + (((chunk0 + (uint32_t)k) ^ swz) << 4)), "r"(pack2<BF16>(o[0], o[1])),
                     "r"(pack2<BF16>(o[2], o[3])), "r"(pack2<BF16>(o[4], o[5])), "r"(pack2<BF16>(o[6], o[7])) : "memory");
      }
    };
    // the 64 chained output columns of M tile ordinal cc: group 1 (it finishes every M tile), 32 columns per warp; the two
    // warps of a lane quarter share the 128-byte staging rows and one store (64-thread named barrier 1 + quarter)
    auto chain_epilogue = [&](int cc) {
      mbar_wait(&chain_full[cc & 1], (uint32_t)((cc >> 1) & 1));
      tc_fence_after();
      if (g == 0 && lane == 0) tma_store_wait_read();        // the previous chained store (issued by this lane) has read the rows
      named_bar_sync(1 + quarter, 64);
      drain32(tmem_chain + ((uint32_t)(quarter * 32) << 16) + (cc & 1) * Cfg::N2 + g * 32, smem_u32(sEpi2) + row * 128, (uint32_t)(4 * g),
              p.scale2 + g * 32, p.shift2 + g * 32, p.relu2, false);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&chain_empty[cc & 1]);
      fence_proxy_async_smem();
      named_bar_sync(1 + quarter, 64);
      if (g == 0 && lane == 0) {
        if (!dbg_nostore2) tma_store_2d(map_out2, smem_u32(sEpi2) + quarter * (32 * 128), 0, ((int)blockIdx.x + cc * (int)gridDim.x) * kBlockM + quarter * 32);
        tma_store_commit();
      }
    };
    if (lane == 0 && !dbg_nores && par < total_main) issue_res(par);
#pragma unroll 1
    for (int q = par; q < total_main; q += 2) {
      const int c = q >> 1, b = q & (NB - 1);
      if (lane == 0 && !dbg_nores && q + 2 < total_main) {
        // the group's other staging tile takes the residual of sub-tile q + 2: the store of q - 2 (committed a round ago) must
        // have read it and the chained MMAs on it must have retired
        if (q >= 2) {
          tma_store_wait_read();
          mbar_wait(&buf_free[b ^ 2], (uint32_t)((q - 2) >> 2) & 1u);
        }
        issue_res(q + 2);
      }
      if (par == 1 && c >= 1) chain_epilogue(c - 1);         // its MMAs were issued a round ago
      mbar_wait(&tmem_full[par], (uint32_t)(c & 1));
      tc_fence_after();
      if (!dbg_nores) mbar_wait(&res_bar[ew * 2 + (b >> 1)], (uint32_t)(q >> 2) & 1u);
      const uint32_t my_row_u32 = row0_u32 + (uint32_t)b * BUF_BYTES;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + par * BN + g * 64;
      const int n0 = par * BN + g * 64;
      drain32(taddr, my_row_u32, 0u, p.scale1 + n0, p.shift1 + n0, p.relu1, !dbg_nores);
      drain32(taddr + 32, my_row_u32, 4u, p.scale1 + n0 + 32, p.shift1 + n0 + 32, p.relu1, !dbg_nores);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[par]);
      fence_proxy_async_smem();                              // generic-proxy writes -> visible to the tensor core and the TMA engine
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&staged[b]);
        if (!dbg_nostore)
          tma_store_2d(map_out1, slice0_u32 + (uint32_t)b * BUF_BYTES, n0, ((int)blockIdx.x + c * (int)gridDim.x) * kBlockM + quarter * 32);
        tma_store_commit();
      }
    }
    if (par == 1 && J > 0) chain_epilogue(J - 1);
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------
// conv_chain_stream_kernel: the on-chip hand-over for ResNet50 stage 3 (K1 = 128, N1 = 512, N2 = 128), where the two weight
// matrices (128 KB each) cannot stay in shared memory.  Same structure as conv_chain_smem_kernel -- staging tile = A operand of
// the chained GEMM, two alternating epilogue groups, four staging tiles -- with the weights STREAMED through a small ring of
// [128][64] boxes in the static order the MMA warp consumes them: per 128-column sub-tile q the two K blocks of the main weights,
// then the two K blocks of the chained weights that belong to sub-tile q - 1.  The M tile's activations (two K blocks, 32 KB) are
// loaded once and serve its four sub-tiles; the single buffer is enough because the MMA warp runs one to two sub-tiles ahead of
// the epilogue.  The L2 hand-over moved 800 KB per 128 pixels between L2 and the SM (the ~10 TB/s L2 cap); this one moves 576 KB.
// TMEM: 2 x 128 columns main + 2 x 128 chained = all 512.
// ---------------------------------------------------------------------------------------
struct ChainStreamCfg {
  static constexpr int BN = 128;
  static constexpr int N1 = 512, K1 = 128, N2 = 128;
  static constexpr int kBStages = 2;
  static constexpr int kBBox = BN * kBlockK * 2;                          // 16 KB
  static constexpr int kABytes = 2 * kAStageBytes;                        // the M tile's two K blocks
  static constexpr int kEpiBytes = kChainSmemBufs * kBlockM * BN * 2;
  static constexpr int kEpi2Bytes = kBlockM * N2 * 2;
  static constexpr int kTmemCols = 512;
  static constexpr int kSmemBytes = kABytes + kBStages * kBBox + kEpiBytes + kEpi2Bytes + 1024 + 1024;
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_chain_stream_kernel(const __grid_constant__ ChainParams p) {
  using Cfg = ChainStreamCfg;
  constexpr int BN = Cfg::BN, NB = kChainSmemBufs, SB = Cfg::kBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                        // [2 K blocks][128 rows][128 B]
  uint8_t* sB = sA + Cfg::kABytes;                           // SB x [128 rows][128 B] weight boxes
  uint8_t* sEpi = sB + SB * Cfg::kBBox;                      // NB x [2 column halves][128 rows][128 B]
  uint8_t* sEpi2 = sEpi + Cfg::kEpiBytes;                    // [2 column halves][128 rows][128 B]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sEpi2 + Cfg::kEpi2Bytes);
  uint64_t* a_empty = a_full + 1;
  uint64_t* b_full = a_empty + 1;
  uint64_t* b_empty = b_full + SB;
  uint64_t* tmem_full = b_empty + SB;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* chain_full = tmem_empty + 2;
  uint64_t* chain_empty = chain_full + 2;
  uint64_t* staged = chain_empty + 2;                        // [NB]
  uint64_t* buf_free = staged + NB;                          // [NB]
  uint64_t* res_bar = buf_free + NB;                         // [16 warps][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * kEpiWarps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_main = 4 * J;                              // four 128-column sub-tiles per M tile
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.a1); tma_prefetch_desc(&p.b1); tma_prefetch_desc(&p.out1); tma_prefetch_desc(&p.res);
    tma_prefetch_desc(&p.b2); tma_prefetch_desc(&p.out2);
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8);
      mbar_init(&chain_full[i], 1); mbar_init(&chain_empty[i], 8);
    }
    for (int i = 0; i < NB; ++i) { mbar_init(&staged[i], 8); mbar_init(&buf_free[i], 1); }
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_chain = tmem_base + 2 * BN;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (elect_one()) {
      const uint64_t mapA1 = reinterpret_cast<uint64_t>(&p.a1), mapB1 = reinterpret_cast<uint64_t>(&p.b1), mapB2 = reinterpret_cast<uint64_t>(&p.b2);
      const uint32_t af = smem_u32(a_full), ae = smem_u32(a_empty), sA0 = smem_u32(sA), sB0 = smem_u32(sB);
      int bs = 0; uint32_t bpar = 1;
      auto load_b = [&](uint64_t map, int c0, int c1) {
        bar_wait_u32(smem_u32(&b_empty[bs]), bpar);
        const uint32_t fb = smem_u32(&b_full[bs]);
        bar_expect_tx_u32(fb, (uint32_t)Cfg::kBBox);
        tma2d_u32(sB0 + bs * Cfg::kBBox, map, fb, c0, c1);
        if (++bs == SB) { bs = 0; bpar ^= 1; }
      };
#pragma unroll 1
      for (int q = 0; q < total_main; ++q) {
        const int c = q >> 2, n = q & 3;
        if (n == 0) {                                          // the M tile's activations: both K blocks behind one barrier
          const int row = ((int)blockIdx.x + c * (int)gridDim.x) * kBlockM;
          bar_wait_u32(ae, (uint32_t)(c & 1) ^ 1u);
          bar_expect_tx_u32(af, (uint32_t)Cfg::kABytes);
          tma2d_u32(sA0, mapA1, af, 0, row);
          tma2d_u32(sA0 + kAStageBytes, mapA1, af, kBlockK, row);
        }
        load_b(mapB1, 0, n * BN);
        load_b(mapB1, kBlockK, n * BN);
        if (q >= 1) {                                          // chained weights for the previous sub-tile's 128 channels
          const int pn = (q - 1) & 3;
          load_b(mapB2, (2 * pn) * kBlockK, 0);
          load_b(mapB2, (2 * pn + 1) * kBlockK, 0);
        }
      }
      if (total_main > 0) { load_b(mapB2, 6 * kBlockK, 0); load_b(mapB2, 7 * kBlockK, 0); }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    const bool leader = elect_one();
    const uint32_t a_lo = desc_lo(smem_u32(sA)), b_lo0 = desc_lo(smem_u32(sB)), e_lo0 = desc_lo(smem_u32(sEpi));
    const uint32_t tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
    const uint32_t cfull0 = smem_u32(chain_full), cempty0 = smem_u32(chain_empty);
    const uint32_t staged0 = smem_u32(staged), bfree0 = smem_u32(buf_free);
    const uint32_t bfull0 = smem_u32(b_full), bempty0 = smem_u32(b_empty);
    int bs = 0; uint32_t bpar = 0;
    // one K block: A descriptor a, the next weight box of the ring, into accumulator d
    auto kblock = [&](uint32_t d, uint32_t a, uint32_t idesc, uint32_t accumulate) {
      bar_wait_u32(bfull0 + bs * 8, bpar);
      tc_fence_after();
      if (leader) {
        umma_kblock(d, a, b_lo0 + (uint32_t)bs * (Cfg::kBBox >> 4), idesc, accumulate);
        commit_u32(bempty0 + bs * 8);
      }
      if (++bs == SB) { bs = 0; bpar ^= 1; }
    };
    auto chain_part = [&](int qq) {
      const int c = qq >> 2, n = qq & 3, b = qq & (NB - 1);
      bar_wait_u32(staged0 + b * 8, (uint32_t)(qq >> 2) & 1u);
      if (n == 0) bar_wait_u32(cempty0 + (c & 1) * 8, (uint32_t)((c >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_chain + (c & 1) * Cfg::N2;
      const uint32_t e_lo = e_lo0 + (uint32_t)b * (kBlockM * BN * 2 >> 4);
      kblock(d, e_lo, p.idesc2, n != 0 ? 1u : 0u);
      kblock(d, e_lo + (kBlockM * 128 >> 4), p.idesc2, 1u);
      if (leader) {
        commit_u32(bfree0 + b * 8);
        if (n == 3) commit_u32(cfull0 + (c & 1) * 8);
      }
    };
#pragma unroll 1
    for (int q = 0; q < total_main; ++q) {
      const int c = q >> 2, n = q & 3;
      if (n == 0) bar_wait_u32(smem_u32(a_full), (uint32_t)(c & 1));
      const uint32_t acc = q & 1;
      bar_wait_u32(tempty0 + acc * 8, (uint32_t)((q >> 1) & 1) ^ 1u);
      tc_fence_after();
      kblock(tmem_base + acc * BN, a_lo, p.idesc1, 0u);
      kblock(tmem_base + acc * BN, a_lo + (kAStageBytes >> 4), p.idesc1, 1u);
      if (leader) {
        commit_u32(tfull0 + acc * 8);
        if (n == 3) commit_u32(smem_u32(a_empty));
      }
      if (q >= 1) chain_part(q - 1);
    }
    if (total_main > 0) chain_part(total_main - 1);
  } else {
    // ------------------------------- epilogue warps -------------------------------
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int g = half & 1;                                  // 64-column half of a sub-tile = one [128 rows][128 B] staging block
    const int par = half >> 1;                               // group: sub-tiles q with q & 1 == par (column sub-tiles n = par, par + 2)
    constexpr int BUF_BYTES = kBlockM * BN * 2;
    const int row = quarter * 32 + lane;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t slice0_u32 = smem_u32(sEpi) + g * (kBlockM * 128) + quarter * (32 * 128);
    const uint32_t row0_u32 = smem_u32(sEpi) + g * (kBlockM * 128) + row * 128;
    const uint32_t slice2_u32 = smem_u32(sEpi2) + g * (kBlockM * 128) + quarter * (32 * 128);
    const uint32_t row2_u32 = smem_u32(sEpi2) + g * (kBlockM * 128) + row * 128;
    const uint64_t map_out1 = reinterpret_cast<uint64_t>(&p.out1), map_out2 = reinterpret_cast<uint64_t>(&p.out2);
    const uint64_t map_res = reinterpret_cast<uint64_t>(&p.res);
    auto tile_row = [&](int c) { return ((int)blockIdx.x + c * (int)gridDim.x) * kBlockM + quarter * 32; };
    auto issue_res = [&](int qq) {                           // lane 0: this warp's 32 x 64 residual slice of main sub-tile qq (same group)
      const int b = qq & (NB - 1);
      const uint32_t bar = smem_u32(&res_bar[ew * 2 + (b >> 1)]);
      bar_expect_tx_u32(bar, 32u * 128u);
      tma2d_u32(slice0_u32 + (uint32_t)b * BUF_BYTES, map_res, bar, (qq & 3) * BN + g * 64, tile_row(qq >> 2));
    };
    auto drain32 = [&](uint32_t taddr, uint32_t my_row_u32, uint32_t chunk0, const float* scale, const float* shift, int relu, bool has_res) {
      uint32_t v[32];
      tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
      tmem_ld16(taddr + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
      uint32_t rw[16];
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rw[4 * j]), "=r"(rw[4 * j + 1]), "=r"(rw[4 * j + 2]), "=r"(rw[4 * j + 3])
                       : "r"(my_row_u32 + (((chunk0 + (uint32_t)j) ^ swz) << 4)));
      }
      tmem_ld_wait();
      const float4* sc = reinterpret_cast<const float4*>(scale);
      const float4* sh = reinterpret_cast<const float4*>(shift);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 s0 = __ldg(sc + 2 * k), s1 = __ldg(sc + 2 * k + 1), t0 = __ldg(sh + 2 * k), t1 = __ldg(sh + 2 * k + 1);
        float o[8] = {__uint_as_float(v[8 * k]) * s0.x + t0.x, __uint_as_float(v[8 * k + 1]) * s0.y + t0.y,
                      __uint_as_float(v[8 * k + 2]) * s0.z + t0.z, __uint_as_float(v[8 * k + 3]) * s0.w + t0.w,
                      __uint_as_float(v[8 * k + 4]) * s1.x + t1.x, __uint_as_float(v[8 * k + 5]) * s1.y + t1.y,
                      __uint_as_float(v[8 * k + 6]) * s1.z + t1.z, __uint_as_float(v[8 * k + 7]) * s1.w + t1.w};
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack2<BF16>(rw[4 * k + j]);
            o[2 * j] += f.x; o[2 * j + 1] += f.y;
          }
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row_u32 + (((chunk0 + (uint32_t)k) ^ swz) << 4)), "r"(pack2<BF16>(o[0], o[1])),
                     "r"(pack2<BF16>(o[2], o[3])), "r"(pack2<BF16>(o[4], o[5])), "r"(pack2<BF16>(o[6], o[7])) : "memory");
      }
    };
    // the 128 chained output columns of M tile ordinal cc: group 1 (it finishes every M tile), 64 columns (whole staging rows) per warp
    auto chain_epilogue = [&](int cc) {
      mbar_wait(&chain_full[cc & 1], (uint32_t)((cc >> 1) & 1));
      tc_fence_after();
      if (lane == 0) tma_store_wait_read();                  // this warp's previous chained store has read its slice
      __syncwarp();
      const uint32_t taddr = tmem_chain + ((uint32_t)(quarter * 32) << 16) + (cc & 1) * Cfg::N2 + g * 64;
      drain32(taddr, row2_u32, 0u, p.scale2 + g * 64, p.shift2 + g * 64, p.relu2, false);
      drain32(taddr + 32, row2_u32, 4u, p.scale2 + g * 64 + 32, p.shift2 + g * 64 + 32, p.relu2, false);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&chain_empty[cc & 1]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(map_out2, slice2_u32, g * 64, tile_row(cc));
        tma_store_commit();
      }
    };
    if (lane == 0 && par < total_main) issue_res(par);
#pragma unroll 1
    for (int q = par; q < total_main; q += 2) {
      const int c = q >> 2, n = q & 3, b = q & (NB - 1);
      if (lane == 0 && q + 2 < total_main) {
        if (q >= 2) {                                        // the group's other staging tile: store read, chained MMAs on it retired
          tma_store_wait_read();
          mbar_wait(&buf_free[b ^ 2], (uint32_t)((q - 2) >> 2) & 1u);
        }
        issue_res(q + 2);
      }
      if (par == 1 && n == 1 && c >= 1) chain_epilogue(c - 1);   // its MMAs were issued behind sub-tile (c - 1, 3)
      mbar_wait(&tmem_full[par], (uint32_t)((q >> 1) & 1));
      tc_fence_after();
      mbar_wait(&res_bar[ew * 2 + (b >> 1)], (uint32_t)(q >> 2) & 1u);
      const uint32_t my_row_u32 = row0_u32 + (uint32_t)b * BUF_BYTES;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + par * BN + g * 64;
      const int n0 = n * BN + g * 64;
      drain32(taddr, my_row_u32, 0u, p.scale1 + n0, p.shift1 + n0, p.relu1, true);
      drain32(taddr + 32, my_row_u32, 4u, p.scale1 + n0 + 32, p.shift1 + n0 + 32, p.relu1, true);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[par]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&staged[b]);
        tma_store_2d(map_out1, slice0_u32 + (uint32_t)b * BUF_BYTES, n0, tile_row(c));
        tma_store_commit();
      }
    }
    if (par == 1 && J > 0) chain_epilogue(J - 1);
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------
// conv1_7x7_s2 as a "line" kernel.  After space-to-depth the layer is a 4x4 / stride-1 convolution
// over 16-channel pixels.  One tile = one output line (112 pixels): the 4-row input patch is loaded once
// and each of the 16 taps is a K=16 UMMA on a row-shifted view of it; the whole 64 x 256 weight matrix
// (32 KB) stays resident.
// The input is stored chunk-planar, [B][115 rows][2 chunks][115 px][8 ch] (nn_kernels.cuh): for a fixed
// 16-byte channel chunk the pixels of a row are contiguous, which IS the un-swizzled K-major UMMA layout
// (8-row core matrices of 128 contiguous bytes, SBO = 128 B; the second K chunk LBO = 1840 B further).
// Four consecutive rows are one contiguous 14 720-byte block, fetched by a single cp.async.bulk.  The first
// version of this kernel used a SWIZZLE_32B tensor map whose 460 32-byte box rows per tile kept the TMA unit
// busy for ~1900 cycles (the kernel ran at 1.0 us per line against 0.27 us of MMA time).
// ---------------------------------------------------------------------------------------
constexpr int kLineABytes = 16384;
constexpr int kLineAStages = 6;
constexpr int kLineWBytes = 32768;
constexpr int kLinePlane = 115 * 16;                         // one (row, chunk) plane: 115 pixels x 16 B
constexpr uint32_t kDescHi32 = 16u | (1u << 14) | (6u << 29);   // weights: SBO = 256 B, version 1, SWIZZLE_32B
constexpr uint32_t kDescHiFlat = 8u | (1u << 14);               // patch: SBO = 128 B, version 1, no swizzle
constexpr int kLineEpiBytes = 2 * kBlockM * 64 * 2;           // double-buffered staging
constexpr int kLineSmemBytes = kLineWBytes + kLineAStages * kLineABytes + kLineEpiBytes + 512 + 1024;

__device__ __forceinline__ void bulk_load_u32(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void umma_one(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate,
                                         uint32_t a_hi, uint32_t b_hi) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %6};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(b_hi)
      : "memory");
}

// POOL = true additionally fuses pool1_3x3_s2 (MaxPool 3x3, stride 2, pad 0, ceil_mode; 112x112 -> 56x56) into the
// epilogue, so the 1.6 MB-per-image conv1 output is never written or re-read.  A CTA then walks bands of 8 pooled
// rows (conv rows 16b .. 16b+16): the vertical maximum over three conv rows is an elementwise max of three
// consecutive tiles in the SAME thread (thread = output column, 32 channels; the shared row 2i+2 is carried over in
// registers), and the horizontal 3-wide / stride-2 maximum goes through a double-buffered shared-memory line.  The
// maximum is taken on the 16-bit rounded post-BN/ReLU values, i.e. exactly what a separate pooling kernel would read.
constexpr int kPoolBand = 8;                                 // pooled rows per work unit
constexpr int kPoolBands = 56 / kPoolBand;

template <bool BF16>
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  if (BF16) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

template <bool BF16, bool POOL>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv1_line_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ConvParams p) {
  constexpr int SA = kLineAStages, BLOCK_N = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sA = smem + kLineWBytes;
  uint8_t* sEpi = sA + SA * kLineABytes;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(sEpi + kLineEpiBytes);
  uint64_t* empty_a = full_a + SA;
  uint64_t* w_full = empty_a + SA;
  uint64_t* tmem_full = w_full + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmB);
    if (!POOL) tma_prefetch_desc(&tmOut);
    for (int i = 0; i < SA; ++i) { mbar_init(&full_a[i], 1); mbar_init(&empty_a[i], 1); }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], POOL ? 16 : 8); }   // arrivals = active epilogue warps
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int lines = p.tiles_h;                               // 112 output lines per image
  // Work list of this CTA.  !POOL: tiles (image, line) round-robin.  POOL: units (image, band) round-robin, each
  // expanding to its 17 (last band: 16) consecutive conv lines.  Every role walks the same list.
  const int num_units = POOL ? p.Nimg * kPoolBands : p.m_tiles;
  auto first_line = [&](int unit, int& n_img, int& h0, int& count) {
    if (POOL) {
      n_img = unit / kPoolBands;
      const int band = unit - n_img * kPoolBands;
      h0 = band * 2 * kPoolBand;
      count = band == kPoolBands - 1 ? 2 * kPoolBand : 2 * kPoolBand + 1;      // conv row 112 does not exist (ceil_mode clips)
    } else {
      n_img = unit / lines; h0 = unit - n_img * lines; count = 1;
    }
  };
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t a_bytes = 4u * 2u * kLinePlane;           // 4 rows x 2 chunks x 115 pixels x 16 B, contiguous in memory
    const uint32_t sA0 = smem_u32(sA), fa0 = smem_u32(full_a), ea0 = smem_u32(empty_a);
    const uint64_t mapB = reinterpret_cast<uint64_t>(&tmB);
    const uint8_t* a_src = reinterpret_cast<const uint8_t*>(p.a_ptr);
    if (leader) {
      bar_expect_tx_u32(smem_u32(w_full), (uint32_t)kLineWBytes);
      tma3d_u32(smem_u32(sW), mapB, smem_u32(w_full), 0, 0, 0);
    }
    int sa = 0; uint32_t pa = 1;
#pragma unroll 1
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      int n_img, h0, count;
      first_line(unit, n_img, h0, count);
#pragma unroll 1
      for (int h = h0; h < h0 + count; ++h) {
        bar_wait_u32(ea0 + sa * 8, pa);
        if (leader) {
          bar_expect_tx_u32(fa0 + sa * 8, a_bytes);
          bulk_load_u32(sA0 + sa * kLineABytes, a_src + ((size_t)n_img * 115 + h) * (2 * kLinePlane), a_bytes, fa0 + sa * 8);
        }
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = p.idesc;
    const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFF) >> 4) | ((uint32_t)(kLinePlane >> 4) << 16);   // LBO = one chunk plane
    const uint32_t w_lo0 = desc_lo(smem_u32(sW));
    const uint32_t fa0 = smem_u32(full_a), ea0 = smem_u32(empty_a);
    const uint32_t tfull0 = smem_u32(tmem_full), tempty0 = smem_u32(tmem_empty);
    bar_wait_u32(smem_u32(w_full), 0);
    int sa = 0; uint32_t pa = 0;
    int local = 0;
#pragma unroll 1
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      int n_img, h0, count;
      first_line(unit, n_img, h0, count);
#pragma unroll 1
      for (int h = h0; h < h0 + count; ++h, ++local) {
        const uint32_t acc = local & 1;
        bar_wait_u32(tempty0 + acc * 8, ((local >> 1) & 1) ^ 1);
        bar_wait_u32(fa0 + sa * 8, pa);
        tc_fence_after();
        if (leader) {
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          const uint32_t a_base = a_lo0 + sa * (kLineABytes >> 4);
#pragma unroll
          for (int kh = 0; kh < 4; ++kh)
#pragma unroll
            for (int kw = 0; kw < 4; ++kw)
              if (!(p.debug & 2) || kh == 0) umma_one(d_tmem, a_base + (uint32_t)(kh * (2 * kLinePlane >> 4) + kw), w_lo0 + (uint32_t)(kh * 4 + kw) * 128u, idesc,
                       (kh | kw) != 0 ? 1u : 0u, kDescHiFlat, kDescHi32);
          commit_u32(ea0 + sa * 8);
          commit_u32(tfull0 + acc * 8);
        }
        if (++sa == SA) { sa = 0; pa ^= 1; }
      }
    }
  } else if (!POOL) {
    epilogue_warps<BLOCK_N, BF16, 0, 2>(p, &tmOut, warp, lane, sEpi, nullptr, tmem_full, tmem_empty, tmem_base, p.m_tiles, blockIdx.x, gridDim.x);
  } else {
    // ---- fused BN + ReLU + max pool epilogue: all 16 warps, thread = output column dw, 16 channels.  (With 8 warps
    // of 32 channels the epilogue ran at 0.86 us per line against 0.64 us of load + MMA: two warps per scheduler
    // cannot hide the latencies of this ~250-instruction dependent chain; ncu source view, profiles/.) ----
    const int ew = warp - 2, quarter = warp & 3, part = ew >> 2;        // part: channels [16*part, 16*part + 16)
    const int dw = quarter * 32 + lane;
    const int et = part * 128 + dw;                          // 0..511: index among the epilogue threads
    const uint32_t stage0 = smem_u32(sEpi);                  // two lines of [128 columns][64 ch] 16-bit, 128 B rows, XOR-swizzled 16 B chunks
    const float4* sc = reinterpret_cast<const float4*>(p.scale + part * 16);
    const float4* sh = reinterpret_cast<const float4*>(p.shift + part * 16);
    uint16_t* out = reinterpret_cast<uint16_t*>(p.out);
    int local = 0, emits = 0;
#pragma unroll 1
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      int n_img, h0, count;
      first_line(unit, n_img, h0, count);
      uint32_t carry[8], m[8];
#pragma unroll 1
      for (int pos = 0; pos < count; ++pos, ++local) {
        const int acc = local & 1;
        mbar_wait(&tmem_full[acc], (local >> 1) & 1);
        tc_fence_after();
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + part * 16, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        uint32_t w[8];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 s4 = __ldg(sc + g), t4 = __ldg(sh + g);
          const float o0 = fmaxf(__uint_as_float(v[4 * g]) * s4.x + t4.x, 0.f), o1 = fmaxf(__uint_as_float(v[4 * g + 1]) * s4.y + t4.y, 0.f);
          const float o2 = fmaxf(__uint_as_float(v[4 * g + 2]) * s4.z + t4.z, 0.f), o3 = fmaxf(__uint_as_float(v[4 * g + 3]) * s4.w + t4.w, 0.f);
          w[2 * g] = pack2<BF16>(o0, o1);
          w[2 * g + 1] = pack2<BF16>(o2, o3);
        }
        const bool last = pos == count - 1;
        bool emit = false;
        if (pos == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) carry[j] = w[j];
        } else if (pos & 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = hmax2_u32<BF16>(carry[j], w[j]);
          emit = last;                                         // rows 110, 111 only: the bottom pooled row of the image
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { m[j] = hmax2_u32<BF16>(m[j], w[j]); carry[j] = w[j]; }
          emit = true;
        }
        if (emit) {                                            // CTA-uniform: pos and count are
          const int prow = (h0 >> 1) + ((pos - 1) >> 1);       // pooled row
          const uint32_t buf = stage0 + (uint32_t)(emits & 1) * (128 * 128);
          ++emits;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t chunk = (uint32_t)(part * 2 + g);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + dw * 128 + ((chunk ^ (uint32_t)(dw & 7)) << 4)),
                         "r"(m[4 * g]), "r"(m[4 * g + 1]), "r"(m[4 * g + 2]), "r"(m[4 * g + 3]) : "memory");
          }
          named_bar_sync(1, 512);
          // horizontal 3-wide / stride-2 maximum: item = (pooled column j, 16-byte channel chunk c)
          if (et < 56 * 8) {
            const int j = et >> 3, c = et & 7;
            uint32_t r0[4], r1[4];
            {
              const int x = 2 * j;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3])
                           : "r"(buf + x * 128 + (((uint32_t)c ^ (uint32_t)(x & 7)) << 4)));
            }
#pragma unroll
            for (int d = 1; d < 3; ++d) {
              const int x = 2 * j + d;
              if (x < 112) {
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3])
                             : "r"(buf + x * 128 + (((uint32_t)c ^ (uint32_t)(x & 7)) << 4)));
#pragma unroll
                for (int q = 0; q < 4; ++q) r0[q] = hmax2_u32<BF16>(r0[q], r1[q]);
              }
            }
            *reinterpret_cast<uint4*>(out + ((((size_t)n_img * 56 + prow) * 56 + j) * 64 + c * 8)) = make_uint4(r0[0], r0[1], r0[2], r0[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int encode_map(CUtensorMap* map, ElemType elem, int rank, const void* base, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr,
                      CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_fn();
  MM_REQUIRE(fn, MIMAMO_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bd[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bd[i] = box[i]; es[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(map, elem == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                  const_cast<void*>(base), gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MM_REQUIRE(r == CUDA_SUCCESS, MIMAMO_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, box %u %u %u %u)",
             (int)r, rank, box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return MIMAMO_OK;
}

static uint32_t make_idesc(int block_n, ElemType elem) {
  // cute::UMMA::InstrDescriptor: c_format F32 (1) @ [4,6); a/b format @ [7,10)/[10,13) (F16 = 0,
  // BF16 = 1); a/b K-major (0) @ 15/16; N>>3 @ [17,23); M>>4 @ [24,29).
  const uint32_t fmt = elem == kBF16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}

static uint32_t make_idesc2(int block_n, ElemType elem) {    // cta_group::2: M = 256 across the CTA pair
  const uint32_t fmt = elem == kBF16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

int out_size(int in, int k, int s, int p) { return (in + 2 * p - k) / s + 1; }

static int num_sms() {
  static std::atomic<int> cache[64];
  const int dev = current_device();
  int n = cache[dev & 63].load(std::memory_order_relaxed);
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

// optional per-launch CUDA-event timing of the GEMM kernel (bench.py's roofline leg)
static bool g_profile = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static size_t g_prof_used = 0;
static double g_prof_flops = 0.0;

template <int BLOCK_N, bool BF16, int RES>
static int launch_cfg(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const ConvParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, RES>;
  static DeviceOnce attr_set;
  if (attr_set.need()) {
    MM_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, BF16, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.mark();
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile) {
    if (g_prof_used == g_prof_events.size()) {
      cudaEvent_t a0, a1;
      MM_CUDA(cudaEventCreate(&a0));
      MM_CUDA(cudaEventCreate(&a1));
      g_prof_events.emplace_back(a0, a1);
    }
    e0 = g_prof_events[g_prof_used].first; e1 = g_prof_events[g_prof_used].second;
    ++g_prof_used;
    g_prof_flops += 2.0 * (double)p.m_tiles * kBlockM * (double)p.n_tiles * BLOCK_N * (double)p.num_k_blocks * kBlockK;
    MM_CUDA(cudaEventRecord(e0, stream));
  }
  if (p.pair) {
    cudaLaunchConfig_t cfg = {};
    static std::atomic<int> pairs_cache[64];
    int max_pairs = pairs_cache[current_device() & 63].load(std::memory_order_relaxed);
    if (!max_pairs) {
      max_pairs = max_pairs_for(conv_gemm_kernel<BLOCK_N, BF16, RES>, Cfg::kSmemBytes);
      pairs_cache[current_device() & 63].store(max_pairs, std::memory_order_relaxed);
    }
    const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
    cfg.gridDim = dim3((unsigned)(2 * (pairs < max_pairs ? pairs : max_pairs)), 1, 1);   // whole, co-resident clusters only
    cfg.blockDim = dim3(kGemmThreads, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    MM_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BLOCK_N, BF16, RES>, a, b, o, p));
  } else {
    conv_gemm_kernel<BLOCK_N, BF16, RES><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(a, b, o, p);
  }
  MM_LAUNCH_OK();
  if (e1) MM_CUDA(cudaEventRecord(e1, stream));
  return MIMAMO_OK;
}

// Persistent 2-CTA-cluster kernels must not launch more clusters than can be co-resident (GPCs with an odd number of
// free SMs cannot host a pair): ask the occupancy calculator once per kernel.
template <typename Kernel>
static int max_pairs_for(Kernel kernel, size_t smem_bytes) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(num_sms() & ~1), 1, 1);
  cfg.blockDim = dim3(kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = num_sms() / 2; }
  const int cap = num_sms() / 2;
  if (getenv("MIMAMO_VERBOSE")) fprintf(stderr, "mimamo: %d co-resident 2-CTA clusters (of %d SM pairs), %zu B of shared memory per CTA\n", n, cap, smem_bytes);
  return n < cap ? n : cap;
}

// cta_group::2 launch (256-wide tiles only): clusters of two CTAs, M = 256 instruction descriptor
template <bool BF16, int RES>
static int launch_cfg2(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const ConvParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<256, RES>;
  static DeviceOnce attr_set;
  if (attr_set.need()) {
    MM_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel<256, BF16, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.mark();
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile) {
    if (g_prof_used == g_prof_events.size()) {
      cudaEvent_t a0, a1;
      MM_CUDA(cudaEventCreate(&a0));
      MM_CUDA(cudaEventCreate(&a1));
      g_prof_events.emplace_back(a0, a1);
    }
    e0 = g_prof_events[g_prof_used].first; e1 = g_prof_events[g_prof_used].second;
    ++g_prof_used;
    g_prof_flops += 2.0 * (double)p.m_tiles * kBlockM * (double)p.n_tiles * 256 * (double)p.num_k_blocks * kBlockK;
    MM_CUDA(cudaEventRecord(e0, stream));
  }
  cudaLaunchConfig_t cfg = {};
  const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
  static std::atomic<int> pairs_cache[64];
  int max_pairs = pairs_cache[current_device() & 63].load(std::memory_order_relaxed);
  if (!max_pairs) {
    max_pairs = max_pairs_for(conv_gemm2_kernel<256, BF16, RES>, Cfg::kSmemBytes);
    pairs_cache[current_device() & 63].store(max_pairs, std::memory_order_relaxed);
  }
  cfg.gridDim = dim3((unsigned)(2 * (pairs < max_pairs ? pairs : max_pairs)), 1, 1);
  cfg.blockDim = dim3(kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  MM_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm2_kernel<256, BF16, RES>, a, b, o, p));
  count_launch();
  if (e1) MM_CUDA(cudaEventRecord(e1, stream));
  return MIMAMO_OK;
}

template <int BLOCK_N>
static int launch_n(bool bf, bool res, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const ConvParams& p, cudaStream_t s) {
  if (BLOCK_N == 256 && res && p.res_tma == 2)
    return bf ? launch_cfg<BLOCK_N == 256 ? 256 : BLOCK_N, true, BLOCK_N == 256 ? 2 : 1>(a, b, o, p, s)
              : launch_cfg<BLOCK_N == 256 ? 256 : BLOCK_N, false, BLOCK_N == 256 ? 2 : 1>(a, b, o, p, s);
  if (bf) return res ? launch_cfg<BLOCK_N, true, 1>(a, b, o, p, s) : launch_cfg<BLOCK_N, true, 0>(a, b, o, p, s);
  return res ? launch_cfg<BLOCK_N, false, 1>(a, b, o, p, s) : launch_cfg<BLOCK_N, false, 0>(a, b, o, p, s);
}

// BLOCK_N actually launched: residual layers use at most 128 columns (the residual ring takes the
// shared memory of two 256-wide pipeline stages).
static int effective_block_n(const ConvLayer& L, bool has_res) {
  const char* e = getenv("MIMAMO_RES_BLOCK_N");        // read per call: the switches below are experiment / test knobs, never cached
  const int res_bn = (e && atoi(e) == 128) ? 128 : 256;      // 256: each tile writes whole 512-byte pixel rows (measured 33.3 -> 32.4 ms per 2048 images)
  return (has_res && L.block_n > res_bn) ? res_bn : L.block_n;
}

// Output tensor map for the epilogue's TMA store: the 16-bit NHWC output viewed as [rows][Cout] (flat) or
// [B][Ho][Wo][Cout] (spatial), one box = one column group (32 or 64 channels) of one tile's pixels.
static int out_cols(int block_n) { return block_n >= 256 ? 64 : 32; }
static int out_map_flat(CUtensorMap* map, ElemType elem, void* out, int ldc, int cout, long long rows, int block_n, int box_rows) {
  const uint64_t dims[2] = {(uint64_t)cout, (uint64_t)rows};
  const uint64_t strides[1] = {(uint64_t)ldc * 2};
  const uint32_t box[2] = {(uint32_t)out_cols(block_n), (uint32_t)box_rows};
  const uint32_t es[2] = {1, 1};
  return encode_map(map, elem, 2, out, dims, strides, box, es, out_cols(block_n) == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}
static int out_map_spatial(CUtensorMap* map, ElemType elem, void* out, int ldc, int cout, int Wo, int Ho, int B, int bw, int bh,
                           int bn, int block_n) {
  const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)ldc * 2, (uint64_t)Wo * ldc * 2, (uint64_t)Ho * Wo * ldc * 2};
  const uint32_t box[4] = {(uint32_t)out_cols(block_n), (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
  const uint32_t es[4] = {1, 1, 1, 1};
  return encode_map(map, elem, 4, out, dims, strides, box, es, out_cols(block_n) == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

// MIMAMO_STORE_MODE = "<flat><strided 1x1><3x3>" digits (default "012") selects the epilogue store granularity per layer kind
static int store_mode_setting(int kind) {
  int modes[3];
  const char* e = getenv("MIMAMO_STORE_MODE");
  const char* d = (e && strlen(e) == 3) ? e : "012";
  for (int i = 0; i < 3; ++i) modes[i] = (d[i] >= '0' && d[i] <= '2') ? d[i] - '0' : i;
  if (modes[1] == 0) modes[1] = 1;                            // per-warp boxes only exist for flat layers
  if (modes[2] == 0) modes[2] = 1;
  return modes[kind];
}

static int launch(const ConvLayer& L, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const ConvParams& p, cudaStream_t s) {
  const bool bf = L.elem == kBF16, res = p.residual != nullptr;
  if (p.pair == 2) {
    if (res && p.res_tma == 2) return bf ? launch_cfg2<true, 2>(a, b, o, p, s) : launch_cfg2<false, 2>(a, b, o, p, s);
    if (bf) return res ? launch_cfg2<true, 1>(a, b, o, p, s) : launch_cfg2<true, 0>(a, b, o, p, s);
    return res ? launch_cfg2<false, 1>(a, b, o, p, s) : launch_cfg2<false, 0>(a, b, o, p, s);
  }
  switch (effective_block_n(L, res)) {
    case 64:  return launch_n<64>(bf, res, a, b, o, p, s);
    case 128: return launch_n<128>(bf, res, a, b, o, p, s);
    case 256: return launch_n<256>(bf, res, a, b, o, p, s);
  }
  set_error("unsupported BLOCK_N %d", L.block_n);
  return MIMAMO_E_RUNTIME;
}

template <int BLOCK_N, bool BF16>
static int launch_halo_cfg(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const ConvParams& p, cudaStream_t stream) {
  using Cfg = HaloCfg<BLOCK_N>;
  static DeviceOnce attr_set;
  if (attr_set.need()) {
    MM_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<BLOCK_N, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set.mark();
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile) {
    if (g_prof_used == g_prof_events.size()) {
      cudaEvent_t a0, a1;
      MM_CUDA(cudaEventCreate(&a0));
      MM_CUDA(cudaEventCreate(&a1));
      g_prof_events.emplace_back(a0, a1);
    }
    e0 = g_prof_events[g_prof_used].first; e1 = g_prof_events[g_prof_used].second;
    ++g_prof_used;
    g_prof_flops += 2.0 * (double)p.m_tiles * kBlockM * (double)p.n_tiles * BLOCK_N * (double)p.num_k_blocks * kBlockK;
    MM_CUDA(cudaEventRecord(e0, stream));
  }
  conv3x3_halo_kernel<BLOCK_N, BF16><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(a, b, o, p);
  MM_LAUNCH_OK();
  if (e1) MM_CUDA(cudaEventRecord(e1, stream));
  return MIMAMO_OK;
}

// MIMAMO_CONV3X3_HALO: 0 = box-per-tap only; 1 = halo mode (default).  Round-1 experiment on B200: setting the
// descriptor's base_offset field to (addr >> 7) & 7 for the row-shifted views gives WRONG results; the 128B
// swizzle is a function of the absolute shared-memory address, so base_offset stays 0.
static int halo_setting() {
  const char* e = getenv("MIMAMO_CONV3X3_HALO");
  return e ? atoi(e) : 1;
}

// ---- 16-bit rounding of the weights ----
static inline float f16_bits_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else { float f = ldexpf((float)man, -24); memcpy(&bits, &f, 4); bits |= sign; }
  } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp + 112u) << 23) | (man << 13);
  float f; memcpy(&f, &bits, 4); return f;
}
static inline float w16_to_float(uint16_t b, ElemType elem) {
  if (elem == kBF16) { const uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
  return f16_bits_to_float(b);
}
static inline uint16_t float_to_w16(float v, ElemType elem) {
  uint16_t bits;
  if (elem == kBF16) { __nv_bfloat16 h = __float2bfloat16(v); memcpy(&bits, &h, 2); }
  else { __half h = __float2half(v); memcpy(&bits, &h, 2); }
  return bits;
}
// the representable neighbour of a finite 16-bit value towards -inf (down) or +inf
static inline uint16_t w16_neighbor(uint16_t b, bool down) {
  const bool neg = (b & 0x8000u) != 0;
  if ((b & 0x7FFFu) == 0) return down ? 0x8001u : 0x0001u;
  return (uint16_t)((neg == down) ? b + 1 : b - 1);          // larger magnitude when moving away from zero
}

static void quantize_rows(const float* w, uint16_t* out, int row0, int row1, size_t K, int cin_p, ElemType elem, int mode, const float* mu) {
  std::vector<uint16_t> alt(K);
  std::vector<double> delta(K);
  std::vector<float> cost(K);
  std::vector<int> order;
  for (int o = row0; o < row1; ++o) {
    const float* wr = w + (size_t)o * K;
    uint16_t* q = out + (size_t)o * K;
    double tot = 0.0;
    order.clear();
    for (size_t k = 0; k < K; ++k) {
      q[k] = float_to_w16(wr[k], elem);
      if (mode == 0 || wr[k] == 0.f) continue;               // zero weights (channel padding) stay zero
      const double m = mu ? (double)mu[k % cin_p] : 1.0;
      const double r = (double)w16_to_float(q[k], elem) - (double)wr[k];
      tot += r * m;
      if (r == 0.0 || m == 0.0) continue;
      alt[k] = w16_neighbor(q[k], r > 0.0);
      const float av = w16_to_float(alt[k], elem);
      if (!(fabsf(av) < 6.0e4f)) continue;
      const double ar = (double)av - (double)wr[k];
      delta[k] = (ar - r) * m;
      cost[k] = (float)(fabs(ar) - fabs(r));
      order.push_back((int)k);
    }
    if (mode == 0) continue;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] < cost[b] || (cost[a] == cost[b] && a < b); });
    for (int k : order) {                                      // cheapest flips first; take those that shrink the weighted residual sum
      if (fabs(tot + delta[k]) < fabs(tot)) { tot += delta[k]; q[k] = alt[k]; }
    }
  }
}

int conv_layer_quantize(ConvLayer& L, int mode, const float* mu) {
  const size_t K = (size_t)L.ksize * L.ksize * L.Cin_p;
  MM_REQUIRE(L.w_f32.size() == (size_t)L.Cout * K, MIMAMO_E_VALUE, "conv_layer_quantize: layer holds no fp32 weights");
  std::vector<uint16_t> packed((size_t)L.Cout * K);
  unsigned nthr = std::thread::hardware_concurrency();
  if (nthr < 1) nthr = 1;
  if (nthr > 32) nthr = 32;
  if (mode == 0 || (size_t)L.Cout * K < (1u << 16)) nthr = 1;
  std::vector<std::thread> pool;
  const int per = (L.Cout + (int)nthr - 1) / (int)nthr;
  for (unsigned t = 0; t < nthr; ++t) {
    const int r0 = (int)t * per, r1 = r0 + per < L.Cout ? r0 + per : L.Cout;
    if (r0 >= r1) break;
    if (nthr == 1) quantize_rows(L.w_f32.data(), packed.data(), r0, r1, K, L.Cin_p, L.elem, mode, mu);
    else pool.emplace_back(quantize_rows, L.w_f32.data(), packed.data(), r0, r1, K, L.Cin_p, L.elem, mode, mu);
  }
  for (auto& th : pool) th.join();
  if (!L.w_dev) MM_CUDA(cudaMalloc(&L.w_dev, packed.size() * 2));
  MM_CUDA(cudaMemcpy(L.w_dev, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
  return MIMAMO_OK;
}

int conv_layer_init(ConvLayer& L, const float* w_host, const float* scale_host, const float* shift_host, int Cout,
                    int Cin, int ksize, int stride, int pad, int relu, ElemType elem) {
  MM_REQUIRE(Cout % 64 == 0, MIMAMO_E_RUNTIME, "Cout=%d must be a multiple of 64 for the tcgen05 engine", Cout);
  L.Cin = Cin; L.Cin_p = (Cin + 63) / 64 * 64; L.Cout = Cout;
  L.ksize = ksize; L.stride = stride; L.pad = pad; L.relu = relu; L.elem = elem;
  L.block_n = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
  const int taps = ksize * ksize;
  const size_t K = (size_t)taps * L.Cin_p;
  L.w_f32.assign((size_t)Cout * K, 0.f);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < taps; ++t)
        L.w_f32[(size_t)o * K + (size_t)t * L.Cin_p + c] = w_host[((size_t)o * Cin + c) * taps + t];
  int rc = conv_layer_quantize(L, 0, nullptr);
  if (rc) return rc;
  rc = upload(&L.scale_dev, scale_host, (size_t)Cout);
  if (rc) return rc;
  return upload(&L.shift_dev, shift_host, (size_t)Cout);
}

void conv_layer_free(ConvLayer& L) {
  cudaFree(L.w_dev); cudaFree(L.scale_dev); cudaFree(L.shift_dev);
  L.w_dev = nullptr; L.scale_dev = L.shift_dev = nullptr;
  std::vector<float>().swap(L.w_f32);
}

// Pair mode (2-CTA clusters, multicast weight boxes) pays where the weight tile dominates the bytes a tile pulls
// through L2 -> SM: 256-wide tiles with K >= 256 and enough M tiles to keep every cluster busy (measured: -3 % on the stage-4
// 3x3 and increase layers; those layers turned out not to be weight-traffic-bound).  MIMAMO_PAIR=0 disables, 2 forces (tests).
// Returns 0 (single CTAs), 1 (pair mode, multicast weights) or 2 (cta_group::2 UMMA) for a 256-wide layer.
// Measured per 2048 images (profiles/layers_r1_pair_modes.txt; modes 0 / 1 / 2): stage-4 reduce 212 / 215 / 184 us,
// stage-4 3x3 417 / 412 / 367, stage-5 reduce 184 / 184 / 152, stage-5 3x3 398 / 392 / 357, stage-5 proj 433 / 413 / 357,
// stage-5 increase 283 / 275 / 247 -- but stage-4 increase 375 / 365 / 420 and the strided stage-4 proj 492 / 482 / 507:
// with few K blocks per tile the layer is epilogue-bound and coupling the two CTAs' epilogues to one MMA issuer costs
// more than the halved weight ingress saves.  Hence: cta_group::2 from 16 K blocks (8 on flat layers), multicast pairs
// from 4.  MIMAMO_PAIR = "<mode><force>": first digit caps the mode (default 2), a second digit 1 forces exactly that
// mode on every 256-wide layer (tests).
static int pair_wanted(int block_n, int num_k_blocks, int m_tiles, bool flat) {
  const char* e = getenv("MIMAMO_PAIR");
  const int cap = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
  const bool force = e && e[0] && e[1] == '1';
  if (cap == 0 || block_n != 256) return 0;
  if (force) return cap;
  if (m_tiles < num_sms() || num_k_blocks < 4) return 0;
  const int want = (num_k_blocks >= 16 || (flat && num_k_blocks >= 8)) ? 2 : 1;
  return want < cap ? want : cap;
}

static int weight_map(const ConvLayer& L, CUtensorMap* map, int block_n) {
  const uint64_t K = (uint64_t)L.ksize * L.ksize * L.Cin_p;
  const uint64_t dims[2] = {K, (uint64_t)L.Cout};
  const uint64_t strides[1] = {K * 2};
  const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)block_n};
  const uint32_t es[2] = {1, 1};
  return encode_map(map, L.elem, 2, L.w_dev, dims, strides, box, es);
}

static int fill_common(ConvParams& p, const ConvLayer& L, void* out, int ldc, const void* residual, int ld_res) {
  const int bn = effective_block_n(L, residual != nullptr);
  p.idesc = make_idesc(bn, L.elem);
  p.scale = L.scale_dev; p.shift = L.shift_dev;
  p.residual = residual; p.out = out; p.ldc = ldc; p.ld_res = ld_res; p.relu = L.relu;
  { const char* e = getenv("MIMAMO_DEBUG"); p.debug = e ? atoi(e) : 0; }
  p.n_tiles = L.Cout / bn;
  p.cin_blocks = L.Cin_p / kBlockK;
  p.taps_w = L.ksize;
  p.num_k_blocks = L.ksize * L.ksize * p.cin_blocks;
  p.stride = L.stride; p.pad = L.pad;
  return bn;
}

int gemm_forward(const ConvLayer& L, const void* a, int M, void* out, int ldc, const void* residual, int ld_res,
                 cudaStream_t stream) {
  MM_REQUIRE(L.ksize == 1 && L.stride == 1 && L.pad == 0, MIMAMO_E_VALUE, "gemm_forward needs a 1x1 stride-1 layer");
  MM_REQUIRE(ldc % 8 == 0 && (residual == nullptr || ld_res % 8 == 0), MIMAMO_E_VALUE, "row pitches must be multiples of 8");
  if (M == 0) return MIMAMO_OK;
  CUtensorMap ma, mb;
  const uint64_t dims[2] = {(uint64_t)L.Cin_p, (uint64_t)M};
  const uint64_t strides[1] = {(uint64_t)L.Cin_p * 2};
  const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)kBlockM};
  const uint32_t es[2] = {1, 1};
  int rc = encode_map(&ma, L.elem, 2, a, dims, strides, box, es);
  if (rc) return rc;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  const int bn = fill_common(p, L, out, ldc, residual, ld_res);
  p.mode = 0; p.M_total = M; p.a_rows = kBlockM;
  p.m_tiles = (M + kBlockM - 1) / kBlockM;
  p.pair = pair_wanted(bn, p.num_k_blocks, p.m_tiles, true);
  if (p.pair == 2) p.idesc = make_idesc2(bn, L.elem);
  rc = weight_map(L, &mb, p.pair ? bn / 2 : bn);              // pair mode: each CTA fetches half of the N rows of a weight box
  if (rc) return rc;
  CUtensorMap mo;
  p.store_mode = store_mode_setting(0);
  rc = out_map_flat(&mo, L.elem, out, ldc, L.Cout, M, effective_block_n(L, residual != nullptr), p.store_mode == 0 ? 32 : kBlockM);
  if (rc) return rc;
  {
    // residual through TMA into the staging tile.  MIMAMO_RES_TMA: 0 = the per-lane cp.async ring instead, 1 = two staging
    // buffers (the ring's memory is the second), 2 = one buffer + a deeper operand pipeline; default: 2 where a tile has
    // at least four K blocks (the pipeline depth matters there), 1 otherwise
    const char* e = getenv("MIMAMO_RES_TMA");
    const int want = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : (p.num_k_blocks >= 4 ? 2 : 1);
    if (residual != nullptr && bn == 256 && p.store_mode == 0 && want != 0) {
      rc = out_map_flat(&p.res_map, L.elem, const_cast<void*>(residual), ld_res, L.Cout, M, bn, 32);
      if (rc) return rc;
      p.res_tma = want;
    }
  }
  return launch(L, ma, mb, mo, p, stream);
}

// MIMAMO_CHAIN=0 runs the two layers of a chain as separate launches (cross-check / A-B measurements)
bool chain_enabled() {
  const char* e = getenv("MIMAMO_CHAIN");
  return !(e && e[0] == '0');
}

bool chain_supported(const ConvLayer& L1, const ConvLayer& L2) {
  return L1.ksize == 1 && L1.stride == 1 && L1.pad == 0 && L2.ksize == 1 && L2.stride == 1 && L2.pad == 0 && L1.elem == L2.elem &&
         L1.Cout % 128 == 0 && L2.Cin_p == L1.Cout && L2.Cout <= 128 && L2.Cout % 64 == 0;
}

// out1 = act(L1(a) + residual) [M][L1.Cout] (dense rows) and out2 = act(L2(out1)) [M][L2.Cout] in one launch.
int chain_forward(const ConvLayer& L1, const ConvLayer& L2, const void* a, int M, void* out1, const void* residual, int ld_res,
                  void* out2, int ldc2, cudaStream_t stream) {
  MM_REQUIRE(chain_supported(L1, L2), MIMAMO_E_VALUE, "chain_forward: unsupported layer pair (%d -> %d -> %d)", L1.Cin_p, L1.Cout, L2.Cout);
  MM_REQUIRE(residual != nullptr && ld_res % 8 == 0 && ldc2 % 8 == 0, MIMAMO_E_VALUE, "chain_forward needs a residual and row pitches that are multiples of 8");
  if (M == 0) return MIMAMO_OK;
  constexpr int BN = 128;
  ChainParams p;
  memset(&p, 0, sizeof(p));
  p.m_tiles = (M + kBlockM - 1) / kBlockM;
  p.n1_tiles = L1.Cout / BN; p.nkb1 = L1.Cin_p / kBlockK;
  p.n2_cols = L2.Cout; p.nkb2 = L2.Cin_p / kBlockK;
  p.idesc1 = make_idesc(BN, L1.elem); p.idesc2 = make_idesc(L2.Cout, L2.elem);
  p.relu1 = L1.relu; p.relu2 = L2.relu;
  p.scale1 = L1.scale_dev; p.shift1 = L1.shift_dev; p.scale2 = L2.scale_dev; p.shift2 = L2.shift_dev;
  // on-chip hand-over with resident weights where both matrices fit (ResNet50 stage 2: 32 + 32 KB); MIMAMO_CHAIN_SMEM=0 forces
  // the hand-over through L2 (A-B measurements, cross-check)
  { const char* e = getenv("MIMAMO_CHAIN_DEBUG"); p.debug = e ? atoi(e) : 0; }
  const char* rwe = getenv("MIMAMO_CHAIN_SMEM");
  const bool rw = p.nkb1 == 1 && L1.Cout == ChainSmemCfg::N1 && L2.Cout == ChainSmemCfg::N2 && !(rwe && rwe[0] == '0');
  // ... and with streamed weights for the stage-3 shape
  const bool st = L1.Cin_p == ChainStreamCfg::K1 && L1.Cout == ChainStreamCfg::N1 && L2.Cout == ChainStreamCfg::N2 && !(rwe && rwe[0] == '0');
  const uint32_t es[2] = {1, 1};
  const uint32_t abox[2] = {(uint32_t)kBlockK, (uint32_t)kBlockM};
  {
    const uint64_t dims[2] = {(uint64_t)L1.Cin_p, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)L1.Cin_p * 2};
    int rc = encode_map(&p.a1, L1.elem, 2, a, dims, strides, abox, es);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)L1.Cout, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)L1.Cout * 2};
    int rc = encode_map(&p.a2, L1.elem, 2, out1, dims, strides, abox, es);
    if (rc) return rc;
  }
  int rc = weight_map(L1, &p.b1, BN);
  if (!rc) rc = weight_map(L2, &p.b2, L2.Cout);
  // store / residual boxes: 32 rows x 32 columns per warp (L2 hand-over) or x 64 columns per warp pair (on-chip hand-over)
  const int box_bn = (rw || st) ? 256 : BN;
  if (!rc) rc = out_map_flat(&p.out1, L1.elem, out1, L1.Cout, L1.Cout, M, box_bn, 32);
  if (!rc) rc = out_map_flat(&p.res, L1.elem, const_cast<void*>(residual), ld_res, L1.Cout, M, box_bn, 32);
  if (!rc) rc = out_map_flat(&p.out2, L2.elem, out2, ldc2, L2.Cout, M, box_bn, 32);
  if (rc) return rc;
  const bool bf = L1.elem == kBF16;
  static DeviceOnce attr_set;
  if (attr_set.need()) {
    MM_CUDA(cudaFuncSetAttribute(conv_chain_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg<BN>::kSmemBytes));
    MM_CUDA(cudaFuncSetAttribute(conv_chain_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg<BN>::kSmemBytes));
    MM_CUDA(cudaFuncSetAttribute(conv_chain_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainSmemCfg::kSmemBytes));
    MM_CUDA(cudaFuncSetAttribute(conv_chain_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainSmemCfg::kSmemBytes));
    MM_CUDA(cudaFuncSetAttribute(conv_chain_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainStreamCfg::kSmemBytes));
    MM_CUDA(cudaFuncSetAttribute(conv_chain_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainStreamCfg::kSmemBytes));
    attr_set.mark();
  }
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile) {
    if (g_prof_used == g_prof_events.size()) {
      cudaEvent_t a0, a1;
      MM_CUDA(cudaEventCreate(&a0));
      MM_CUDA(cudaEventCreate(&a1));
      g_prof_events.emplace_back(a0, a1);
    }
    e0 = g_prof_events[g_prof_used].first; e1 = g_prof_events[g_prof_used].second;
    ++g_prof_used;
    g_prof_flops += 2.0 * (double)p.m_tiles * kBlockM * ((double)L1.Cout * L1.Cin_p + (double)L2.Cout * L2.Cin_p);
    MM_CUDA(cudaEventRecord(e0, stream));
  }
  if (rw) {
    if (bf) conv_chain_smem_kernel<true><<<grid, kGemmThreads, ChainSmemCfg::kSmemBytes, stream>>>(p);
    else conv_chain_smem_kernel<false><<<grid, kGemmThreads, ChainSmemCfg::kSmemBytes, stream>>>(p);
  } else if (st) {
    if (bf) conv_chain_stream_kernel<true><<<grid, kGemmThreads, ChainStreamCfg::kSmemBytes, stream>>>(p);
    else conv_chain_stream_kernel<false><<<grid, kGemmThreads, ChainStreamCfg::kSmemBytes, stream>>>(p);
  } else {
    if (bf) conv_chain_kernel<BN, true><<<grid, kGemmThreads, ChainCfg<BN>::kSmemBytes, stream>>>(p);
    else conv_chain_kernel<BN, false><<<grid, kGemmThreads, ChainCfg<BN>::kSmemBytes, stream>>>(p);
  }
  MM_LAUNCH_OK();
  if (e1) MM_CUDA(cudaEventRecord(e1, stream));
  return MIMAMO_OK;
}

int conv_forward(const ConvLayer& L, const void* x, int B, int H, int W, void* out, int ldc, const void* residual,
                 int ld_res, cudaStream_t stream, int in_pitch, int in_channels) {
  if (in_pitch == 0) in_pitch = L.Cin_p;
  if (in_channels == 0) in_channels = L.Cin_p;
  MM_REQUIRE(in_pitch % 8 == 0 && in_channels >= 1 && in_channels <= L.Cin_p && in_channels <= in_pitch, MIMAMO_E_VALUE,
             "bad input pitch / channel count (%d / %d for %d padded channels)", in_pitch, in_channels, L.Cin_p);
  if (L.ksize == 1 && L.stride == 1 && L.pad == 0) {
    MM_REQUIRE(in_pitch == L.Cin_p && in_channels == L.Cin_p, MIMAMO_E_VALUE, "flat 1x1 layers need the dense Cin_p input layout");
    return gemm_forward(L, x, B * H * W, out, ldc, residual, ld_res, stream);
  }
  MM_REQUIRE(ldc % 8 == 0 && (residual == nullptr || ld_res % 8 == 0), MIMAMO_E_VALUE, "row pitches must be multiples of 8");
  if (B == 0) return MIMAMO_OK;
  int Ho = out_size(H, L.ksize, L.stride, L.pad), Wo = out_size(W, L.ksize, L.stride, L.pad);
  // fill-bound 3x3 layers (Cout <= 128): halo-resident kernel
  if (halo_setting() != 0 && L.ksize == 3 && L.stride == 1 && L.pad == 1 && residual == nullptr && L.block_n <= 128 &&
      W + 2 <= 64) {
    const int line = W + 2;
    int bh = kBlockM / line;
    if (bh > H) bh = H;
    while ((bh + 2) * line > kHaloABytes / 128) --bh;
    if (bh >= 1) {
      CUtensorMap ma, mb;
      const uint64_t dims[4] = {(uint64_t)in_channels, (uint64_t)W, (uint64_t)H, (uint64_t)B};
      const uint64_t strides[3] = {(uint64_t)in_pitch * 2, (uint64_t)W * in_pitch * 2, (uint64_t)H * W * in_pitch * 2};
      const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)line, (uint32_t)(bh + 2), 1};
      const uint32_t es[4] = {1, 1, 1, 1};
      int rc = encode_map(&ma, L.elem, 4, x, dims, strides, box, es);
      if (rc) return rc;
      ConvParams p;
      memset(&p, 0, sizeof(p));
      const int bn_cols = fill_common(p, L, out, ldc, nullptr, 0);
      {
        // weights viewed as [tap][Cout][Cin_p]: one box = the three taps of a kernel row, each tap's
        // [BLOCK_N][64] tile contiguous in shared memory
        const uint64_t K = (uint64_t)9 * L.Cin_p;
        const uint64_t wdims[3] = {(uint64_t)L.Cin_p, (uint64_t)L.Cout, 9};
        const uint64_t wstr[2] = {K * 2, (uint64_t)L.Cin_p * 2};
        const uint32_t wbox[3] = {(uint32_t)kBlockK, (uint32_t)bn_cols, 3};
        const uint32_t wes[3] = {1, 1, 1};
        rc = encode_map(&mb, L.elem, 3, L.w_dev, wdims, wstr, wbox, wes);
        if (rc) return rc;
      }
      p.mode = 2;
      p.Wo = Wo; p.Ho = Ho; p.Nimg = B;
      p.bw = line; p.bh = bh; p.bn = 1;
      p.tiles_w = 1; p.tiles_h = (Ho + bh - 1) / bh;
      p.a_rows = (bh + 2) * line;
      p.m_tiles = p.tiles_h * B;
      {
        const char* e = getenv("MIMAMO_HALO_RESIDENT_W");
        p.resident_w = (p.cin_blocks == 1 && p.n_tiles == 1 && !(e && e[0] == '0')) ? 1 : 0;
      }
      const bool bf = L.elem == kBF16;
      p.store_mode = store_mode_setting(2);
      CUtensorMap mo;                                          // one box = bh padded lines; the two extra columns per line fall outside Wo and are clipped
      rc = out_map_spatial(&mo, L.elem, out, ldc, L.Cout, Wo, Ho, B, line, bh, 1, bn_cols);
      if (rc) return rc;
      if (L.block_n == 64) return bf ? launch_halo_cfg<64, true>(ma, mb, mo, p, stream) : launch_halo_cfg<64, false>(ma, mb, mo, p, stream);
      return bf ? launch_halo_cfg<128, true>(ma, mb, mo, p, stream) : launch_halo_cfg<128, false>(ma, mb, mo, p, stream);
    }
  }
  // Strided 1x1 layers (ResNet50's stride-2 `_reduce` / `_proj`): the layer only reads every stride-th pixel of every
  // stride-th row, which is a DENSE box over a strided VIEW of the tensor (global strides multiplied by the stride) --
  // TMA element strides (traversal strides) make the unit walk the skipped pixels as well (measured: those layers ran at
  // 61 % of their floor).  MIMAMO_STRIDED_VIEW=0 restores the element-stride boxes (cross-check).
  int vstride = L.stride;                       // stride still expressed through the box / element strides
  int Wv = W, Hv = H;                           // extent of the tensor the map describes
  uint64_t pix_pitch = (uint64_t)in_pitch;      // elements between horizontally adjacent pixels of that tensor
  {
    const char* e = getenv("MIMAMO_STRIDED_VIEW");
    if (L.ksize == 1 && L.pad == 0 && L.stride > 1 && !(e && e[0] == '0')) {
      vstride = 1; Wv = Wo; Hv = Ho; pix_pitch = (uint64_t)in_pitch * L.stride;
      // When H is a multiple of the stride the sampled rows of consecutive images are evenly spaced as well, so (row, image)
      // merge into ONE dimension and a tile's rows may run across images: 14x14 / 7x7 outputs fill 126 of the 128 MMA rows
      // instead of 98 (9 + 5 row boxes per image).  The dense output tensor merges the same way.
      if (H == Ho * L.stride && residual == nullptr) { Hv = Ho * B; Ho = Hv; B = 1; }
    }
  }
  // Stride-2 k x k layers (PhaseNet's second convolution of every block, api/mimamo_net.py:72): input pixel 2y + kh - pad lies
  // on the sub-lattice of row parity (kh - pad) & 1 at index y + ((kh - pad) >> 1), so every tap is a DENSE box over one of
  // four strided views (row parity x column parity) instead of an element-strided box that walks the skipped pixels (those
  // layers ran at 23-55 % of their tensor floor).  Indices outside a view are the convolution's zero padding.
  bool sub = false;
  {
    const char* e = getenv("MIMAMO_STRIDED_VIEW");
    sub = L.ksize > 1 && L.stride == 2 && W >= 2 && H >= 2 && !(e && e[0] == '0');
    if (sub) vstride = 1;
  }
  // choose the output box (bw x bh x bn <= 128 pixels) that wastes the fewest MMA rows
  int best_bw = 1, best_bh = 1, best_bn = 1;
  long long best_tiles = -1;
  for (int bw = 1; bw <= Wo && bw <= 128; ++bw) {
    if (bw * vstride > 256) break;
    for (int bh = 1; bh <= Ho && bw * bh <= 128; ++bh) {
      if (bh * vstride > 256) break;
      int bn = 1;
      if (bw == Wo && bh == Ho) { bn = 128 / (bw * bh); if (bn > B) bn = B; if (bn < 1) bn = 1; }
      const long long tiles = (long long)((Wo + bw - 1) / bw) * ((Ho + bh - 1) / bh) * ((B + bn - 1) / bn);
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && bw > best_bw)) {
        best_tiles = tiles; best_bw = bw; best_bh = bh; best_bn = bn;
      }
    }
  }
  CUtensorMap ma, mb;
  const uint64_t dims[4] = {(uint64_t)in_channels, (uint64_t)Wv, (uint64_t)Hv, (uint64_t)B};
  const uint64_t strides[3] = {pix_pitch * 2, (uint64_t)W * in_pitch * 2 * (uint64_t)(L.stride / vstride), (uint64_t)H * W * in_pitch * 2};
  const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)(best_bw * vstride), (uint32_t)(best_bh * vstride), (uint32_t)best_bn};
  const uint32_t es[4] = {1, (uint32_t)vstride, (uint32_t)vstride, 1};
  int rc = sub ? 0 : encode_map(&ma, L.elem, 4, x, dims, strides, box, es);
  if (rc) return rc;
  ConvParams p;
  memset(&p, 0, sizeof(p));
  const int bn_eff = fill_common(p, L, out, ldc, residual, ld_res);
  p.mode = 1;
  p.Wo = Wo; p.Ho = Ho; p.Nimg = B;
  p.bw = best_bw; p.bh = best_bh; p.bn = best_bn;
  p.tiles_w = (Wo + best_bw - 1) / best_bw;
  p.tiles_h = (Ho + best_bh - 1) / best_bh;
  p.a_rows = best_bw * best_bh * best_bn;
  p.m_tiles = (int)best_tiles;
  p.stride = vstride;                           // the producer steps boxes in units of the tensor the map describes
  if (sub) {
    p.sub = 1; p.sub_pad = L.pad; p.pad = 0;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const uint64_t sdims[4] = {(uint64_t)in_channels, (uint64_t)((W - pw + 1) / 2), (uint64_t)((H - ph + 1) / 2), (uint64_t)B};
        const uint64_t sstr[3] = {(uint64_t)in_pitch * 4, (uint64_t)W * in_pitch * 4, (uint64_t)H * W * in_pitch * 2};
        const uint32_t sbox[4] = {(uint32_t)kBlockK, (uint32_t)best_bw, (uint32_t)best_bh, (uint32_t)best_bn};
        const uint32_t ses[4] = {1, 1, 1, 1};
        const uint16_t* base = reinterpret_cast<const uint16_t*>(x) + ((size_t)ph * W + pw) * in_pitch;
        rc = encode_map(&p.sub_map[ph * 2 + pw], L.elem, 4, base, sdims, sstr, sbox, ses);
        if (rc) return rc;
      }
    ma = p.sub_map[0];
  }
  p.pair = pair_wanted(bn_eff, p.num_k_blocks, p.m_tiles, false);
  if (p.pair == 2) p.idesc = make_idesc2(bn_eff, L.elem);
  rc = weight_map(L, &mb, p.pair ? bn_eff / 2 : bn_eff);
  if (rc) return rc;
  CUtensorMap mo;
  p.store_mode = store_mode_setting(L.ksize == 1 ? 1 : 2);
  rc = out_map_spatial(&mo, L.elem, out, ldc, L.Cout, Wo, Ho, B, best_bw, best_bh, best_bn, effective_block_n(L, residual != nullptr));
  if (rc) return rc;
  return launch(L, ma, mb, mo, p, stream);
}

// conv1 over the space-to-depth'ed, chunk-planar input (nn_kernels.cuh): the line kernel.
int conv1_s2d_forward(const ConvLayer& L, const void* s2d, int B, void* out, int ldc, cudaStream_t stream, bool pool) {
  MM_REQUIRE(L.Cin_p == 256 && L.ksize == 1 && L.Cout == 64, MIMAMO_E_VALUE, "conv1_s2d_forward needs the packed [64][256] layer");
  MM_REQUIRE(!pool || ldc == 64, MIMAMO_E_VALUE, "the fused pool1 output is a dense [B][56][56][64] tensor");
  if (B == 0) return MIMAMO_OK;
  const int S2D = 115, Wo = 112, Ho = 112;
  CUtensorMap ma, mb;
  {
    // weights [64][256] viewed as [tap (16)][n (64)][16 ch]: one box = the whole matrix, tap-major in smem
    const uint64_t wdims[3] = {16, 64, 16};
    const uint64_t wstr[2] = {512, 32};
    const uint32_t wbox[3] = {16, 64, 16};
    const uint32_t wes[3] = {1, 1, 1};
    int rc = encode_map(&mb, L.elem, 3, L.w_dev, wdims, wstr, wbox, wes, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc) return rc;
    ma = mb;                                                   // the patch needs no tensor map (plain bulk copies)
  }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  fill_common(p, L, out, ldc, nullptr, 0);
  p.mode = 2;
  p.Wo = Wo; p.Ho = Ho; p.Nimg = B;
  p.bw = S2D; p.bh = 1; p.bn = 1;
  p.tiles_w = 1; p.tiles_h = Ho;
  p.a_rows = 4 * S2D;
  p.a_ptr = s2d;
  p.m_tiles = Ho * B; p.n_tiles = 1;
  p.num_k_blocks = 4;                                        // K = 256 for the flop accounting
  static DeviceOnce attr_set;
  const bool bf = L.elem == kBF16;
  if (attr_set.need()) {
    MM_CUDA(cudaFuncSetAttribute(conv1_line_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLineSmemBytes));
    MM_CUDA(cudaFuncSetAttribute(conv1_line_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLineSmemBytes));
    MM_CUDA(cudaFuncSetAttribute(conv1_line_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLineSmemBytes));
    MM_CUDA(cudaFuncSetAttribute(conv1_line_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLineSmemBytes));
    attr_set.mark();
  }
  const int units = pool ? B * kPoolBands : p.m_tiles;
  const int grid = units < num_sms() ? units : num_sms();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_profile) {
    if (g_prof_used == g_prof_events.size()) {
      cudaEvent_t a0, a1;
      MM_CUDA(cudaEventCreate(&a0));
      MM_CUDA(cudaEventCreate(&a1));
      g_prof_events.emplace_back(a0, a1);
    }
    e0 = g_prof_events[g_prof_used].first; e1 = g_prof_events[g_prof_used].second;
    ++g_prof_used;
    g_prof_flops += 2.0 * (double)p.m_tiles * kBlockM * 64.0 * 256.0;
    MM_CUDA(cudaEventRecord(e0, stream));
  }
  p.store_mode = 1;
  CUtensorMap mo;                                              // one box = one 115-pixel padded line; columns >= 112 are clipped
  {
    int rc = out_map_spatial(&mo, L.elem, out, ldc, 64, Wo, Ho, B, S2D, 1, 1, 64);
    if (rc) return rc;
  }
  if (pool) {
    if (bf) conv1_line_kernel<true, true><<<grid, kGemmThreads, kLineSmemBytes, stream>>>(ma, mb, mo, p);
    else conv1_line_kernel<false, true><<<grid, kGemmThreads, kLineSmemBytes, stream>>>(ma, mb, mo, p);
  } else {
    if (bf) conv1_line_kernel<true, false><<<grid, kGemmThreads, kLineSmemBytes, stream>>>(ma, mb, mo, p);
    else conv1_line_kernel<false, false><<<grid, kGemmThreads, kLineSmemBytes, stream>>>(ma, mb, mo, p);
  }
  MM_LAUNCH_OK();
  if (e1) MM_CUDA(cudaEventRecord(e1, stream));
  return MIMAMO_OK;
}

}  // namespace mimamo

using namespace mimamo;

extern "C" int mimamo_profile_gemm(int32_t enable) {
  g_profile = enable != 0;
  g_prof_used = 0;
  g_prof_flops = 0.0;
  return MIMAMO_OK;
}

// Sum of the event-timed durations of every GEMM launch since mimamo_profile_gemm(1); the caller
// must have synchronised the stream.  issued_flops counts the MMA work actually issued (padded tiles).
extern "C" int mimamo_profile_gemm_read(double* total_ms, uint64_t* launches, double* issued_flops) {
  double tot = 0.0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    float ms = 0.f;
    MM_CUDA(cudaEventElapsedTime(&ms, g_prof_events[i].first, g_prof_events[i].second));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = g_prof_used;
  if (issued_flops) *issued_flops = g_prof_flops;
  return MIMAMO_OK;
}

// Per-launch durations (ms) of the launches recorded since mimamo_profile_gemm(1), in launch order; returns the count.
extern "C" int mimamo_profile_gemm_launches(float* ms_out, int32_t max_launches) {
  int n = 0;
  for (size_t i = 0; i < g_prof_used && n < max_launches; ++i, ++n) {
    float ms = 0.f;
    MM_CUDA(cudaEventElapsedTime(&ms, g_prof_events[i].first, g_prof_events[i].second));
    ms_out[n] = ms;
  }
  return n;
}

// Test hook (include/mimamo_b200.h): one convolution through the engine, bf16 NHWC in/out.
extern "C" int mimamo_conv_bf16(const void* x, int32_t B, int32_t H, int32_t W, int32_t Cin, const float* w_host,
                                const float* scale_host, const float* shift_host, int32_t Cout, int32_t ksize,
                                int32_t stride, int32_t pad, int32_t relu, const void* residual, void* out, void* stream) {
  MM_REQUIRE(x && w_host && scale_host && shift_host && out, MIMAMO_E_VALUE, "null argument");
  MM_REQUIRE(Cin % 64 == 0, MIMAMO_E_VALUE, "test hook needs Cin %% 64 == 0 (input pitch == padded Cin)");
  ConvLayer L;
  int rc = conv_layer_init(L, w_host, scale_host, shift_host, Cout, Cin, ksize, stride, pad, relu, kBF16);
  if (rc == MIMAMO_OK) rc = conv_forward(L, x, B, H, W, out, Cout, residual, Cout, (cudaStream_t)stream);
  if (rc == MIMAMO_OK && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
    set_error("conv kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    rc = MIMAMO_E_CUDA;
  }
  conv_layer_free(L);
  return rc;
}

// Test hook (include/mimamo_b200.h): increase-with-residual + next reduce through conv_chain_kernel, bf16 rows.
extern "C" int mimamo_conv_chain_bf16(const void* x, int32_t M, int32_t K1, const float* w1_host, const float* scale1_host,
                                      const float* shift1_host, int32_t N1, const void* residual, const float* w2_host,
                                      const float* scale2_host, const float* shift2_host, int32_t N2, void* out1, void* out2,
                                      void* stream) {
  MM_REQUIRE(x && w1_host && w2_host && residual && out1 && out2, MIMAMO_E_VALUE, "null argument");
  MM_REQUIRE(K1 % 64 == 0, MIMAMO_E_VALUE, "test hook needs K1 %% 64 == 0");
  ConvLayer L1, L2;
  int rc = conv_layer_init(L1, w1_host, scale1_host, shift1_host, N1, K1, 1, 1, 0, 1, kBF16);
  if (rc == MIMAMO_OK) rc = conv_layer_init(L2, w2_host, scale2_host, shift2_host, N2, N1, 1, 1, 0, 1, kBF16);
  if (rc == MIMAMO_OK) rc = chain_forward(L1, L2, x, M, out1, residual, N1, out2, N2, (cudaStream_t)stream);
  if (rc == MIMAMO_OK && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
    set_error("chain kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    rc = MIMAMO_E_CUDA;
  }
  conv_layer_free(L1);
  conv_layer_free(L2);
  return rc;
}
