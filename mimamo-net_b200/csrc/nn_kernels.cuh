// Small non-GEMM kernels around the tcgen05 engine: layout conversion, pooling, the fp32 linear
// layers of the head and the GRU recurrence.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "conv_engine.cuh"

namespace mimamo {

// fp32 NCHW [N][C][H][W] -> 16-bit NHWC rows of pitch `ldc`, written at channel offset `c_off`;
// channels [c_off + C, c_off + c_fill) are zero-filled (channel padding for the engine).
int nchw_to_nhwc16(const float* src, int N, int C, int H, int W, void* dst, int ldc, int c_off, int c_fill,
                   ElemType elem, cudaStream_t s);
// conv1_7x7_s2 lowering: fp32 NCHW [B][3][224][224] -> 16-bit [B*112*112][192] (K = c*49+kh*7+kw, zero padded)
int im2col_conv1(const float* x, int B, void* a, ElemType elem, cudaStream_t s);
// conv1 lowering without im2col: fp32 NCHW [B][3][224][224] -> 16-bit space-to-depth'ed, chunk-planar
// [B][115 rows Y][2 chunks][115 px X][8 ch]:  value(b, Y, X, k = (py*2+px)*3 + c) = x[b][c][2Y+py-4][2X+px-4]
// (zero outside the image and for k = 12..15), stored at chunk k / 8, lane k % 8.  For a fixed chunk the
// pixels of a row are contiguous 16-byte units -- the un-swizzled K-major UMMA operand layout -- and four
// consecutive rows form one contiguous block (conv_engine.cu, conv1_line_kernel).
int conv1_space_to_depth(const float* x, int B, void* s2d, ElemType elem, cudaStream_t s);
// pool1_3x3_s2: MaxPool(3, stride 2, pad 0, ceil_mode) on NHWC 16-bit
int maxpool3x3s2_ceil(const void* x, int B, int H, int W, int C, void* out, ElemType elem, cudaStream_t s);
// mean over HW positions of NHWC 16-bit -> fp32 [N][ldo] at column offset
int avgpool_to_f32(const void* x, int N, int HW, int C, float* out, int ldo, int relu, ElemType elem, cudaStream_t s);

// per-channel mean over the M rows of an NHWC 16-bit tensor [M][ld] -> mean_dev f32[C]; deterministic (fixed-order two-stage sum);
// scratch: at least C doubles per block (up to 1024 blocks are used)
int channel_means(const void* x, long long M, int C, int ld, float* mean_dev, double* scratch, size_t scratch_doubles, ElemType elem, cudaStream_t s);

// out[m][n] = post(relu?(pre(sum_k a[m][k] w[n][k] + bias[n])))  with per-column affine pre/post
struct LinearLayer {
  int in_f = 0, out_f = 0, relu = 0;
  float* w = nullptr;        // [out][in]
  float* bias = nullptr;     // [out]
  float* pre_s = nullptr;  float* pre_t = nullptr;    // applied before ReLU (Linear -> BN -> ReLU)
  float* post_s = nullptr; float* post_t = nullptr;   // applied after ReLU  (Linear -> ReLU -> BN)
};
int linear_init(LinearLayer& L, const float* w, const float* bias, int out_f, int in_f, int relu,
                const float* pre_s, const float* pre_t, const float* post_s, const float* post_t);
void linear_free(LinearLayer& L);
int linear_forward(const LinearLayer& L, const float* a, int lda, int M, float* out, int ldo, cudaStream_t s);

// One bidirectional GRU layer over dim 0.  xproj [S][Bt][2][3H] = W_ih x + b_ih per direction,
// whhT [2][H][3H], bhh [2][3H];  y [S][Bt][2H].
int gru_layer(const float* xproj, const float* whhT, const float* bhh, int S, int Bt, int Hd, float* y, cudaStream_t s);

}  // namespace mimamo
