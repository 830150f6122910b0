// tcgen05 implicit-GEMM convolution engine (sm_100a): declarations shared by resnet50.cu / head.cu.
//
// Replaces every cuDNN / cuBLAS convolution the reference issues through nn.Conv2d
// (ResNet50: api/resnet50_extractor.py:81; PhaseNet: api/mimamo_net.py:84-90).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace mimamo {

enum ElemType { kBF16 = 0, kF16 = 1 };

// One convolution (or plain GEMM) lowered to D[M=pixels][N=Cout] = A[M][K] * B[N][K]^T with
// K = taps * Cin_p, all operands 16-bit K-major, fp32 accumulation in TMEM.
struct ConvLayer {
  // static description
  int Cin = 0, Cin_p = 0;        // logical / padded (multiple of 64) input channels
  int Cout = 0;
  int ksize = 1, stride = 1, pad = 0;
  int relu = 0;
  ElemType elem = kBF16;
  void* w_dev = nullptr;         // [Cout][taps*Cin_p] 16-bit
  float* scale_dev = nullptr;    // [Cout]  (eval BatchNorm folded, or ones)
  float* shift_dev = nullptr;    // [Cout]
  int block_n = 64;
  std::vector<float> w_f32;      // the packed [Cout][taps*Cin_p] weights in fp32, kept so they can be re-rounded (conv_layer_quantize)
};

// pack torch-layout fp32 weights [Cout][Cin][k][k] into the engine layout and upload.
int conv_layer_init(ConvLayer& L, const float* w_host, const float* scale_host, const float* shift_host,
                    int Cout, int Cin, int ksize, int stride, int pad, int relu, ElemType elem);
void conv_layer_free(ConvLayer& L);

// Round L.w_f32 to the layer's 16-bit type and upload.  mode 0: round to nearest.  mode 1: mean-compensated rounding --
// per output row the rounding residuals r_k = q_k - w_k are steered (by rounding a few weights, those closest to a
// tie, the other way) so that sum_k r_k * mu[k % Cin_p] ~ 0, where mu is the expected value of input channel c
// (nullptr: all ones, i.e. post-ReLU inputs of comparable mean).  A conv's output error from weight rounding is
// sum_k r_k x_k; its part sum_k r_k E[x_k] is the SAME at every pixel of every image, so it survives every spatial
// average downstream (pool5) -- measured on the synthetic ResNet50 it is 90 % of the whole 16-bit error of the pooled
// features.  Costs nothing at run time.
int conv_layer_quantize(ConvLayer& L, int mode, const float* mu);

// x: NHWC 16-bit [B][H][W][Cin_p] (row pitch == Cin_p).  out: NHWC 16-bit, row pitch `ldc` elements,
// written at channel offset 0 of `out` (pass out + offset for concatenation).  residual: optional,
// same geometry as out with pitch ld_res.  Returns MIMAMO_OK or an error code.
// in_pitch / in_channels (3x3 and strided layers only): the input tensor holds `in_channels` <= Cin_p channels per pixel at a
// row pitch of `in_pitch` elements (a multiple of 8); the TMA box still spans Cin_p channels and the missing ones are
// zero-filled by the tensor map's out-of-bounds handling instead of being stored.  0 = the default dense Cin_p layout.
int conv_forward(const ConvLayer& L, const void* x, int B, int H, int W, void* out, int ldc,
                 const void* residual, int ld_res, cudaStream_t stream, int in_pitch = 0, int in_channels = 0);

// Plain GEMM view of the same kernel: A [M][K_p] 16-bit row-major (K_p multiple of 64).
int gemm_forward(const ConvLayer& L, const void* a, int M, void* out, int ldc, const void* residual,
                 int ld_res, cudaStream_t stream);

// Two flat 1x1 layers in one launch (conv_chain_kernel): out1 = act(L1(a) + residual), dense [M][L1.Cout] rows, and
// out2 = act(L2(out1)), [M][L2.Cout] at pitch ldc2.  L2 reads out1 back through L2 right after the tile was stored, so its
// activation read never reaches HBM.  chain_supported: both stride-1 1x1, L1.Cout % 128 == 0, L2.Cout <= 128.
bool chain_enabled();
bool chain_supported(const ConvLayer& L1, const ConvLayer& L2);
int chain_forward(const ConvLayer& L1, const ConvLayer& L2, const void* a, int M, void* out1, const void* residual, int ld_res,
                  void* out2, int ldc2, cudaStream_t stream);

int out_size(int in, int k, int s, int p);

// conv1_7x7_s2 over a space-to-depth'ed input (see nn_kernels: conv1_space_to_depth); L holds the
// re-packed [64][256] weights.
// pool = true fuses pool1_3x3_s2 (ceil_mode) into the epilogue: out is then the pooled [B][56][56][64] tensor
int conv1_s2d_forward(const ConvLayer& L, const void* s2d, int B, void* out, int ldc, cudaStream_t stream, bool pool = false);

}  // namespace mimamo
