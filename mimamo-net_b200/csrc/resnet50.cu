// R: ResNet50 (resnet50_ferplus_dag architecture) pool5 features on the tcgen05 engine.
//
// Replaces Resnet50_Extractor.get_vec (api/resnet50_extractor.py:74-83), i.e. the forward of the
// third-party module loaded at api/utils/model_utils.py:65-79 with a hook on `pool5_7x7_s1`.
// Architecture restated from SURVEY.md section 8(a) row R: Caffe-style ResNet-50, stride 2 on the
// first 1x1 (`_reduce`) and on `_proj` of stages 3-5, pool1 = MaxPool(3,2,pad 0,ceil_mode).
//
// Data layout: activations NHWC 16-bit (fp16 by default, MIMAMO_RESNET_DTYPE=bf16 selects bf16) resident in a caller-provided workspace,
// eval-mode BatchNorm folded into per-channel fp32 scale/shift applied in the GEMM epilogue,
// residual add + ReLU fused into the `_increase` epilogue.  The batch is processed in chunks so the
// workspace stays bounded and activations stay close to L2.
#include "common.cuh"
#include "conv_engine.cuh"
#include "nn_kernels.cuh"
#include "tensor_table.cuh"
#include <stdlib.h>

using namespace mimamo;

namespace {
struct ResBlock {
  ConvLayer reduce, conv3, increase, proj;
  bool has_proj = false;
  int stride = 1, mid = 0, cout = 0;
};
const int kStageBlocks[4] = {3, 4, 6, 3};
const int kStageMid[4] = {64, 128, 256, 512};
}  // namespace

struct mimamo_resnet50 {
  ElemType elem = kF16;            // fp16 activations (10-bit mantissa): the 1e-3 valence/arousal budget leaves no room for bf16 (measured
                                   // end to end: bf16 3.1e-3 .. 4.8e-3, fp16 within the budget); same UMMA kind::f16 rate, fp32 accumulation
  ConvLayer conv1;                 // 7x7 s2 as a 4x4 s1 conv over the space-to-depth'ed input (K = 256)
  ConvLayer conv1_im2col;          // fallback lowering: K = 147 -> 192 GEMM over an im2col buffer
  bool use_im2col = false;         // MIMAMO_CONV1=im2col
  bool fuse_pool = true;           // MIMAMO_CONV1_POOL=0: separate pool1 kernel (cross-check of the fused epilogue)
  bool chain = true;               // MIMAMO_CHAIN=0: `_increase` and the next `_reduce` as separate launches (stages 2-3 chain them, conv_engine.cuh)
  int calib = 2;                   // MIMAMO_RESNET_CALIB: weight rounding 0 = to nearest, 1 = zero-sum residuals, 2 = mean-compensated against
                                   // channel means measured on a built-in synthetic batch at create time (conv_layer_quantize)
  std::vector<ResBlock> blocks;
  int device = 0;                  // the CUDA device the weights live on
  int chunk = 2048;                // images per pass: larger chunks amortise per-launch ramp/tail (measured per 2048 images: 128: 41.4 ms, 512: 37.6 ms
                                   // on the first engine; 512: 28.3, 1024: 27.2, 2048: 26.5 ms now); 12 MB of workspace per image
};

static const size_t kPerImageElems =
    (size_t)12544 * 192 /*A0*/ + (size_t)12544 * 64 /*C1*/ + 3 * (size_t)802816 /*X,Y,SK*/ + 2 * (size_t)200704 /*T1,T2*/;

static int make_conv(const TensorTable& T, const std::string& name, int cout, int cin, int k, int stride, int pad,
                     int relu, ElemType elem, ConvLayer& L, int gemm_k = 0) {
  const float* w = T.get(name + ".weight", (int64_t)cout * cin * k * k);
  if (!w) return MIMAMO_E_VALUE;
  std::vector<float> sc, sh;
  if (!fold_bn(T, name + "_bn", cout, 1e-5f, nullptr, sc, sh)) return MIMAMO_E_VALUE;
  if (gemm_k) return conv_layer_init(L, w, sc.data(), sh.data(), cout, gemm_k, 1, 1, 0, relu, elem);
  return conv_layer_init(L, w, sc.data(), sh.data(), cout, cin, k, stride, pad, relu, elem);
}

extern "C" void mimamo_resnet50_destroy(mimamo_resnet50* net) {
  if (!net) return;
  conv_layer_free(net->conv1);
  conv_layer_free(net->conv1_im2col);
  for (auto& b : net->blocks) {
    conv_layer_free(b.reduce); conv_layer_free(b.conv3); conv_layer_free(b.increase);
    if (b.has_proj) conv_layer_free(b.proj);
  }
  delete net;
}

static int resnet50_self_calibrate(mimamo_resnet50* net);

extern "C" int mimamo_resnet50_create(const mimamo_tensor_desc* tensors, int32_t n_tensors, mimamo_resnet50** net_out) {
  MM_REQUIRE(tensors && net_out && n_tensors > 0, MIMAMO_E_VALUE, "null argument");
  TensorTable T{tensors, n_tensors};
  mimamo_resnet50* net = new mimamo_resnet50();
  net->device = current_device();
  const char* dt = getenv("MIMAMO_RESNET_DTYPE");
  net->elem = (dt && strcmp(dt, "bf16") == 0) ? kBF16 : kF16;
  const char* ck = getenv("MIMAMO_RESNET_CHUNK");
  if (ck && atoi(ck) > 0) net->chunk = atoi(ck);
  const char* c1 = getenv("MIMAMO_CONV1");
  net->use_im2col = c1 && strcmp(c1, "im2col") == 0;
  const char* fp = getenv("MIMAMO_CONV1_POOL");
  net->fuse_pool = !(fp && fp[0] == '0');
  net->chain = chain_enabled();
  const char* cb = getenv("MIMAMO_RESNET_CALIB");
  if (cb && cb[0] >= '0' && cb[0] <= '2') net->calib = cb[0] - '0';
  int rc = make_conv(T, "conv1_7x7_s2", 64, 3, 7, 2, 3, 1, net->elem, net->conv1_im2col, 147);
  if (!rc) {
    // re-pack [64][3][7][7] into the s2d kernel [64][kh'(4)][kw'(4)][(py*2+px)*3+c (16)]
    const float* w = T.get("conv1_7x7_s2.weight", 64 * 147);
    std::vector<float> sc, sh, w2((size_t)64 * 256, 0.f);
    if (!w || !fold_bn(T, "conv1_7x7_s2_bn", 64, 1e-5f, nullptr, sc, sh)) rc = MIMAMO_E_VALUE;
    for (int o = 0; o < 64 && !rc; ++o)
      for (int khp = 0; khp < 4; ++khp)
        for (int kwp = 0; kwp < 4; ++kwp)
          for (int q = 0; q < 4; ++q) {
            const int kh = 2 * khp + (q >> 1) - 1, kw = 2 * kwp + (q & 1) - 1;
            if (kh < 0 || kh > 6 || kw < 0 || kw > 6) continue;
            for (int c = 0; c < 3; ++c)
              w2[(size_t)o * 256 + khp * 64 + kwp * 16 + q * 3 + c] = w[((size_t)o * 3 + c) * 49 + kh * 7 + kw];
          }
    if (!rc) rc = conv_layer_init(net->conv1, w2.data(), sc.data(), sh.data(), 64, 256, 1, 1, 0, 1, net->elem);
  }
  int cin = 64;
  for (int s = 0; s < 4 && rc == MIMAMO_OK; ++s) {
    const int mid = kStageMid[s], cout = mid * 4;
    for (int b = 1; b <= kStageBlocks[s] && rc == MIMAMO_OK; ++b) {
      net->blocks.emplace_back();
      ResBlock& blk = net->blocks.back();
      blk.stride = (b == 1 && s > 0) ? 2 : 1;
      blk.mid = mid; blk.cout = cout; blk.has_proj = (b == 1);
      char p[64];
      snprintf(p, sizeof(p), "conv%d_%d_", s + 2, b);
      const std::string pre(p);
      rc = make_conv(T, pre + "1x1_reduce", mid, cin, 1, blk.stride, 0, 1, net->elem, blk.reduce);
      if (!rc) rc = make_conv(T, pre + "3x3", mid, mid, 3, 1, 1, 1, net->elem, blk.conv3);
      if (!rc) rc = make_conv(T, pre + "1x1_increase", cout, mid, 1, 1, 0, 1 /*relu after residual*/, net->elem, blk.increase);
      if (!rc && blk.has_proj) rc = make_conv(T, pre + "1x1_proj", cout, cin, 1, blk.stride, 0, 0, net->elem, blk.proj);
      cin = cout;
    }
  }
  if (rc == MIMAMO_OK && net->calib == 1) {
    for (auto& b : net->blocks) {                              // every layer past conv1 reads post-ReLU activations
      ConvLayer* ls[4] = {&b.reduce, &b.conv3, &b.increase, b.has_proj ? &b.proj : nullptr};
      for (ConvLayer* l : ls) if (l && !rc) rc = conv_layer_quantize(*l, 1, nullptr);
    }
  }
  if (rc == MIMAMO_OK && net->calib == 2) rc = resnet50_self_calibrate(net);
  if (rc != MIMAMO_OK) { mimamo_resnet50_destroy(net); return rc; }
  *net_out = net;
  return MIMAMO_OK;
}

extern "C" int mimamo_resnet50_workspace_bytes(const mimamo_resnet50* net, int32_t batch, size_t* bytes_out) {
  MM_REQUIRE(net && bytes_out && batch >= 0, MIMAMO_E_VALUE, "bad arguments");
  const int chunk = batch < net->chunk ? (batch > 0 ? batch : 1) : net->chunk;
  *bytes_out = (size_t)chunk * kPerImageElems * 2 + 4096;
  return MIMAMO_OK;
}

struct mimamo_preproc;
namespace mimamo {
int crops_rgb_launch(const mimamo_preproc* p, const uint8_t* crops, int64_t n, void* out, int mode, cudaStream_t s);
}
extern "C" int mimamo_preproc_geometry(const mimamo_preproc* p, int32_t* src, int32_t* gray_size, int32_t* crop);

// x != nullptr: fp32 NCHW frames; otherwise uint8 face crops preprocessed on the fly (preproc.cu)
// Channel means of every layer's input, measured during a forward pass (weight-rounding calibration).
namespace {
struct Calib {
  std::vector<std::pair<const ConvLayer*, std::vector<float>>> mus;
  float* mean_dev = nullptr;
  double* scratch = nullptr;
  size_t scratch_doubles = 0;
  int measure(const ConvLayer* L, const void* x, long long M, cudaStream_t s) {
    const int C = L->Cin_p;
    int rc = channel_means(x, M, C, C, mean_dev, scratch, scratch_doubles, L->elem, s);
    if (rc) return rc;
    std::vector<float> mu((size_t)C);
    MM_CUDA(cudaMemcpyAsync(mu.data(), mean_dev, sizeof(float) * C, cudaMemcpyDeviceToHost, s));
    MM_CUDA(cudaStreamSynchronize(s));
    mus.emplace_back(L, std::move(mu));
    return MIMAMO_OK;
  }
};
}  // namespace

static int resnet50_forward(const mimamo_resnet50* net, const float* x, const mimamo_preproc* pre, const uint8_t* crops,
                            int crop_edge, int32_t batch, float* out, void* workspace, size_t workspace_bytes,
                            cudaStream_t stream, Calib* cal = nullptr) {
  size_t need = 0;
  mimamo_resnet50_workspace_bytes(net, batch, &need);
  MM_REQUIRE(workspace && workspace_bytes >= need, MIMAMO_E_VALUE, "workspace too small: need %zu bytes", need);
  MM_CHECK_DEVICE(net->device);
  const int chunk = batch < net->chunk ? batch : net->chunk;
  uint16_t* base = reinterpret_cast<uint16_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  uint16_t* A0 = base;
  uint16_t* C1 = A0 + (size_t)chunk * 12544 * 192;
  uint16_t* X = C1 + (size_t)chunk * 12544 * 64;
  uint16_t* Y = X + (size_t)chunk * 802816;
  uint16_t* SK = Y + (size_t)chunk * 802816;
  uint16_t* T1 = SK + (size_t)chunk * 802816;
  uint16_t* T2 = T1 + (size_t)chunk * 200704;
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int Bc = batch - b0 < chunk ? batch - b0 : chunk;
    int rc;
    bool pooled = false;                                       // conv1 kernel already produced pool1's output in X
    if (crops) {
      // resize + centre crop + mean subtraction straight into the space-to-depth'ed conv1 operand
      rc = crops_rgb_launch(pre, crops + (size_t)b0 * crop_edge * crop_edge * 3, Bc, A0, net->elem == kBF16 ? 1 : 2, stream);
      pooled = net->fuse_pool;
      if (!rc) rc = conv1_s2d_forward(net->conv1, A0, Bc, pooled ? X : C1, 64, stream, pooled);
    } else if (net->use_im2col) {
      rc = im2col_conv1(x + (size_t)b0 * 3 * 224 * 224, Bc, A0, net->elem, stream);
      if (!rc) rc = gemm_forward(net->conv1_im2col, A0, Bc * 12544, C1, 64, nullptr, 0, stream);
    } else {
      rc = conv1_space_to_depth(x + (size_t)b0 * 3 * 224 * 224, Bc, A0, net->elem, stream);
      pooled = net->fuse_pool;
      if (!rc) rc = conv1_s2d_forward(net->conv1, A0, Bc, pooled ? X : C1, 64, stream, pooled);
    }
    if (!rc && !pooled) rc = maxpool3x3s2_ceil(C1, Bc, 112, 112, 64, X, net->elem, stream);
    uint16_t* cur = X;
    uint16_t* nxt = Y;
    int H = 56;
    bool chained = false;                                      // this block's `_reduce` was computed by the previous block's chain launch
    for (size_t i = 0; i < net->blocks.size() && !rc; ++i) {
      const ResBlock& blk = net->blocks[i];
      const int Ho = out_size(H, 1, blk.stride, 0);
      if (cal) {
        rc = cal->measure(&blk.reduce, cur, (long long)Bc * H * H, stream);
        if (!rc && blk.has_proj) cal->mus.emplace_back(&blk.proj, cal->mus.back().second);    // same input tensor
      }
      if (!rc && !chained) rc = conv_forward(blk.reduce, cur, Bc, H, H, T1, blk.mid, nullptr, 0, stream);
      if (!rc && cal) rc = cal->measure(&blk.conv3, T1, (long long)Bc * Ho * Ho, stream);
      if (!rc) rc = conv_forward(blk.conv3, T1, Bc, Ho, Ho, T2, blk.mid, nullptr, 0, stream);
      const uint16_t* res = cur;
      if (!rc && blk.has_proj) {
        rc = conv_forward(blk.proj, cur, Bc, H, H, SK, blk.cout, nullptr, 0, stream);
        res = SK;
      }
      if (!rc && cal) rc = cal->measure(&blk.increase, T2, (long long)Bc * Ho * Ho, stream);
      // `_increase` of this block and `_reduce` of the next one in one launch (conv_chain_kernel): the next block then
      // finds its T1 ready.  T1 is free here: this block's 3x3 has consumed it.
      chained = !cal && net->chain && i + 1 < net->blocks.size() && net->blocks[i + 1].stride == 1 &&
                chain_supported(blk.increase, net->blocks[i + 1].reduce);
      if (!rc && chained)
        rc = chain_forward(blk.increase, net->blocks[i + 1].reduce, T2, Bc * Ho * Ho, nxt, res, blk.cout, T1, net->blocks[i + 1].mid, stream);
      else if (!rc) rc = conv_forward(blk.increase, T2, Bc, Ho, Ho, nxt, blk.cout, res, blk.cout, stream);
      uint16_t* t = cur; cur = nxt; nxt = t;
      H = Ho;
    }
    // pool5_7x7_s1 + the (no-op) relu of get_vec, straight to fp32
    if (!rc) rc = avgpool_to_f32(cur, Bc, 49, 2048, out + (size_t)b0 * 2048, 2048, 1, net->elem, stream);
    if (rc) return rc;
  }
  return MIMAMO_OK;
}

extern "C" int mimamo_resnet50_pool5(const mimamo_resnet50* net, const float* x, int32_t batch, float* out, void* workspace,
                                     size_t workspace_bytes, void* stream_) {
  MM_REQUIRE(net && x && out && batch >= 0, MIMAMO_E_VALUE, "bad arguments");
  if (batch == 0) return MIMAMO_OK;
  return resnet50_forward(net, x, nullptr, nullptr, 0, batch, out, workspace, workspace_bytes, (cudaStream_t)stream_);
}

// Re-round every layer's weights (past conv1, whose input is zero-centred) against the channel means of ITS input on
// the given images: the part of the weight-rounding error that is common to all pixels cancels (conv_engine.cuh).
extern "C" int mimamo_resnet50_calibrate(mimamo_resnet50* net, const float* x, int32_t batch, void* workspace,
                                         size_t workspace_bytes, void* stream_) {
  MM_REQUIRE(net && x && batch >= 1, MIMAMO_E_VALUE, "bad arguments");
  MM_REQUIRE(batch <= net->chunk, MIMAMO_E_VALUE, "calibration batch must fit one pass (%d images)", net->chunk);
  cudaStream_t stream = (cudaStream_t)stream_;
  Calib cal;
  cal.scratch_doubles = (size_t)1024 * 2048;
  float* feats = nullptr;
  MM_CUDA(cudaMalloc(&cal.mean_dev, sizeof(float) * 2048));
  MM_CUDA(cudaMalloc(&cal.scratch, sizeof(double) * cal.scratch_doubles));
  MM_CUDA(cudaMalloc(&feats, sizeof(float) * 2048 * (size_t)batch));
  int rc = resnet50_forward(net, x, nullptr, nullptr, 0, batch, feats, workspace, workspace_bytes, stream, &cal);
  if (!rc && cudaStreamSynchronize(stream) != cudaSuccess) { set_error("calibration pass failed: %s", cudaGetErrorString(cudaGetLastError())); rc = MIMAMO_E_CUDA; }
  for (auto& m : cal.mus)
    if (!rc) rc = conv_layer_quantize(*const_cast<ConvLayer*>(m.first), 1, m.second.data());
  cudaFree(cal.mean_dev); cudaFree(cal.scratch); cudaFree(feats);
  return rc;
}

// Built-in calibration batch: 8 images of uniform 0..255 noise minus the channel means (what the synthetic parity inputs
// look like; recalibrate on real face crops with mimamo_resnet50_calibrate).  Deterministic, so every rank / run rounds
// its weights identically.
static int resnet50_self_calibrate(mimamo_resnet50* net) {
  const int n = 8;
  const float mean[3] = {131.0912f, 103.8827f, 91.4953f};
  std::vector<float> img((size_t)n * 3 * 224 * 224);
  uint64_t state = 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < img.size(); ++i) {
    state = state * 6364136223846793005ull + 1442695040888963407ull;
    img[i] = (float)((state >> 56) & 0xFF) - mean[(i / (224 * 224)) % 3];
  }
  float* x_dev = nullptr;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  mimamo_resnet50_workspace_bytes(net, n, &ws_bytes);
  MM_CUDA(cudaMalloc(&x_dev, img.size() * sizeof(float)));
  if (cudaMalloc(&ws, ws_bytes) != cudaSuccess) { cudaFree(x_dev); set_error("cudaMalloc of the calibration workspace failed"); return MIMAMO_E_CUDA; }
  cudaMemcpy(x_dev, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice);
  const int rc = mimamo_resnet50_calibrate(net, x_dev, n, ws, ws_bytes, nullptr);
  cudaFree(x_dev); cudaFree(ws);
  return rc;
}

// uint8 face crops [batch, S, S, 3] -> pool5: Image_Sampler's transform (api/utils/model_utils.py:26-40) runs on the
// device, bit-exact with PIL, and feeds conv1 directly (no fp32 frame is ever materialised).
extern "C" int mimamo_resnet50_pool5_crops(const mimamo_resnet50* net, const mimamo_preproc* pre, const uint8_t* crops,
                                           int32_t batch, float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  MM_REQUIRE(net && pre && crops && out && batch >= 0, MIMAMO_E_VALUE, "bad arguments");
  if (batch == 0) return MIMAMO_OK;
  int32_t src = 0, crop = 0;
  mimamo_preproc_geometry(pre, &src, nullptr, &crop);
  MM_REQUIRE(crop == 224, MIMAMO_E_RUNTIME, "ResNet50 needs 224x224 centre crops, the plan produces %d", crop);
  return resnet50_forward(net, nullptr, pre, crops, src, batch, out, workspace, workspace_bytes, (cudaStream_t)stream_);
}
