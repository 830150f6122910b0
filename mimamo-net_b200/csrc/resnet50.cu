// R: ResNet50 (resnet50_ferplus_dag architecture) pool5 features on the tcgen05 engine.
//
// Replaces Resnet50_Extractor.get_vec (api/resnet50_extractor.py:74-83), i.e. the forward of the
// third-party module loaded at api/utils/model_utils.py:65-79 with a hook on `pool5_7x7_s1`.
// Architecture restated from SURVEY.md section 8(a) row R: Caffe-style ResNet-50, stride 2 on the
// first 1x1 (`_reduce`) and on `_proj` of stages 3-5, pool1 = MaxPool(3,2,pad 0,ceil_mode).
//
// Data layout: activations NHWC 16-bit (bf16 by default) resident in a caller-provided workspace,
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
  ElemType elem = kBF16;
  ConvLayer conv1;                 // 7x7 s2 as a 4x4 s1 conv over the space-to-depth'ed input (K = 256)
  ConvLayer conv1_im2col;          // fallback lowering: K = 147 -> 192 GEMM over an im2col buffer
  bool use_im2col = false;         // MIMAMO_CONV1=im2col
  bool fuse_pool = true;           // MIMAMO_CONV1_POOL=0: separate pool1 kernel (cross-check of the fused epilogue)
  std::vector<ResBlock> blocks;
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

extern "C" int mimamo_resnet50_create(const mimamo_tensor_desc* tensors, int32_t n_tensors, mimamo_resnet50** net_out) {
  MM_REQUIRE(tensors && net_out && n_tensors > 0, MIMAMO_E_VALUE, "null argument");
  TensorTable T{tensors, n_tensors};
  mimamo_resnet50* net = new mimamo_resnet50();
  const char* dt = getenv("MIMAMO_RESNET_DTYPE");
  net->elem = (dt && strcmp(dt, "fp16") == 0) ? kF16 : kBF16;
  const char* ck = getenv("MIMAMO_RESNET_CHUNK");
  if (ck && atoi(ck) > 0) net->chunk = atoi(ck);
  const char* c1 = getenv("MIMAMO_CONV1");
  net->use_im2col = c1 && strcmp(c1, "im2col") == 0;
  const char* fp = getenv("MIMAMO_CONV1_POOL");
  net->fuse_pool = !(fp && fp[0] == '0');
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
static int resnet50_forward(const mimamo_resnet50* net, const float* x, const mimamo_preproc* pre, const uint8_t* crops,
                            int crop_edge, int32_t batch, float* out, void* workspace, size_t workspace_bytes,
                            cudaStream_t stream) {
  size_t need = 0;
  mimamo_resnet50_workspace_bytes(net, batch, &need);
  MM_REQUIRE(workspace && workspace_bytes >= need, MIMAMO_E_VALUE, "workspace too small: need %zu bytes", need);
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
    for (size_t i = 0; i < net->blocks.size() && !rc; ++i) {
      const ResBlock& blk = net->blocks[i];
      const int Ho = out_size(H, 1, blk.stride, 0);
      rc = conv_forward(blk.reduce, cur, Bc, H, H, T1, blk.mid, nullptr, 0, stream);
      if (!rc) rc = conv_forward(blk.conv3, T1, Bc, Ho, Ho, T2, blk.mid, nullptr, 0, stream);
      const uint16_t* res = cur;
      if (!rc && blk.has_proj) {
        rc = conv_forward(blk.proj, cur, Bc, H, H, SK, blk.cout, nullptr, 0, stream);
        res = SK;
      }
      if (!rc) rc = conv_forward(blk.increase, T2, Bc, Ho, Ho, nxt, blk.cout, res, blk.cout, stream);
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
