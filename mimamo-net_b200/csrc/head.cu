// H: two-stream head (MLP || PhaseNet -> transform -> 2-layer bidirectional GRU -> classifier).
//
// Replaces Two_Stream_RNN.forward in eval mode (api/mimamo_net.py:129-143; MLP :22-26, PhaseNet
// :79-95).  Dropout layers are identities and every BatchNorm is folded into a per-channel affine.
// PhaseNet's six 3x3 convolutions (99 % of the head's flops) run on the tcgen05 engine in fp16
// (10-bit mantissa: the 1e-3 valence/arousal budget leaves no room for bf16 here); the stride-2
// layers use TMA element strides, the level-1 skip concat (mimamo_net.py:85) is a channel offset
// into a shared NHWC buffer.  The small dense layers and the GRU stay in fp32.
#include "common.cuh"
#include "conv_engine.cuh"
#include "nn_kernels.cuh"
#include "tensor_table.cuh"
#include <stdlib.h>
#include <string>

using namespace mimamo;

namespace {
// MLP (api/mimamo_net.py:6-26): [Dropout, Linear, BN, ReLU] per hidden layer
struct MlpPart {
  std::vector<LinearLayer> layers;
  int in_f = 0, max_f = 0;
};
// PhaseNet (api/mimamo_net.py:27-95): input size 48 -> three conv blocks (64, 128, 256 channels), 96 / 112 -> four (.., 512);
// every block is conv3x3 + BN + ReLU, conv3x3 stride 2 + BN + ReLU; the level-1 maps join after block 0
struct PhaseNetPart {
  int cin0 = 24;                          // phase channels = nbands * num_phase
  int size = 48;                          // level-0 map edge; level-1 maps are size / 2
  int n_blocks = 3;
  ConvLayer conv[8];                      // conv_net.{0..3}.{0,3}
  LinearLayer fc0, fc4, cls;
  bool has_cls = false;
  int chunk = 1024;                       // windows per convolution pass
};
}  // namespace

struct mimamo_mlp { MlpPart p; };
struct mimamo_phasenet { PhaseNetPart p; };

struct mimamo_head {
  int num_phase = 12;
  int device = 0;                         // the CUDA device the weights live on
  MlpPart mlp;
  PhaseNetPart pn;
  LinearLayer transform, xproj[2], classifier;
  float* whhT[2] = {nullptr, nullptr};    // [2 dir][128][384] per layer
  float* bhh[2] = {nullptr, nullptr};     // [2 dir][384] per layer
};

static const ElemType kHeadElem = kF16;

static int make_phase_conv(const TensorTable& T, const std::string& conv, const std::string& bn, int cout, int cin,
                           int stride, ConvLayer& L) {
  const float* w = T.get(conv + ".weight", (int64_t)cout * cin * 9);
  const float* b = T.get(conv + ".bias", cout);
  if (!w || !b) return MIMAMO_E_VALUE;
  std::vector<float> sc, sh;
  if (!fold_bn(T, bn, cout, 1e-5f, b, sc, sh)) return MIMAMO_E_VALUE;
  return conv_layer_init(L, w, sc.data(), sh.data(), cout, cin, 3, stride, 1, 1, kHeadElem);
}

// Weight rounding (conv_engine.cuh, conv_layer_quantize): PhaseNet's convolutions past the first read post-ReLU
// activations, so their rounding residuals are steered to sum to zero over those input channels; the phase-difference
// inputs themselves (conv_net[0][0], and the level-1 channels of the skip concat) are zero-mean by construction
// (api/phase_difference_extractor.py:127-131) and stay round-to-nearest.  MIMAMO_HEAD_CALIB=0 disables.
static int compensate_phase_conv(ConvLayer& L, int relu_channels) {
  std::vector<float> mu((size_t)L.Cin_p, 0.f);
  for (int c = 0; c < relu_channels && c < L.Cin_p; ++c) mu[c] = 1.f;
  return conv_layer_quantize(L, 1, mu.data());
}

// Linear -> BN -> ReLU (bn_first) or Linear -> ReLU -> BN, or plain Linear (+BN)
static int make_linear(const TensorTable& T, const std::string& lin, const std::string& bn, int out_f, int in_f, int relu,
                       bool bn_first, LinearLayer& L, float bn_eps = 1e-5f) {
  const float* w = T.get(lin + ".weight", (int64_t)out_f * in_f);
  const float* b = T.get(lin + ".bias", out_f);
  if (!w || !b) return MIMAMO_E_VALUE;
  std::vector<float> sc, sh;
  if (!bn.empty() && !fold_bn(T, bn, out_f, bn_eps, nullptr, sc, sh)) return MIMAMO_E_VALUE;
  const float* s = bn.empty() ? nullptr : sc.data();
  const float* t = bn.empty() ? nullptr : sh.data();
  return bn_first ? linear_init(L, w, b, out_f, in_f, relu, s, t, nullptr, nullptr)
                  : linear_init(L, w, b, out_f, in_f, relu, nullptr, nullptr, s, t);
}

// `pre` = key prefix of the nn.Sequential ("mlp.mlp." inside Two_Stream_RNN, "mlp." for a bare MLP); layer i owns
// keys 4i+1 (Linear) and 4i+2 (BatchNorm1d).  Any depth / widths; the last width must be 256 (reference :12).
static int mlp_init(const TensorTable& T, const std::string& pre, MlpPart& P) {
  for (int i = 0;; ++i) {
    const std::string lin = pre + std::to_string(4 * i + 1), bn = pre + std::to_string(4 * i + 2);
    const mimamo_tensor_desc* d = T.find(lin + ".weight");
    if (!d) break;
    MM_REQUIRE(d->ndim == 2, MIMAMO_E_VALUE, "'%s.weight' must be 2-D", lin.c_str());
    const int out_f = (int)d->shape[0], in_f = (int)d->shape[1];
    MM_REQUIRE(P.layers.empty() || P.layers.back().out_f == in_f, MIMAMO_E_VALUE, "MLP layer %d does not chain", i);
    P.layers.emplace_back();
    int rc = make_linear(T, lin, bn, out_f, in_f, 1, true, P.layers.back());
    if (rc) return rc;
    if (i == 0) P.in_f = in_f;
    P.max_f = out_f > P.max_f ? out_f : P.max_f;
  }
  MM_REQUIRE(!P.layers.empty(), MIMAMO_E_VALUE, "state_dict holds no '%s1.weight'", pre.c_str());
  MM_REQUIRE(P.layers.back().out_f == 256, MIMAMO_E_RUNTIME, "the MLP must end in 256 features (api/mimamo_net.py:12)");
  return MIMAMO_OK;
}
static void mlp_free(MlpPart& P) { for (auto& l : P.layers) linear_free(l); P.layers.clear(); }
static size_t mlp_tmp_floats(const MlpPart& P, int M) { return P.layers.size() > 1 ? 2 * (size_t)M * P.max_f : 0; }
// x [M][in_f] -> out [M][ldo] (256 columns); tmp: mlp_tmp_floats
static int mlp_run(const MlpPart& P, const float* x, int M, float* out, int ldo, float* tmp, cudaStream_t s) {
  const float* cur = x;
  int ld = P.in_f;
  for (size_t i = 0; i < P.layers.size(); ++i) {
    const bool last = i + 1 == P.layers.size();
    float* dst = last ? out : tmp + (i & 1) * (size_t)M * P.max_f;
    const int ldd = last ? ldo : P.layers[i].out_f;
    int rc = linear_forward(P.layers[i], cur, ld, M, dst, ldd, s);
    if (rc) return rc;
    cur = dst; ld = ldd;
  }
  return MIMAMO_OK;
}

static int phasenet_init(const TensorTable& T, const std::string& pre, int cin0, PhaseNetPart& P, int size = 48) {
  MM_REQUIRE(cin0 >= 1 && cin0 <= 64, MIMAMO_E_RUNTIME, "PhaseNet supports 1..64 phase channels (2*num_phase), got %d", cin0);
  MM_REQUIRE(size == 48 || size == 96 || size == 112, MIMAMO_E_VALUE, "Incorrect input size");      // reference :31-32
  P.cin0 = cin0;
  P.size = size;
  P.n_blocks = size == 48 ? 3 : 4;
  { const char* e = getenv("MIMAMO_HEAD_CHUNK"); if (e && atoi(e) > 0) P.chunk = atoi(e); }
  if (size != 48 && P.chunk > 256) P.chunk = 256;              // 4-5x the activations per window
  int rc = MIMAMO_OK;
  for (int b = 0; b < P.n_blocks && !rc; ++b) {
    const int cout = 64 << b, cin = b == 0 ? cin0 : (b == 1 ? cin0 + 64 : 64 << (b - 1));
    const std::string blk = pre + "conv_net." + std::to_string(b) + ".";
    rc = make_phase_conv(T, blk + "0", blk + "1", cout, cin, 1, P.conv[2 * b]);
    if (!rc) rc = make_phase_conv(T, blk + "3", blk + "4", cout, cout, 2, P.conv[2 * b + 1]);
  }
  {
    const char* e = getenv("MIMAMO_HEAD_CALIB");
    if (!(e && e[0] == '0')) {
      // post-ReLU input channels of conv_net.{b}.{0,3}: all of them except the phase inputs (block 0's first conv, and
      // the level-1 channels [64, 64 + cin0) of the skip concatenation read by block 1's first conv)
      for (int i = 1; i < 2 * P.n_blocks && !rc; ++i) rc = compensate_phase_conv(P.conv[i], i == 2 ? 64 : P.conv[i].Cin);
    }
  }
  const int last = 64 << (P.n_blocks - 1);
  if (!rc) rc = make_linear(T, pre + "fc.0", pre + "fc.2", 256, last, 1, false, P.fc0);
  if (!rc) rc = make_linear(T, pre + "fc.4", pre + "fc.6", 256, 256, 1, false, P.fc4);
  if (!rc && T.find(pre + "classifier.0.weight")) {            // Linear(256, 1) + BatchNorm1d(1, eps=1e-6) (reference :62-64)
    rc = make_linear(T, pre + "classifier.0", pre + "classifier.1", 1, 256, 0, true, P.cls, 1e-6f);
    P.has_cls = rc == MIMAMO_OK;
  }
  return rc;
}
static void phasenet_free(PhaseNetPart& P) {
  for (auto& c : P.conv) conv_layer_free(c);
  linear_free(P.fc0); linear_free(P.fc4); linear_free(P.cls);
}

namespace {
struct PhaseNetLayout { size_t pool, fc, a0, mid[8], total; };       // mid[i]: output of conv[i] (mid[1] doubles as the skip concat)
// spatial edge of conv[i]'s output, its channel pitch
inline int pn_edge(const PhaseNetPart& P, int i) { return P.size >> ((i + 1) / 2); }
inline int pn_pitch(const PhaseNetPart& P, int i) { return i == 1 ? 128 : P.conv[i].Cout; }
PhaseNetLayout phasenet_layout(const PhaseNetPart& P, int M) {
  PhaseNetLayout L;
  size_t cur = 0;
  auto take = [&](size_t bytes) { size_t at = cur; cur += align_up(bytes, 1024); return at; };
  const size_t Mc = (size_t)(M < P.chunk ? M : P.chunk);
  const int last = 64 << (P.n_blocks - 1);
  L.pool = take((size_t)M * last * 4);
  L.fc = take((size_t)M * 256 * 4);
  L.a0 = take(Mc * P.size * P.size * 64 * 2);                // phase_0 as NHWC (cin0 -> 64 channels)
  for (int i = 0; i < 2 * P.n_blocks; ++i) {
    const size_t e = (size_t)pn_edge(P, i);
    L.mid[i] = take(Mc * e * e * pn_pitch(P, i) * 2);        // mid[1] = [conv_net[0][3] (64) | phase_1 (cin0) | zero pad] at pitch 128
  }
  L.total = cur;
  return L;
}
}  // namespace

// phase_0 f32[M,cin0,S,S], phase_1 f32[M,cin0,S/2,S/2] -> out[m][0..256) at pitch ldo (the `feature=True` output).
// Alternatively (a0_nhwc != nullptr, S = 48) the inputs arrive as the phase tail writes them for this net: a0_nhwc f16
// [M][48][48][a0_pitch] holding the cin0 level-0 channels, cat_nhwc f16 [M][24][24][128] holding the level-1 channels at
// [64, 64 + cin0) and zeros above (channels [0,64) are overwritten here by conv_net[0][3]).
static int phasenet_run(const PhaseNetPart& P, const float* phase_0, const float* phase_1, int M, float* out, int ldo,
                        char* ws, cudaStream_t s, const uint16_t* a0_nhwc = nullptr, int a0_pitch = 0, uint16_t* cat_nhwc = nullptr) {
  const PhaseNetLayout L = phasenet_layout(P, M);
  float* pool = (float*)(ws + L.pool);
  float* fc = (float*)(ws + L.fc);
  const int c0 = P.cin0, S = P.size, S1 = P.size / 2, nc = 2 * P.n_blocks, last = 64 << (P.n_blocks - 1);
  int rc = MIMAMO_OK;
  for (int m0 = 0; m0 < M && !rc; m0 += P.chunk) {
    const int Mc = M - m0 < P.chunk ? M - m0 : P.chunk;
    void* a0 = ws + L.a0;
    void* mid[8];
    for (int i = 0; i < nc; ++i) mid[i] = ws + L.mid[i];
    if (a0_nhwc) {
      mid[1] = cat_nhwc + (size_t)m0 * S1 * S1 * 128;
      rc = conv_forward(P.conv[0], a0_nhwc + (size_t)m0 * S * S * a0_pitch, Mc, S, S, mid[0], 64, nullptr, 0, s, a0_pitch, c0);
    } else {
      rc = nchw_to_nhwc16(phase_0 + (size_t)m0 * c0 * S * S, Mc, c0, S, S, a0, 64, 0, 64, kHeadElem, s);
      if (!rc) rc = nchw_to_nhwc16(phase_1 + (size_t)m0 * c0 * S1 * S1, Mc, c0, S1, S1, mid[1], 128, 64, 64, kHeadElem, s);
      if (!rc) rc = conv_forward(P.conv[0], a0, Mc, S, S, mid[0], 64, nullptr, 0, s);
    }
    for (int i = 1; i < nc && !rc; ++i) {                      // conv[1] writes channels [0,64) of the skip concat (pitch 128)
      const int e_in = pn_edge(P, i - 1);
      rc = conv_forward(P.conv[i], mid[i - 1], Mc, e_in, e_in, mid[i], pn_pitch(P, i), nullptr, 0, s);
    }
    const int e_out = pn_edge(P, nc - 1);
    if (!rc) rc = avgpool_to_f32(mid[nc - 1], Mc, e_out * e_out, last, pool + (size_t)m0 * last, last, 0, kHeadElem, s);
  }
  if (!rc) rc = linear_forward(P.fc0, pool, last, M, fc, 256, s);
  if (!rc) rc = linear_forward(P.fc4, fc, 256, M, out, ldo, s);
  return rc;
}

extern "C" void mimamo_head_destroy(mimamo_head* h) {
  if (!h) return;
  mlp_free(h->mlp);
  phasenet_free(h->pn);
  linear_free(h->transform); linear_free(h->xproj[0]); linear_free(h->xproj[1]); linear_free(h->classifier);
  for (int l = 0; l < 2; ++l) { cudaFree(h->whhT[l]); cudaFree(h->bhh[l]); }
  delete h;
}

extern "C" int mimamo_head_create(const mimamo_tensor_desc* tensors, int32_t n_tensors, int32_t num_phase,
                                  mimamo_head** head_out) {
  MM_REQUIRE(tensors && head_out && n_tensors > 0, MIMAMO_E_VALUE, "null argument");
  MM_REQUIRE(num_phase >= 1 && num_phase <= 32, MIMAMO_E_RUNTIME, "num_phase must be in [1,32] (2*num_phase phase channels), got %d", num_phase);
  TensorTable T{tensors, n_tensors};
  mimamo_head* h = new mimamo_head();
  h->num_phase = num_phase;
  h->device = current_device();
  int rc = mlp_init(T, "mlp.mlp.", h->mlp);
  if (!rc) rc = phasenet_init(T, "phasenet.", 2 * num_phase, h->pn);
  if (!rc) rc = make_linear(T, "transform.0", "transform.2", 256, 512, 1, false, h->transform);
  if (!rc) rc = make_linear(T, "classifier.1", "classifier.2", 2, 256, 0, true, h->classifier);
  for (int l = 0; l < 2 && !rc; ++l) {
    // stack both directions' input projections into one [768][256] linear; transpose W_hh
    std::vector<float> wih((size_t)768 * 256), bih(768), whhT((size_t)2 * 128 * 384), bhh(768);
    for (int d = 0; d < 2 && !rc; ++d) {
      char sfx[32];
      snprintf(sfx, sizeof(sfx), "_l%d%s", l, d ? "_reverse" : "");
      const float* wi = T.get(std::string("rnns.weight_ih") + sfx, 384 * 256);
      const float* wh = T.get(std::string("rnns.weight_hh") + sfx, 384 * 128);
      const float* bi = T.get(std::string("rnns.bias_ih") + sfx, 384);
      const float* bh = T.get(std::string("rnns.bias_hh") + sfx, 384);
      if (!wi || !wh || !bi || !bh) { rc = MIMAMO_E_VALUE; break; }
      memcpy(&wih[(size_t)d * 384 * 256], wi, sizeof(float) * 384 * 256);
      memcpy(&bih[d * 384], bi, sizeof(float) * 384);
      memcpy(&bhh[d * 384], bh, sizeof(float) * 384);
      for (int g = 0; g < 384; ++g)
        for (int k = 0; k < 128; ++k) whhT[((size_t)d * 128 + k) * 384 + g] = wh[(size_t)g * 128 + k];
    }
    if (!rc) rc = linear_init(h->xproj[l], wih.data(), bih.data(), 768, 256, 0, nullptr, nullptr, nullptr, nullptr);
    if (!rc) rc = upload(&h->whhT[l], whhT.data(), whhT.size());
    if (!rc) rc = upload(&h->bhh[l], bhh.data(), bhh.size());
  }
  if (rc) { mimamo_head_destroy(h); return rc; }
  *head_out = h;
  return MIMAMO_OK;
}

namespace {
struct HeadLayout { size_t feat, f2, xp, y0, y1, mlp_tmp, pn, total; };
HeadLayout head_layout(const mimamo_head* h, int M) {
  HeadLayout L;
  size_t cur = 0;
  auto take = [&](size_t bytes) { size_t at = cur; cur += align_up(bytes, 1024); return at; };
  L.feat = take((size_t)M * 512 * 4);
  L.f2 = take((size_t)M * 256 * 4);
  L.xp = take((size_t)M * 768 * 4);
  L.y0 = take((size_t)M * 256 * 4);
  L.y1 = take((size_t)M * 256 * 4);
  L.mlp_tmp = take(mlp_tmp_floats(h->mlp, M) * 4);
  L.pn = take(phasenet_layout(h->pn, M).total);
  L.total = cur + 1024;
  return L;
}
}  // namespace

extern "C" int mimamo_head_workspace_bytes(const mimamo_head* head, int32_t bs, int32_t nf, size_t* bytes_out) {
  MM_REQUIRE(head && bytes_out && bs >= 0 && nf >= 0, MIMAMO_E_VALUE, "bad arguments");
  *bytes_out = head_layout(head, bs * nf > 0 ? bs * nf : 1).total;
  return MIMAMO_OK;
}

static int head_forward_impl(const mimamo_head* h, const float* phase_0, const float* phase_1, const uint16_t* a0_nhwc, int a0_pitch,
                             uint16_t* cat_nhwc, const float* rgb, int32_t bs, int32_t nf, float* out, void* workspace,
                             size_t workspace_bytes, void* stream_) {
  const int M = bs * nf;
  if (M == 0) return MIMAMO_OK;
  MM_CHECK_DEVICE(h->device);
  cudaStream_t s = (cudaStream_t)stream_;
  const HeadLayout L = head_layout(h, M);
  MM_REQUIRE(workspace && workspace_bytes >= L.total, MIMAMO_E_VALUE, "workspace too small: need %zu bytes", L.total);
  char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  float* feat = (float*)(ws + L.feat); float* f2 = (float*)(ws + L.f2); float* xp = (float*)(ws + L.xp);
  float* y0 = (float*)(ws + L.y0); float* y1 = (float*)(ws + L.y1);
  // spatial stream: MLP over the ResNet50 features -> feat[:, 0:256]
  int rc = mlp_run(h->mlp, rgb, M, feat, 512, (float*)(ws + L.mlp_tmp), s);
  // temporal stream: PhaseNet -> feat[:, 256:512]
  if (!rc) rc = phasenet_run(h->pn, phase_0, phase_1, M, feat + 256, 512, ws + L.pn, s, a0_nhwc, a0_pitch, cat_nhwc);
  // fusion + recurrence over dim 0 (= bs; the nf frames are the GRU batch)
  if (!rc) rc = linear_forward(h->transform, feat, 512, M, f2, 256, s);
  if (!rc) rc = linear_forward(h->xproj[0], f2, 256, M, xp, 768, s);
  if (!rc) rc = gru_layer(xp, h->whhT[0], h->bhh[0], bs, nf, 128, y0, s);
  if (!rc) rc = linear_forward(h->xproj[1], y0, 256, M, xp, 768, s);
  if (!rc) rc = gru_layer(xp, h->whhT[1], h->bhh[1], bs, nf, 128, y1, s);
  if (!rc) rc = linear_forward(h->classifier, y1, 256, M, out, 2, s);
  return rc;
}

extern "C" int mimamo_head_forward(const mimamo_head* h, const float* phase_0, const float* phase_1, const float* rgb,
                                   int32_t bs, int32_t nf, float* out, void* workspace, size_t workspace_bytes, void* stream_) {
  MM_REQUIRE(h && phase_0 && phase_1 && rgb && out && bs >= 0 && nf >= 0, MIMAMO_E_VALUE, "bad arguments");
  return head_forward_impl(h, phase_0, phase_1, nullptr, 0, nullptr, rgb, bs, nf, out, workspace, workspace_bytes, stream_);
}

// Same forward, fed by mimamo_pyr_phase_indexed_nhwc16: the phase differences arrive as the fp16 NHWC operands PhaseNet's
// first convolution and its skip concatenation read (phase_tail.cu), so no fp32 NCHW phase tensor is written or transposed.
extern "C" int mimamo_head_forward_nhwc16(const mimamo_head* h, const void* phase0_nhwc, int32_t phase0_pitch, void* cat_nhwc,
                                          const float* rgb, int32_t bs, int32_t nf, float* out, void* workspace,
                                          size_t workspace_bytes, void* stream_) {
  MM_REQUIRE(h && phase0_nhwc && cat_nhwc && rgb && out && bs >= 0 && nf >= 0, MIMAMO_E_VALUE, "bad arguments");
  MM_REQUIRE(phase0_pitch % 8 == 0 && phase0_pitch >= h->pn.cin0, MIMAMO_E_VALUE, "phase_0 pitch must be a multiple of 8 and hold %d channels", h->pn.cin0);
  return head_forward_impl(h, nullptr, nullptr, reinterpret_cast<const uint16_t*>(phase0_nhwc), phase0_pitch,
                           reinterpret_cast<uint16_t*>(cat_nhwc), rgb, bs, nf, out, workspace, workspace_bytes, stream_);
}

// ---- the two streams on their own: MLP.forward / PhaseNet.forward (api/mimamo_net.py:22-26,79-95) ----
extern "C" void mimamo_mlp_destroy(mimamo_mlp* m) { if (m) { mlp_free(m->p); delete m; } }
extern "C" int mimamo_mlp_create(const mimamo_tensor_desc* tensors, int32_t n_tensors, mimamo_mlp** out) {
  MM_REQUIRE(tensors && out && n_tensors > 0, MIMAMO_E_VALUE, "null argument");
  TensorTable T{tensors, n_tensors};
  mimamo_mlp* m = new mimamo_mlp();
  const int rc = mlp_init(T, "mlp.", m->p);
  if (rc) { mimamo_mlp_destroy(m); return rc; }
  *out = m;
  return MIMAMO_OK;
}
extern "C" int mimamo_mlp_workspace_bytes(const mimamo_mlp* m, int32_t rows, size_t* bytes_out) {
  MM_REQUIRE(m && bytes_out && rows >= 0, MIMAMO_E_VALUE, "bad arguments");
  *bytes_out = mlp_tmp_floats(m->p, rows) * 4 + 256;
  return MIMAMO_OK;
}
extern "C" int mimamo_mlp_in_features(const mimamo_mlp* m) { return m ? m->p.in_f : 0; }
extern "C" int mimamo_mlp_forward(const mimamo_mlp* m, const float* x, int32_t rows, float* out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  MM_REQUIRE(m && x && out && rows >= 0, MIMAMO_E_VALUE, "bad arguments");
  if (rows == 0) return MIMAMO_OK;
  size_t need = 0;
  mimamo_mlp_workspace_bytes(m, rows, &need);
  MM_REQUIRE(workspace && workspace_bytes >= need, MIMAMO_E_VALUE, "workspace too small: need %zu bytes", need);
  float* tmp = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  return mlp_run(m->p, x, rows, out, 256, tmp, (cudaStream_t)stream);
}

extern "C" void mimamo_phasenet_destroy(mimamo_phasenet* n) { if (n) { phasenet_free(n->p); delete n; } }
extern "C" int mimamo_phasenet_create(const mimamo_tensor_desc* tensors, int32_t n_tensors, int32_t input_size, int32_t num_channels,
                                      mimamo_phasenet** out) {
  MM_REQUIRE(tensors && out && n_tensors > 0, MIMAMO_E_VALUE, "null argument");
  TensorTable T{tensors, n_tensors};
  mimamo_phasenet* n = new mimamo_phasenet();
  const int rc = phasenet_init(T, "", num_channels, n->p, input_size);
  if (rc) { mimamo_phasenet_destroy(n); return rc; }
  *out = n;
  return MIMAMO_OK;
}
extern "C" int mimamo_phasenet_workspace_bytes(const mimamo_phasenet* n, int32_t rows, size_t* bytes_out) {
  MM_REQUIRE(n && bytes_out && rows >= 0, MIMAMO_E_VALUE, "bad arguments");
  *bytes_out = phasenet_layout(n->p, rows > 0 ? rows : 1).total + align_up((size_t)(rows > 0 ? rows : 1) * 256 * 4, 1024) + 2048;
  return MIMAMO_OK;
}
// feature != 0: out f32[rows,256] (the fc stack); feature == 0: out f32[rows,1] (+ Linear(256,1) + BatchNorm1d(1))
extern "C" int mimamo_phasenet_forward(const mimamo_phasenet* n, const float* phase_0, const float* phase_1, int32_t rows,
                                       int32_t feature, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  MM_REQUIRE(n && phase_0 && phase_1 && out && rows >= 0, MIMAMO_E_VALUE, "bad arguments");
  if (rows == 0) return MIMAMO_OK;
  MM_REQUIRE(feature || n->p.has_cls, MIMAMO_E_VALUE, "this PhaseNet was created without classifier weights");
  size_t need = 0;
  mimamo_phasenet_workspace_bytes(n, rows, &need);
  MM_REQUIRE(workspace && workspace_bytes >= need, MIMAMO_E_VALUE, "workspace too small: need %zu bytes", need);
  char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  const size_t pn_bytes = phasenet_layout(n->p, rows).total;
  float* feat = reinterpret_cast<float*>(ws + pn_bytes);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = phasenet_run(n->p, phase_0, phase_1, rows, feature ? out : feat, 256, ws, s);
  if (!rc && !feature) rc = linear_forward(n->p.cls, feat, 256, rows, out, 1, s);
  return rc;
}
