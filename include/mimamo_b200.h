/*
 * libmimamo_b200.so -- C ABI of the B200-native MIMAMO-Net per-window inference hot path.
 *
 * The reference (wtomin/MIMAMO-Net) is 100 % Python over PyTorch library calls and has no
 * FFI of its own; the drop-in boundary is its `api/` class surface.
 * The Python classes under `mimamo-net_b200/api/` keep that surface and call the entry points
 * below through ctypes.  Each entry point names the reference code it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - every call returns 0 on success, <0 on error; mimamo_last_error() returns a thread-local
 *     message which the Python layer re-raises as the matching exception type
 *     (MIMAMO_E_VALUE -> ValueError, MIMAMO_E_RUNTIME/CUDA -> RuntimeError);
 *   - all data pointers are DEVICE pointers unless the name ends in `_host`; the caller owns
 *     every buffer (outputs and workspaces are never allocated here); plans / nets own only
 *     their immutable tables and weights;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and the calls
 *     never synchronise the device;
 *   - plain pointers and sizes only: no torch / C++ types cross this boundary.
 */
#ifndef MIMAMO_B200_H_
#define MIMAMO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIMAMO_OK          0
#define MIMAMO_E_VALUE    -1   /* bad argument (ValueError)                     */
#define MIMAMO_E_RUNTIME  -2   /* unsupported configuration (RuntimeError)      */
#define MIMAMO_E_CUDA     -3   /* CUDA runtime / driver failure (RuntimeError)  */

#define MIMAMO_MAX_LEVELS  8

int         mimamo_abi_version(void);
const char* mimamo_last_error(void);
/* number of kernels this library has launched on the calling process so far (bench.py's
 * `gpu_launches`). */
uint64_t    mimamo_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * P0 + P1: complex steerable pyramid of mirror-extended frames.
 * Replaces symmetric_extension_batch (api/utils/phase_utils.py:116-129),
 * SCFpyr_PyTorch.build/_build_levels (api/steerable/SCFpyr_PyTorch.py:70-208) and the
 * stack/permute/crop in Phase_Difference_Extractor.build_pyramid
 * (api/phase_difference_extractor.py:38-87).
 * The data-independent tables are built on the host exactly like the reference builds its
 * masks (np.interp; mimamo-net_b200/api/steerable/plan_tables.py) and uploaded once.
 * ---------------------------------------------------------------------------------------- */
typedef struct mimamo_pyr_plan mimamo_pyr_plan;

typedef struct {
  int32_t c;                 /* kept crop: outputs y,x in [0,c)                              */
  int32_t h;                 /* folded frequency count                                        */
  int32_t hp, cp;            /* h, c padded to multiples of 8 (leading dimensions)            */
  const float*   trig_host;  /* [2][hp][cp]  cos/sin(pi k/S + 2 pi k y/s)                     */
  const float*   masks_host; /* [nbands][2 ch][2 half][hp][hp], transposed ([l][k])           */
  const int32_t* inner_sel_host; /* [2 ch][2 half]: 0 = cos table, 1 = sin table              */
} mimamo_pyr_level_desc;

int mimamo_pyr_plan_create(int32_t H, int32_t Hp, int32_t Kp, int32_t nbands,
                           const float* dct_t_host /* [Hp][Kp] */,
                           int32_t n_levels, const mimamo_pyr_level_desc* levels,
                           mimamo_pyr_plan** plan_out);
void mimamo_pyr_plan_destroy(mimamo_pyr_plan* plan);

/* frames f32[n_windows*T, H, H]  ->  coeff_out[level] f32[n_windows, nbands, T, c, c, 2]
 * (the layout Phase_Difference_Extractor.build_pyramid returns).  Frames that fit in shared
 * memory (H <= ~128) need no workspace; larger frames (e.g. 224x224) run a persistent grid over a
 * caller-provided scratch. */
int mimamo_pyr_build_workspace_bytes(const mimamo_pyr_plan* plan, int64_t n_windows, int32_t T,
                                     size_t* bytes_out);
int mimamo_pyr_build(const mimamo_pyr_plan* plan, const float* frames,
                     int64_t n_windows, int32_t T, float* const* coeff_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Full pyramid of arbitrary (un-mirrored) square images and its inverse.
 * Replaces SCFpyr_PyTorch.build / _build_levels (api/steerable/SCFpyr_PyTorch.py:70-208) -- hi0 residual, every
 * oriented band at full size, low residual -- and SCFpyr_PyTorch.reconstruct / _reconstruct_levels (:214-314).
 * Not on the inference hot path (that is mimamo_pyr_build above).  A plan is a list of output "units" in the order
 * of the reference's coeff list [hi0, level 1, ..., level L, lo]; the host supplies, per unit, its size s, the gather
 * index of its natural-order frequencies into the image's natural-order spectrum (all of the reference's
 * fftshift / centre-crop / ifftshift steps composed) and its cumulative real masks
 * (mimamo-net_b200/api/steerable/plan_tables.py::full_pyramid_tables).
 * ---------------------------------------------------------------------------------------- */
typedef struct mimamo_scf_plan mimamo_scf_plan;
typedef struct {
  int32_t s;                      /* unit size (s x s)                                              */
  int32_t planes;                 /* 1 for hi0 / lo, nbands for a level                             */
  int32_t is_real;                /* 1: the reference keeps the real part (hi0, lo)                 */
  int32_t twist_build;            /* multiply by (-i)^twist when building (SCFpyr_PyTorch.py:64)    */
  int32_t twist_recon;            /* ... when reconstructing (:65)                                  */
  const int32_t* src_index_host;  /* [s]                                                            */
  const float*   build_mask_host; /* [planes][s][s], natural frequency order                        */
  const float*   recon_mask_host; /* [planes][s][s]                                                 */
} mimamo_scf_unit_desc;
int  mimamo_scf_plan_create(int32_t S, int32_t n_units, const mimamo_scf_unit_desc* units, mimamo_scf_plan** plan_out);
void mimamo_scf_plan_destroy(mimamo_scf_plan* plan);
int  mimamo_scf_workspace_bytes(const mimamo_scf_plan* plan, int64_t N, size_t* bytes_out);
/* images f32[N,S,S] -> unit_out[u]: f32[N,s,s] (real units) or f32[planes,N,s,s,2] (levels) */
int  mimamo_scf_build(const mimamo_scf_plan* plan, const float* images, int64_t N, float* const* unit_out,
                      void* workspace, size_t workspace_bytes, void* stream);
/* the inverse: unit_in as produced by mimamo_scf_build -> images_out f32[N,S,S] */
int  mimamo_scf_reconstruct(const mimamo_scf_plan* plan, const float* const* unit_in, int64_t N, float* images_out,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * P2: phase tail.  Replaces Phase_Difference_Extractor.extract
 * (api/phase_difference_extractor.py:93-134) with torch_unwrap / torch_diff /
 * amplitude_based_gaussian_blur / gaussian_kernel (api/utils/phase_utils.py:5-40,78-90,108-115).
 * coeff f32[n_maps, T, rows, cols, 2] -> out f32[n_maps, T-1, rows, cols], n_maps = bs*nbands.
 * ---------------------------------------------------------------------------------------- */
int mimamo_phase_extract_workspace_bytes(int64_t n_maps, int32_t T, int32_t rows, int32_t cols,
                                         size_t* bytes_out);
int mimamo_phase_extract(const float* coeff, int64_t n_maps, int32_t T, int32_t rows, int32_t cols,
                         float* out, void* workspace, size_t workspace_bytes, void* stream);
/* Training-side variants of the same tail, Steerable_Pyramid_Phase.extract_phase(coeff, return_phase, return_both)
 * (Aff-wild-exps/utils.py:367-418) with insert_tensors (:419-432; it fills only the first T-1 of its 2(T-1) slots,
 * reproduced):  mode 0 = phase differences (as above), mode 1 = return_phase -> out f32[n_maps, T, rows, cols]
 * (denoised phase minus its spatial mean), mode 2 = return_both -> out f32[n_maps, 2(T-1), rows, cols]. */
int mimamo_phase_extract_ex(const float* coeff, int64_t n_maps, int32_t T, int32_t rows, int32_t cols, int32_t mode,
                            float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * P3: frames -> phase-difference maps without materialising coefficients in the caller.
 * Replaces Tester.phase_diff_output (api/tester.py:122-139).
 * frames f32[n_windows*T,H,H] -> out[level] f32[n_windows, nbands*(T-1), c, c].
 * ---------------------------------------------------------------------------------------- */
int mimamo_pyr_phase_workspace_bytes(const mimamo_pyr_plan* plan, int64_t n_windows, int32_t T,
                                     size_t* bytes_out);
int mimamo_pyr_phase(const mimamo_pyr_plan* plan, const float* frames, int64_t n_windows, int32_t T,
                     float* const* out, void* workspace, size_t workspace_bytes, void* stream);

/* Clip variant of P3: every distinct frame is transformed once and
 * window w, slot t reads frame window_index[w*T + t] -- the clamp-window rule of
 * api/sampler/snippet_sampler.py:144-152, built by the caller.  Bit-identical to materialising
 * the windows and calling mimamo_pyr_phase.
 * frames f32[n_frames,H,H], window_index i32[n_windows*T] (device)
 * -> out[level] f32[n_windows, nbands*(T-1), c, c]. */
int mimamo_pyr_phase_indexed_workspace_bytes(const mimamo_pyr_plan* plan, int64_t n_frames,
                                             int64_t n_windows, int32_t T, size_t* bytes_out);
int mimamo_pyr_phase_indexed(const mimamo_pyr_plan* plan, const float* frames, int64_t n_frames,
                             const int32_t* window_index, int64_t n_windows, int32_t T,
                             float* const* out, void* workspace, size_t workspace_bytes, void* stream);

/* The same clip path feeding PhaseNet directly: level i leaves as fp16 channels band*(T-1) + t of the NHWC tensor
 * out16[i] = f16[n_windows, c_i, c_i, pitch[i]] at channel offset c_off[i] -- the operands mimamo_head_forward_nhwc16
 * reads -- instead of fp32 NCHW maps that the head would transpose (no reference counterpart: the reference's
 * nn.Conv2d takes the fp32 tensor of api/tester.py:131-138).  Needs maps up to 56x56 and (T-1) % 4 == 0; pitch and
 * offsets multiples of 4.  Workspace as mimamo_pyr_phase_indexed_workspace_bytes. */
int mimamo_pyr_phase_indexed_nhwc16(const mimamo_pyr_plan* plan, const float* frames, int64_t n_frames,
                                    const int32_t* window_index, int64_t n_windows, int32_t T,
                                    void* const* out16, const int32_t* pitch, const int32_t* c_off,
                                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Face-crop preprocessing on the device, bit-exact with the
 * reference's PIL / torchvision transforms:
 *   gray: Image.convert('L') -> Resize(gray_size, LANCZOS) -> float / 255
 *         (api/sampler/snippet_sampler.py:156-185, api/utils/data_utils.py:71-120)
 *   RGB : Resize(resize) [PIL bilinear] -> CenterCrop(crop) -> ToTensor -> x*255 -> Normalize(mean, 1)
 *         (api/utils/model_utils.py:26-40, api/sampler/image_sampler.py:118-119)
 * The tap tables are Pillow's (libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc:
 * per output index a first input index, a tap count and 22-bit fixed-point taps), built on the
 * host by api/utils/pil_tables.py.  bounds i32[out][2], kk i32[out][taps].
 * crops u8[n, src, src, 3] (RGB, HWC as OpenFace's bmp files decode).
 * ---------------------------------------------------------------------------------------- */
typedef struct mimamo_preproc mimamo_preproc;
int  mimamo_preproc_create(int32_t src, int32_t gray_size, int32_t gray_taps,
                           const int32_t* gray_bounds_host, const int32_t* gray_kk_host,
                           int32_t resize, int32_t rgb_taps,
                           const int32_t* rgb_bounds_host, const int32_t* rgb_kk_host,
                           int32_t crop, int32_t crop_off, const float* mean_host /* [3] */,
                           mimamo_preproc** plan_out);
void mimamo_preproc_destroy(mimamo_preproc* plan);
int  mimamo_preproc_geometry(const mimamo_preproc* plan, int32_t* src, int32_t* gray_size, int32_t* crop);
/* -> out f32[n, gray_size, gray_size] in [0,1] */
int  mimamo_crops_to_gray(const mimamo_preproc* plan, const uint8_t* crops, int64_t n, float* out, void* stream);
/* -> out f32[n, 3, crop, crop] = (u8/255)*255 - mean, what Image_Sampler yields */
int  mimamo_crops_to_rgb(const mimamo_preproc* plan, const uint8_t* crops, int64_t n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Convolution network engine (rows R and H of SURVEY.md section 8(a)).
 * A net is a flat table of named host tensors (the reference-keyed state_dict) folded once
 * into bf16 NHWC implicit-GEMM weights + fp32 scale/shift (eval-mode BatchNorm).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const char*    name;       /* state_dict key                                               */
  const float*   data_host;  /* fp32, contiguous, torch layout                               */
  int32_t        ndim;
  int64_t        shape[4];
} mimamo_tensor_desc;

/* R: ResNet50 pool5.  Replaces Resnet50_Extractor.get_vec (api/resnet50_extractor.py:74-83)
 * over the third-party resnet50_ferplus_dag module (api/utils/model_utils.py:65-79).
 * x f32[B,3,224,224] (0-255 scale minus mean, NCHW as the reference feeds it)
 * -> out f32[B,2048] = relu(pool5_7x7_s1). */
typedef struct mimamo_resnet50 mimamo_resnet50;
int  mimamo_resnet50_create(const mimamo_tensor_desc* tensors, int32_t n_tensors,
                            mimamo_resnet50** net_out);
void mimamo_resnet50_destroy(mimamo_resnet50* net);
int  mimamo_resnet50_workspace_bytes(const mimamo_resnet50* net, int32_t batch, size_t* bytes_out);
int  mimamo_resnet50_pool5(const mimamo_resnet50* net, const float* x, int32_t batch, float* out,
                           void* workspace, size_t workspace_bytes, void* stream);

/* Weight rounding.  The 16-bit weights are rounded so that, per output channel, the rounding residuals weighted by the
 * mean of each input channel sum to ~0 (the error component that is identical at every pixel and therefore survives
 * pool5's spatial average).  mimamo_resnet50_create calibrates the channel means on a built-in
 * synthetic batch (MIMAMO_RESNET_CALIB=0/1/2: round-to-nearest / uniform means / calibrated, default 2); this call
 * re-calibrates on the caller's images x f32[batch,3,224,224] (batch <= one pass).  No reference counterpart. */
int  mimamo_resnet50_calibrate(mimamo_resnet50* net, const float* x, int32_t batch,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Same, from uint8 face crops [B, src, src, 3]: the RGB transform above runs on the device and
 * feeds conv1 directly (workspace as mimamo_resnet50_workspace_bytes). */
int  mimamo_resnet50_pool5_crops(const mimamo_resnet50* net, const mimamo_preproc* plan,
                                 const uint8_t* crops, int32_t batch, float* out,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* H: two-stream head.  Replaces Two_Stream_RNN.forward (api/mimamo_net.py:129-143; MLP :22-26,
 * PhaseNet :79-95; GRU built without batch_first at :119, so it recurs over dim 0 = bs).
 * phase_0 f32[bs,nf,C,48,48], phase_1 f32[bs,nf,C,24,24] (C = 2*num_phase, 24 by default), rgb f32[bs,nf,2048]
 * -> out f32[bs,nf,2] = [valence, arousal]. */
typedef struct mimamo_head mimamo_head;
int  mimamo_head_create(const mimamo_tensor_desc* tensors, int32_t n_tensors, int32_t num_phase,
                        mimamo_head** head_out);
void mimamo_head_destroy(mimamo_head* head);
int  mimamo_head_workspace_bytes(const mimamo_head* head, int32_t bs, int32_t nf, size_t* bytes_out);
int  mimamo_head_forward(const mimamo_head* head, const float* phase_0, const float* phase_1,
                         const float* rgb, int32_t bs, int32_t nf, float* out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Head forward fed by mimamo_pyr_phase_indexed_nhwc16: phase0_nhwc f16[bs*nf,48,48,phase0_pitch] (the 2*num_phase level-0
 * channels; phase0_pitch a multiple of 8), cat_nhwc f16[bs*nf,24,24,128] with the level-1 channels at [64, 64+2*num_phase)
 * and ZEROS above them (channels [0,64) are scratch: conv_net[0][3] writes its output there, the skip concatenation of
 * api/mimamo_net.py:85). */
int  mimamo_head_forward_nhwc16(const mimamo_head* head, const void* phase0_nhwc, int32_t phase0_pitch, void* cat_nhwc,
                                const float* rgb, int32_t bs, int32_t nf, float* out,
                                void* workspace, size_t workspace_bytes, void* stream);

/* The two streams of the head on their own: MLP.forward (api/mimamo_net.py:22-26; keys `mlp.{1,2,5,6,...}`, any depth,
 * last width 256) and PhaseNet.forward (:27-95; keys `conv_net.*`, `fc.*`, `classifier.*`; input_size S = 48 with three
 * conv blocks, 96 or 112 with four).
 * x f32[rows, in_features] -> out f32[rows,256];  phase_0 f32[rows,C,S,S], phase_1 f32[rows,C,S/2,S/2] ->
 * out f32[rows,256] (feature != 0) or f32[rows,1] (feature == 0: + classifier Linear(256,1) + BatchNorm1d(1)). */
typedef struct mimamo_mlp mimamo_mlp;
int  mimamo_mlp_create(const mimamo_tensor_desc* tensors, int32_t n_tensors, mimamo_mlp** mlp_out);
void mimamo_mlp_destroy(mimamo_mlp* mlp);
int  mimamo_mlp_in_features(const mimamo_mlp* mlp);
int  mimamo_mlp_workspace_bytes(const mimamo_mlp* mlp, int32_t rows, size_t* bytes_out);
int  mimamo_mlp_forward(const mimamo_mlp* mlp, const float* x, int32_t rows, float* out,
                        void* workspace, size_t workspace_bytes, void* stream);
typedef struct mimamo_phasenet mimamo_phasenet;
int  mimamo_phasenet_create(const mimamo_tensor_desc* tensors, int32_t n_tensors, int32_t input_size,
                            int32_t num_channels, mimamo_phasenet** net_out);
void mimamo_phasenet_destroy(mimamo_phasenet* net);
int  mimamo_phasenet_workspace_bytes(const mimamo_phasenet* net, int32_t rows, size_t* bytes_out);
int  mimamo_phasenet_forward(const mimamo_phasenet* net, const float* phase_0, const float* phase_1, int32_t rows,
                             int32_t feature, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Per-launch CUDA-event timing of the tcgen05 GEMM kernel (bench.py's roofline leg).
 * enable=1 resets and starts recording; read after synchronising the stream. */
int mimamo_profile_gemm(int32_t enable);
int mimamo_profile_gemm_read(double* total_ms, uint64_t* launches, double* issued_flops);
/* per-launch durations in ms, launch order; returns how many were written (<= max_launches) */
int mimamo_profile_gemm_launches(float* ms_out, int32_t max_launches);

/* Test hook: one conv layer through the tcgen05 implicit-GEMM engine.
 * x bf16 NHWC [B,H,W,Cin] (Cin multiple of 8), w f32 [Cout,Cin,k,k] (torch layout),
 * scale/shift f32[Cout], optional residual bf16 NHWC of the output shape, out bf16 NHWC. */
int mimamo_conv_bf16(const void* x, int32_t B, int32_t H, int32_t W, int32_t Cin,
                     const float* w_host, const float* scale_host, const float* shift_host,
                     int32_t Cout, int32_t ksize, int32_t stride, int32_t pad, int32_t relu,
                     const void* residual, void* out, void* stream);

/* Test hook: two flat 1x1 layers in one launch (conv_chain_kernel; ResNet50 `_increase` + residual + ReLU followed by the
 * next block's `_reduce` + ReLU, api/resnet50_extractor.py:81).  x bf16 [M,K1], w1 f32 [N1,K1], residual bf16 [M,N1],
 * w2 f32 [N2,N1]; out1 bf16 [M,N1] = relu(bn1(x w1^T) + residual), out2 bf16 [M,N2] = relu(bn2(out1 w2^T)).
 * N1 a multiple of 128, N2 in {64, 128}. */
int mimamo_conv_chain_bf16(const void* x, int32_t M, int32_t K1, const float* w1_host, const float* scale1_host,
                           const float* shift1_host, int32_t N1, const void* residual, const float* w2_host,
                           const float* scale2_host, const float* shift2_host, int32_t N2, void* out1, void* out2,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MIMAMO_B200_H_ */
