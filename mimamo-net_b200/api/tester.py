"""Drop-in for api/tester.py: the end-to-end orchestrator, plus a device-resident clip path.

`Tester.test` / `test_on_dataloader` / `phase_diff_output` keep the reference's signatures and
return types (api/tester.py:53-139).  `infer_clips` is the B200 fast path used by bench.py and
the multi-GPU driver: gray windows + RGB frames go in, valence/arousal comes out, and the ResNet
features never leave the device (no .npy round trip).
"""
import os

import numpy as np
import torch

from phase_difference_extractor import Phase_Difference_Extractor
from mimamo_net import Two_Stream_RNN
from steerable.utils import get_device

device = get_device()          # import-time default, like the reference; methods re-query the current device


class Tester(object):
    def __init__(self,
                 model_path,
                 batch_size,
                 workers=0,
                 save_size=112, nomask=True, grey=False, quiet=True,
                 tracked_vid=False, noface_save=False,
                 OpenFace_exe='OpenFace/build/bin/FeatureExtraction',
                 benchmark_dir='pytorch-benchmarks', model_name='resnet50_ferplus_dag',
                 feature_layer='pool5_7x7_s1',
                 num_phase=12, phase_size=48,
                 length=64, stride=64,
                 height=4, nbands=2, scale_factor=2,
                 extract_level=[1, 2],
                 resnet_model=None, head_state_dict=None):
        '''Reference arguments (:15-33) plus two optional in-memory weight sources:
        resnet_model (nn.Module / state_dict) and head_state_dict.  When both are given no files,
        OpenFace binary or third-party checkpoint are needed (synthetic runs).'''
        from resnet50_extractor import Resnet50_Extractor
        self.batch_size = batch_size
        self.workers = workers
        self.num_phase = num_phase
        self.phase_size = phase_size
        self.length = length
        self.stride = stride
        synthetic = resnet_model is not None and head_state_dict is not None
        if synthetic:
            self.video_processor = None
        else:
            from video_processor import Video_Processor
            self.video_processor = Video_Processor(save_size, nomask, grey, quiet, tracked_vid, noface_save, OpenFace_exe)
        self.resnet50_extractor = Resnet50_Extractor(benchmark_dir, model_name, feature_layer, model=resnet_model)
        self.phase_difference_extractor = Phase_Difference_Extractor(height, nbands, scale_factor, extract_level, not quiet)
        self.model = Two_Stream_RNN(num_phase=num_phase)
        if head_state_dict is None:
            assert os.path.exists(model_path)
            checkpoint = torch.load(model_path, map_location='cpu')
            self.model.load_state_dict(checkpoint['state_dict'])
            print("load checkpoint from {}, epoch:{}".format(model_path, checkpoint['epoch']))
        else:
            self.model.load_state_dict(head_state_dict)
        self.model.to(get_device())
        self.model.eval()
        self.label_name = ['valence', 'arousal']
        self.save_size = save_size
        self._crop_preprocessor = None
        self._window_index = {}
        self._operand_cache = {}

    # PhaseNet operands straight from the phase tail (fp16 NHWC) when the configuration allows it; MIMAMO_PHASE_FP32=1
    # forces the reference-shaped fp32 phase tensors instead (cross-check: both routes give the same bits).
    def _operand_buffers(self, n_windows, device):
        pde = self.phase_difference_extractor
        if os.environ.get('MIMAMO_PHASE_FP32') == '1' or not isinstance(pde.extract_level, list) or len(pde.extract_level) != 2:
            return None
        channels = pde.nbands * self.num_phase
        if self.num_phase % 4 != 0 or channels % 8 != 0 or channels > 64 or channels != 2 * self.num_phase or self.phase_size != 48:
            return None
        key = (n_windows, str(device))
        if key not in self._operand_cache:
            if len(self._operand_cache) > 8:
                self._operand_cache.clear()
            c = self.phase_size
            self._operand_cache[key] = (torch.empty((n_windows, c, c, channels), dtype=torch.float16, device=device),
                                        torch.zeros((n_windows, c // 2, c // 2, 128), dtype=torch.float16, device=device))
        return self._operand_cache[key]

    def _phase_streams(self, gray, idx):
        """('operands', phase0_nhwc, cat_nhwc) or ('fp32', phase_0, phase_1), each with one row per window."""
        bufs = self._operand_buffers(idx.shape[0], gray.device)
        if bufs is not None:
            ops = self.phase_difference_extractor.phasenet_operands(gray, idx, out=bufs)
            if ops is not None:
                return ('operands',) + tuple(ops)
        diffs = self.phase_difference_extractor.phase_difference_indexed(gray, idx)
        return ('fp32',) + tuple(d.view(idx.shape[0], -1, d.shape[-2], d.shape[-1]) for d in diffs)

    def _head(self, streams, rows, feats, b, f):
        """Head forward over windows `rows` (a slice, or an index tensor) of the phase streams, as b snippets of f frames."""
        kind, s0, s1 = streams
        if isinstance(rows, slice):
            s0, s1, feats = s0[rows], s1[rows], feats[rows]
        else:
            s0, s1, feats = s0.index_select(0, rows), s1.index_select(0, rows), feats.index_select(0, rows)
        feats = feats.reshape(b, f, -1)
        if kind == 'operands':
            return self.model.forward_operands(s0, s1, feats)
        return self.model([s0.reshape(b, f, *s0.shape[1:]), s1.reshape(b, f, *s1.shape[1:])], feats)

    # ------------------------------------------------------------------ reference surface
    def test(self, input_video, fast=True):
        """Reference entry point (api/tester.py:53-74): OpenFace -> per-frame features -> snippets -> predictions.
        fast=True (default) decodes the aligned crops once and runs `test_frames` -- same windows, snippets, batches and
        stitching, every transform on the device; it does not leave the reference's `<video>_pool5/%05d.npy` feature
        cache behind.  fast=False follows the reference's file-by-file route (Resnet50_Extractor.run -> .npy ->
        Snippet_Sampler -> DataLoader) on the same kernels."""
        from sampler.snippet_sampler import Snippet_Sampler
        if self.video_processor is None:
            raise RuntimeError('this Tester was built from in-memory weights without OpenFace; use test_frames / infer_crops')
        video_name = os.path.basename(input_video).split('.')[0]
        opface_output_dir = os.path.join(os.path.dirname(input_video), video_name + "_opface")
        self.video_processor.process(input_video, opface_output_dir)
        return self.test_aligned(opface_output_dir, video_name, os.path.join(os.path.dirname(input_video), video_name + "_pool5"), fast)

    def test_aligned(self, opface_output_dir, video_name, feature_dir=None, fast=True):
        """`test` from the OpenFace output directory on (reference :60-74): <dir>/<video>_aligned/frame_det_00_%06d.bmp ->
        {video: DataFrame}.  fast=True decodes the crops once and runs everything on the device; fast=False is the
        reference's file route (features to <feature_dir>/%05d.npy, PIL samplers, DataLoader).  Both give the same
        numbers: the device transforms are bit-exact with PIL and every kernel is independent of the batch composition."""
        from sampler.snippet_sampler import Snippet_Sampler
        if fast:
            return self.test_frames(load_aligned_crops(opface_output_dir, video_name), video_name)
        if feature_dir is None:
            feature_dir = os.path.join(os.path.dirname(os.path.abspath(opface_output_dir)), video_name + "_pool5")
        self.resnet50_extractor.run(opface_output_dir, feature_dir, video_name=video_name)
        dataset = Snippet_Sampler(video_name, opface_output_dir, feature_dir, annot_dir=None,
                                  label_name='valence_arousal', test_mode=True, num_phase=self.num_phase,
                                  phase_size=self.phase_size, length=self.length, stride=self.stride)
        data_loader = torch.utils.data.DataLoader(dataset, batch_size=self.batch_size, num_workers=self.workers,
                                                  pin_memory=True)
        return self.test_on_dataloader(data_loader, self.model)

    def test_on_dataloader(self, dataloader, model, train_mean=None, train_std=None):
        import pandas as pd
        model.eval()
        device = get_device()
        names, preds, ranges = [], [], []
        for phase_f, rgb_f, label, rng, name in dataloader:
            with torch.no_grad():
                phase_f = phase_f.float().to(device, non_blocking=True)
                rgb_f = torch.as_tensor(rgb_f).float().to(device, non_blocking=True)
                phase_0, phase_1 = self.phase_diff_output(phase_f, self.phase_difference_extractor)
                output = model([phase_0, phase_1], rgb_f)
            names.append(np.asarray(name))
            ranges.append(np.asarray(rng))
            preds.append(output.cpu().numpy())
        names = np.concatenate(names, axis=0)
        preds = np.concatenate(preds, axis=0)
        ranges = np.concatenate(ranges, axis=0)
        if train_mean is not None and train_std is not None:
            # the reference calls an undefined `correct(...)` here (api/tester.py:99-101) and dies with a NameError;
            # there is no reference semantics to reproduce, so say so instead of inventing a rescaling
            raise NotImplementedError("train_mean / train_std: the reference's `correct` is undefined (api/tester.py:101)")
        return {video: pd.DataFrame(data=arr, columns=self.label_name)
                for video, arr in stitch_predictions(names, ranges, preds).items()}

    def phase_diff_output(self, phase_batch, steerable_pyramid):
        """(bs, frames, T, W, H) gray windows -> [phase_0 (bs,frames,nb*(T-1),W,H), phase_1 (.., W/2, H/2)]."""
        sp = steerable_pyramid
        bs, num_frames, num_phases, W, H = phase_batch.size()
        diffs = sp.phase_difference(phase_batch.view(bs * num_frames, num_phases, W, H))
        assert isinstance(diffs, list)
        outs = [d.view(bs, num_frames, -1, d.shape[-2], d.shape[-1]) for d in diffs]
        return outs[0], outs[1]

    # ------------------------------------------------------------------ B200 clip path
    def infer_clips(self, gray_windows, rgb_frames):
        """gray_windows (B, F, T, S, S) float32 in [0,1]; rgb_frames (B*F, 3, 224, 224) float32
        (0-255 minus mean).  Both on the GPU.  Returns (B, F, 2) = [valence, arousal]; one GRU
        sequence of length B, exactly as one reference forward with batch B."""
        with torch.no_grad():
            b, f = gray_windows.shape[0], gray_windows.shape[1]
            phase_0, phase_1 = self.phase_diff_output(gray_windows, self.phase_difference_extractor)
            feats = self.resnet50_extractor.features(rgb_frames).view(b, f, 2048)
            return self.model([phase_0, phase_1], feats)


    def infer_clips_host(self, gray_windows, rgb_frames, copy_chunk=256, to_host=True):
        """Same as infer_clips, but from HOST tensors (pinned for true overlap): the RGB batch -- 83 % of
        the input bytes -- is streamed to the device in chunks on a copy stream, double buffered, while
        the pyramid and the ResNet50 of earlier chunks run.  Returns a CPU tensor (B, F, 2) (or the device
        tensor when to_host=False, e.g. to feed a collective)."""
        device = get_device()
        main = torch.cuda.current_stream(device)
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device)
            self._rgb_bufs = [torch.empty((copy_chunk, 3, 224, 224), dtype=torch.float32, device=device) for _ in range(2)]
        copy, bufs = self._copy_stream, self._rgb_bufs
        b, f = gray_windows.shape[0], gray_windows.shape[1]
        n = rgb_frames.shape[0]
        spans = [(s, min(n, s + copy_chunk)) for s in range(0, n, copy_chunk)]
        ready = [torch.cuda.Event() for _ in spans]
        free = [torch.cuda.Event() for _ in spans]
        copy.wait_stream(main)                                   # buffers may still be in use by an earlier call

        def issue(k):
            s, e = spans[k]
            with torch.cuda.stream(copy):
                if k >= 2:
                    copy.wait_event(free[k - 2])
                bufs[k % 2][:e - s].copy_(rgb_frames[s:e], non_blocking=True)
                ready[k].record(copy)

        with torch.no_grad():
            # the gray windows go first: nothing can start before they land, and the pyramid then
            # overlaps the first RGB chunks
            with torch.cuda.stream(copy):
                gray = gray_windows.to(device, non_blocking=True)
                gray_ready = torch.cuda.Event()
                gray_ready.record(copy)
            for k in range(min(2, len(spans))):
                issue(k)
            main.wait_event(gray_ready)
            gray.record_stream(main)
            phase_0, phase_1 = self.phase_diff_output(gray, self.phase_difference_extractor)
            feats = torch.empty((n, 2048), dtype=torch.float32, device=device)
            for k, (s, e) in enumerate(spans):
                main.wait_event(ready[k])
                feats[s:e] = self.resnet50_extractor.features(bufs[k % 2][:e - s])
                free[k].record(main)
                if k + 2 < len(spans):
                    issue(k + 2)
            out = self.model([phase_0, phase_1], feats.view(b, f, 2048))
        return out.cpu() if to_host else out


    # ------------------------------------------------------------------ B200 crop path (uint8 in)
    def crop_preprocessor(self):
        if self._crop_preprocessor is None:
            from utils.crop_preprocessor import Crop_Preprocessor
            meta = self.resnet50_extractor.meta
            self._crop_preprocessor = Crop_Preprocessor(self.save_size, self.phase_size, 256, meta['imageSize'][0],
                                                        meta['mean'])
        return self._crop_preprocessor

    def clip_window_index(self, n_clips, n_frames, device):
        """(n_clips*n_frames, num_phase+1) int32: frame ids of every window, each clip being its own
        video for the clamp rule (api/sampler/snippet_sampler.py:144-152)."""
        key = (n_clips, n_frames, str(device))
        if key not in self._window_index:
            from sampler.snippet_sampler import window_index
            idx = window_index(0, n_frames, n_frames, self.num_phase)                    # (F, T)
            idx = idx[None, :, :] + (torch.arange(n_clips) * n_frames)[:, None, None]
            self._window_index[key] = idx.reshape(n_clips * n_frames, -1).to(device=device, dtype=torch.int32)
        return self._window_index[key]

    def infer_crops(self, crops):
        """crops (B, F, S, S, 3) uint8 on the GPU: B clips of F aligned face crops as OpenFace writes them
        (112x112 RGB).  Everything the reference's samplers do on the host with PIL -- convert('L'),
        LANCZOS 48x48, /255, the 13-frame clamp windows, Resize 256 / CenterCrop 224 / mean -- runs on
        the device, bit-exact; each distinct frame goes through the pyramid once.  Returns (B, F, 2)
        = [valence, arousal], one GRU sequence of length B like one reference forward with batch B."""
        with torch.no_grad():
            b, f = crops.shape[0], crops.shape[1]
            pre = self.crop_preprocessor()
            flat = crops.reshape(b * f, crops.shape[2], crops.shape[3], crops.shape[4])
            gray = pre.gray(flat)                                                       # (B*F, 48, 48)
            idx = self.clip_window_index(b, f, crops.device)
            streams = self._phase_streams(gray, idx)
            feats = self.resnet50_extractor.features_from_crops(flat, pre)
            return self._head(streams, slice(0, b * f), feats, b, f)

    def infer_crops_host(self, crops, to_host=True, parts=4):
        """infer_crops from a HOST uint8 tensor (pinned for an asynchronous copy): 37.6 KB per frame cross
        PCIe instead of the 722 KB of fp32 windows + RGB the reference's DataLoader ships.  The clips are
        copied in `parts` groups on a copy stream; the gray / pyramid / phase stage of a group (clips are
        independent there) starts as soon as its crops have landed, so only the first group's copy is exposed.
        ResNet50 and the head then run once over the whole batch (bit-identical to infer_crops)."""
        device = get_device()
        b, f = crops.shape[0], crops.shape[1]
        main = torch.cuda.current_stream(device)
        if getattr(self, '_crop_copy_stream', None) is None:
            self._crop_copy_stream = torch.cuda.Stream(device)
        copy = self._crop_copy_stream
        parts = max(1, min(parts, b))
        bounds = [(b * k) // parts for k in range(parts + 1)]
        with torch.no_grad():
            dev = torch.empty(crops.shape, dtype=torch.uint8, device=device)
            copy.wait_stream(main)
            landed = []
            with torch.cuda.stream(copy):
                for k in range(parts):
                    dev[bounds[k]:bounds[k + 1]].copy_(crops[bounds[k]:bounds[k + 1]], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy)
                    landed.append(ev)
            dev.record_stream(copy)
            pre = self.crop_preprocessor()
            pde = self.phase_difference_extractor
            bufs = self._operand_buffers(b * f, device)
            phase = None
            for k in range(parts):
                lo, hi = bounds[k], bounds[k + 1]
                if hi == lo:
                    continue
                main.wait_event(landed[k])
                flat = dev[lo:hi].reshape((hi - lo) * f, dev.shape[2], dev.shape[3], dev.shape[4])
                idx = self.clip_window_index(hi - lo, f, device)
                if bufs is not None and pde.phasenet_operands(pre.gray(flat), idx, out=(bufs[0][lo * f:hi * f], bufs[1][lo * f:hi * f])) is not None:
                    continue
                bufs = None                                                   # configuration without the fp16 operand path
                if phase is None:
                    nb, T = pde.nbands, self.num_phase + 1
                    sizes = [self.phase_size, self.phase_size // 2]
                    phase = [torch.empty((b, f, nb * (T - 1), c, c), dtype=torch.float32, device=device) for c in sizes]
                pde.phase_difference_indexed(pre.gray(flat), idx, out=[p[lo:hi] for p in phase])
            flat = dev.reshape(b * f, dev.shape[2], dev.shape[3], dev.shape[4])
            feats = self.resnet50_extractor.features_from_crops(flat, pre)
            if bufs is not None:
                out = self.model.forward_operands(bufs[0], bufs[1], feats.view(b, f, 2048))
            else:
                out = self.model([phase[0], phase[1]], feats.view(b, f, 2048))
        return out.cpu() if to_host else out


    # ------------------------------------------------------------------ B200 video path (uint8 in, DataFrame out)
    def predict_frames(self, frames):
        """One video's aligned face crops, uint8 (n, S, S, 3) on the GPU -> (n, 2) float32 CUDA predictions with
        exactly the semantics of `test` (api/tester.py:53-121): 13-frame windows clamped to the VIDEO
        (snippet_sampler.py:144-152), 64-frame snippets plus the overlapping tail snippet (:116-126), the video's
        snippets forwarded as one batch of at most `batch_size` (the GRU runs over the snippets of a batch) and
        later snippets overwriting the overlap when stitched (:104-118).  Every frame is preprocessed, transformed
        and pushed through ResNet50 once, however many windows / snippets it belongs to."""
        from sampler.snippet_sampler import snippet_ranges, window_index
        n = frames.shape[0]
        if n == 0:
            raise ValueError("number of frames of video should not be zero.")
        device = frames.device
        with torch.no_grad():
            pre = self.crop_preprocessor()
            idx = window_index(0, n, n, self.num_phase).to(device=device, dtype=torch.int32)
            streams = self._phase_streams(pre.gray(frames), idx)
            feats = self.resnet50_extractor.features_from_crops(frames, pre)
            ranges = snippet_ranges(n, self.length, self.stride)
            out = torch.zeros((n, len(self.label_name)), dtype=torch.float32, device=device)
            for b0 in range(0, len(ranges), self.batch_size):             # DataLoader batches, in order
                batch = ranges[b0:b0 + self.batch_size]
                rows = torch.cat([torch.arange(s, e, device=device) for s, e in batch])
                L = batch[0][1] - batch[0][0]
                pred = self._head(streams, rows, feats, len(batch), L)
                for k, (s, e) in enumerate(batch):                        # in order: the tail snippet overwrites its overlap
                    out[s:e] = pred[k]
        return out

    def predict_videos(self, videos, group_frames=4096):
        """Several videos at once: list of uint8 (n_v, S, S, 3) device tensors -> list of (n_v, 2) predictions, each
        bit-identical to predict_frames(video).  The per-frame work (preprocessing, pyramid, ResNet50 -- nothing in it
        couples frames of different videos, and the kernels are independent of the batch composition) runs over groups
        of whole videos of about `group_frames` frames so the tensor-core passes stay full; the head then runs per
        video, one forward per DataLoader batch of that video's snippets, exactly as `test` does
        (api/tester.py:65-73: a batch never mixes videos)."""
        from sampler.snippet_sampler import snippet_ranges, window_index
        out = [None] * len(videos)
        i = 0
        with torch.no_grad():
            pre = self.crop_preprocessor()
            while i < len(videos):
                j, total = i, 0
                while j < len(videos) and (j == i or total + videos[j].shape[0] <= group_frames):
                    if videos[j].shape[0] == 0:
                        raise ValueError("number of frames of video should not be zero.")
                    total += videos[j].shape[0]
                    j += 1
                group = videos[i:j]
                device = group[0].device
                frames = group[0] if len(group) == 1 else torch.cat(group, 0)
                offs = [0]
                for v in group:
                    offs.append(offs[-1] + v.shape[0])
                key = ('videos', tuple(v.shape[0] for v in group), str(device))
                if key not in self._window_index:
                    if len(self._window_index) > 64:
                        self._window_index.clear()
                    idx = torch.cat([window_index(0, v.shape[0], v.shape[0], self.num_phase) + o
                                     for v, o in zip(group, offs)])
                    self._window_index[key] = idx.to(device=device, dtype=torch.int32)
                idx = self._window_index[key]
                streams = self._phase_streams(pre.gray(frames), idx)
                feats = self.resnet50_extractor.features_from_crops(frames, pre)
                # Head.  The GRU recurs over the snippets of ONE video's batch and treats the frames as independent batch
                # rows (api/mimamo_net.py:119,138-139), so videos of equal length whose snippets form a single DataLoader batch
                # can share a forward: their frames simply become more GRU batch rows -- (S snippets, V*L frames).  Nothing in
                # the head couples rows, so this is bit-identical to one forward per video.
                g = 0
                while g < len(group):
                    n = group[g].shape[0]
                    ranges = snippet_ranges(n, self.length, self.stride)
                    h = g + 1
                    if len(ranges) <= self.batch_size:
                        while h < len(group) and group[h].shape[0] == n and (h - g) < 32:
                            h += 1
                    V = h - g
                    if V > 1:
                        L = ranges[0][1] - ranges[0][0]
                        base = torch.tensor([[offs[g + v] + s for v in range(V)] for s, _ in ranges], device=device)      # (S, V)
                        rows = (base[:, :, None] + torch.arange(L, device=device)[None, None, :]).reshape(-1)
                        pred = self._head(streams, rows, feats, len(ranges), V * L).view(len(ranges), V, L, -1)
                        for v in range(V):
                            pred_v = torch.zeros((n, len(self.label_name)), dtype=torch.float32, device=device)
                            for q, (s0, e0) in enumerate(ranges):                 # in order: the tail snippet overwrites its overlap
                                pred_v[s0:e0] = pred[q, v]
                            out[i + g + v] = pred_v
                        g = h
                        continue
                    o = offs[g]
                    pred_v = torch.zeros((n, len(self.label_name)), dtype=torch.float32, device=device)
                    for b0 in range(0, len(ranges), self.batch_size):
                        batch = ranges[b0:b0 + self.batch_size]
                        L = batch[0][1] - batch[0][0]
                        whole = all(e - s == L and s == batch[0][0] + q * L for q, (s, e) in enumerate(batch))
                        if whole:                                             # consecutive snippets: plain views, no gather
                            rows = slice(o + batch[0][0], o + batch[-1][1])
                        else:
                            rows = torch.cat([torch.arange(o + s, o + e, device=device) for s, e in batch])
                        pred = self._head(streams, rows, feats, len(batch), L)
                        for q, (s, e) in enumerate(batch):
                            pred_v[s:e] = pred[q]
                    out[i + g] = pred_v
                    g += 1
                i = j
        return out

    def test_frames(self, frames, video_name='video'):
        """`test` without OpenFace / files: frames = decoded aligned crops of one video, uint8 (n, S, S, 3), host or
        device.  Returns {video_name: DataFrame(columns=['valence', 'arousal'])} like the reference."""
        import pandas as pd
        frames = torch.as_tensor(frames)
        pred = self.predict_frames(frames.to(get_device(), non_blocking=True))
        return {video_name: pd.DataFrame(data=pred.cpu().numpy().astype(np.float64), columns=self.label_name)}


def load_aligned_crops(opface_output_dir, video_name):
    """Decoded face crops of an OpenFace output directory, uint8 (n, S, S, 3) RGB in OpenFace frame order
    (<dir>/<video>_aligned/frame_det_00_%06d.bmp; the order Image_Sampler / Snippet_Sampler use:
    api/sampler/image_sampler.py:104-107, snippet_sampler.py:100-103).  Host I/O only."""
    import glob
    from PIL import Image
    paths = glob.glob(os.path.join(opface_output_dir, video_name + "_aligned", '*.bmp'))
    paths = sorted(paths, key=lambda x: os.path.basename(x).split(".")[0].split("_")[-1])
    if len(paths) == 0:
        raise ValueError("number of frames of video {} should not be zero.".format(video_name))
    frames = np.stack([np.asarray(Image.open(p).convert('RGB'), dtype=np.uint8) for p in paths])
    out = torch.from_numpy(frames)
    return out.pin_memory() if torch.cuda.is_available() else out


def stitch_predictions(names, ranges, preds):
    """Per-video stitching of snippet predictions (api/tester.py:104-118): later snippets overwrite the
    overlap; asserts full coverage [0, max_len)."""
    out = {}
    for video in names:
        if video in out:
            continue
        mask = names == video
        v_ranges, v_preds = ranges[mask], preds[mask]
        max_len = max(r[-1] for r in v_ranges)
        arr = np.zeros((max_len, preds.shape[-1]))
        lo, hi = 0, 0
        for (start, end), p in zip(v_ranges, v_preds):
            arr[start:end, :] = p
            lo, hi = min(lo, start), max(hi, end)
        assert lo == 0 and hi == max_len
        out[video] = arr
    return out
