"""Snippet / window index rules of api/sampler/snippet_sampler.py plus a file-backed sampler.

The index rules are part of the hot-path contract; the PIL/BMP
loading is host I/O kept only so `Tester.test` can run on an OpenFace output directory.
"""
import glob
import os

import numpy as np
import torch
import torch.utils.data as data


def snippet_ranges(n_frames, length=64, stride=64):
    """[start,end) ranges of a video's snippets (reference parse_video, :107-128): stride-`stride`
    snippets while they fit, plus one tail snippet ending at the last frame if uncovered; videos
    shorter than `length` give one snippet of their own length."""
    if n_frames < length:
        length = stride = n_frames
    ranges = []
    start = 0
    while start + length <= n_frames and start < n_frames:
        ranges.append([start, start + length])
        start += stride
    assert len(ranges) != 0, "No snippet is sampled."
    if ranges[-1][1] < n_frames:
        ranges.append([n_frames - length, n_frames])
    return ranges


def window_frame_ids(frame, n_frames, num_phase=12):
    """The num_phase+1 frame ids of a frame's temporal window, clamped to the video (:144-152)."""
    lo = frame - num_phase // 2
    return [min(max(0, lo + i), n_frames - 1) for i in range(num_phase + 1)]


def window_index(start, end, n_frames, num_phase=12):
    """LongTensor (end-start, num_phase+1) of frame ids for a snippet."""
    return torch.tensor([window_frame_ids(f, n_frames, num_phase) for f in range(start, end)], dtype=torch.long)


class Snippet_Sampler(data.Dataset):
    """Test-mode sampler over <root_path>/<video>_aligned/frame_det_00_%06d.bmp and
    <feature_path>/%05d.npy; returns (phase_images (L,T,S,S), features (L,2048), labels, [start,end], video)."""

    def __init__(self, video_name, root_path, feature_path, annot_dir=None, label_name=None, test_mode=True,
                 num_phase=12, phase_size=48, length=64, stride=64, verbose=False):
        if not test_mode:
            raise NotImplementedError('training-mode sampling (labels, augmentation) is out of scope')
        self.video_name, self.root_path, self.feature_path = video_name, root_path, feature_path
        self.label_name = label_name
        self.num_phase, self.phase_size = num_phase, phase_size
        self.frames = sorted(glob.glob(os.path.join(feature_path, '*.npy')),
                             key=lambda x: os.path.basename(x).split(".")[0])
        if len(self.frames) == 0:
            raise ValueError("number of frames of video {} should not be zero.".format(video_name))
        if len(self.frames) < length:
            print("The length exceeds the number of exsisting frames, the sampling length has been changed to {}".format(len(self.frames)))
        self.seq_ranges = snippet_ranges(len(self.frames), length, stride)
        self.length = min(length, len(self.frames))
        self.n_labels = 1 if label_name is None else len(label_name.split("_"))

    def __len__(self):
        return len(self.seq_ranges)

    def _gray(self, frame_path):
        from PIL import Image
        f_index = int(os.path.basename(frame_path).split(".")[0])
        path = os.path.join(self.root_path, self.video_name + "_aligned", 'frame_det_00_{:06d}.bmp'.format(f_index))
        try:
            img = Image.open(path).convert('L')
        except Exception:
            raise ValueError("incorrect face path")
        img = img.resize((self.phase_size, self.phase_size), Image.LANCZOS)     # GroupScale, data_utils.py:71-84
        return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).float().div(255)

    def __getitem__(self, index):
        start, end = self.seq_ranges[index]
        feats = np.array([np.load(f) for f in self.frames[start:end]])
        ids = window_index(start, end, len(self.frames), self.num_phase)
        cache = {int(i): self._gray(self.frames[int(i)]) for i in ids.unique()}
        phase = torch.stack([torch.stack([cache[int(i)] for i in row]) for row in ids])
        labels = np.array([[-100] * self.n_labels] * (end - start))
        return phase, feats, labels, np.array([start, end]), self.video_name
