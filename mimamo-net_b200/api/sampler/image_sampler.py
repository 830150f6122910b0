"""Test-mode image sampler over an OpenFace output directory (api/sampler/image_sampler.py:63-141):
yields (transformed RGB tensor, dummy label, frame path, video name).  Host I/O only."""
import glob
import os

import numpy as np
import torch.utils.data as data


class Image_Sampler(data.Dataset):
    def __init__(self, video_name, root_path, test_mode=False, annot_dir=None, label_name=None,
                 transform=None, verbose=False, size=224):
        if not test_mode:
            raise NotImplementedError('training-mode sampling (labels, augmentation) is out of scope')
        assert transform is not None
        self.video_name, self.transform = video_name, transform
        frames = glob.glob(os.path.join(root_path, video_name + "_aligned", '*.bmp'))
        self.frames = sorted(frames, key=lambda x: os.path.basename(x).split(".")[0].split("_")[-1])
        if len(self.frames) == 0:
            raise ValueError("number of frames of video {} should not be zero.".format(video_name))

    def __len__(self):
        return len(self.frames)

    def __getitem__(self, index):
        from PIL import Image
        frame = self.frames[index]
        return self.transform(Image.open(frame)), np.array([-100]), frame, self.video_name
