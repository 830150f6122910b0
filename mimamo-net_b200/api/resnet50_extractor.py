"""Drop-in for api/resnet50_extractor.py: ResNet50 `pool5_7x7_s1` features on the tcgen05 engine."""
import os

import numpy as np
import torch

import _nets
from steerable.utils import get_device

device = get_device()          # import-time default, like the reference; methods re-query the current device

MEAN = (131.0912, 103.8827, 91.4953)       # resnet50_ferplus_dag meta: 0-255 scale, std 1


def _load_third_party(model_name, model_dir):
    """api/utils/model_utils.py:65-79: import <model_dir>/<name>.py and call <name>(weights_path=...)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(model_name, os.path.join(model_dir, model_name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return getattr(mod, model_name)(weights_path=os.path.join(model_dir, model_name + '.pth'))


class Resnet50_Extractor(object):
    def __init__(self, benchmark_dir='pytorch-benchmarks', model_name='resnet50_ferplus_dag',
                 feature_layer='pool5_7x7_s1', model=None):
        '''Same arguments as the reference (:14-41).  `model` (an nn.Module or a state_dict with the
        resnet50_ferplus_dag key names) bypasses the on-disk third-party checkpoint, e.g. for
        synthetic weights.'''
        self.benchmark_dir = os.path.abspath(benchmark_dir)
        self.model_name = model_name
        self.feature_layer = feature_layer
        if feature_layer != 'pool5_7x7_s1':
            raise NotImplementedError('only the pool5_7x7_s1 tap is implemented')
        if model is None:
            assert os.path.exists(self.benchmark_dir), 'benchmark_dir must exits'
            model = _load_third_party(self.model_name, os.path.abspath(os.path.join(self.benchmark_dir, 'ferplus')))
        self.meta = getattr(model, 'meta', {'mean': list(MEAN), 'std': [1, 1, 1], 'imageSize': [224, 224, 3]})
        state = model if isinstance(model, dict) else model.state_dict()
        self.model = _nets.NativeResNet50(state)
        self.transform = self._compose_transforms(self.meta)

    @staticmethod
    def _compose_transforms(meta, resize=256):
        # api/utils/model_utils.py:6-40 with center_crop=True: Resize(256) -> CenterCrop -> ToTensor
        # -> x255 (std == [1,1,1]) -> Normalize(mean, std)
        import torchvision.transforms as transforms
        size = meta['imageSize']
        steps = [transforms.Resize(resize), transforms.CenterCrop(size=(size[0], size[1])), transforms.ToTensor()]
        if meta['std'] == [1, 1, 1]:
            steps.append(lambda x: x * 255.0)
        steps.append(transforms.Normalize(mean=meta['mean'], std=meta['std']))
        return transforms.Compose(steps)

    def features(self, image):
        """Device-resident variant of get_vec: float32 CUDA in, (bs,2048) float32 CUDA out."""
        return self.model.pool5(image)

    def features_host(self, image_host, chunk=128, to_host=True):
        """get_vec for a HOST batch (pinned for a truly asynchronous copy): (bs,3,224,224) float32 -> (bs,2048).  The
        images are streamed to the device in chunks on a copy stream, double buffered, while ResNet50 runs on the
        previous chunk; the features come back in one copy (or stay on the device with to_host=False)."""
        device = get_device()
        main = torch.cuda.current_stream(device)
        if getattr(self, '_copy_stream', None) is None or self._bufs[0].shape[0] != chunk:
            self._copy_stream = torch.cuda.Stream(device)
            self._bufs = [torch.empty((chunk, 3, 224, 224), dtype=torch.float32, device=device) for _ in range(2)]
        copy, bufs = self._copy_stream, self._bufs
        n = image_host.shape[0]
        spans = [(s, min(n, s + chunk)) for s in range(0, n, chunk)]
        ready = [torch.cuda.Event() for _ in spans]
        free = [torch.cuda.Event() for _ in spans]
        copy.wait_stream(main)

        def issue(k):
            s, e = spans[k]
            with torch.cuda.stream(copy):
                if k >= 2:
                    copy.wait_event(free[k - 2])
                bufs[k % 2][:e - s].copy_(image_host[s:e], non_blocking=True)
                ready[k].record(copy)

        feats = torch.empty((n, 2048), dtype=torch.float32, device=device)
        for k in range(min(2, len(spans))):
            issue(k)
        for k, (s, e) in enumerate(spans):
            main.wait_event(ready[k])
            feats[s:e] = self.features(bufs[k % 2][:e - s])
            free[k].record(main)
            if k + 2 < len(spans):
                issue(k + 2)
        return feats.cpu() if to_host else feats

    def features_from_crops(self, crops, preprocessor):
        """(bs,S,S,3) uint8 CUDA face crops -> (bs,2048) float32 CUDA: `self.transform` (Resize 256 ->
        CenterCrop 224 -> ToTensor -> x255 -> Normalize) runs on the device, bit-exact with PIL."""
        return self.model.pool5_crops(preprocessor, crops)

    def get_vec(self, image):
        """(bs,3,224,224) -> relu(pool5) as a CPU tensor, like the reference's hook-to-CPU (:74-83),
        including its .squeeze() (bs == 1 collapses the batch dim)."""
        return self.features(image.to(get_device())).cpu().squeeze()

    def run(self, input_dir, output_dir, batch_size=64, video_name=''):
        '''Write one %05d.npy (float32[2048]) per aligned face of <input_dir>/<video>_aligned (:42-73).'''
        assert os.path.exists(input_dir), 'input dir must exsit!'
        assert len(os.listdir(input_dir)) != 0, 'input dir must not be empty!'
        assert len(video_name) != 0, 'input video name cannot be empty!'
        from sampler.image_sampler import Image_Sampler
        dataset = Image_Sampler(video_name, input_dir, test_mode=True, transform=self.transform)
        loader = torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=False, drop_last=False,
                                             num_workers=0, pin_memory=True)
        if not os.path.exists(output_dir):
            os.makedirs(output_dir)
        elif len(os.listdir(output_dir)) != 0 and '.npy' in os.listdir(output_dir)[0]:
            print("output_dir {} already exists, feature extraction skipped.".format(output_dir))
            return
        device = get_device()
        with torch.no_grad():
            for ims, target, img_path, names in loader:
                feats = self.features(ims.to(device, non_blocking=True)).cpu().numpy()
                for feature, path in zip(feats, img_path):
                    np.save(os.path.join(output_dir, "%05d.npy" % self.get_frame_index(path)), feature)

    def get_frame_index(self, frame_path):
        frame_name = frame_path.split('/')[-1]
        return int(frame_name.split('.')[0].split('_')[-1])
