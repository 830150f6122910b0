"""Drop-in for api/mimamo_net.py: same module tree / state_dict keys, CUDA-native forward.

`Two_Stream_RNN` stays an `nn.Module` so `load_state_dict(checkpoint['state_dict'])`
(api/tester.py:47-48), `.eval()` and `.to(device)` behave as before; its eval-mode forward is one
call into libmimamo_b200.so (mimamo_head_forward).  `MLP` and `PhaseNet` also run on their own
(mimamo_mlp_forward / mimamo_phasenet_forward).  Training is out of scope.
"""
import torch
import torch.nn as nn

import _nets


def _dense(i, o, dropout):           # [Dropout, Linear, BN, ReLU] -> keys 1, 2 (reference MLP :13-19)
    return [nn.Dropout(dropout), nn.Linear(i, o), nn.BatchNorm1d(o), nn.ReLU(inplace=True)]


class MLP(nn.Module):
    def __init__(self, hidden_units, dropout=0.3):
        super(MLP, self).__init__()
        assert len(hidden_units) > 1 and hidden_units[-1] == 256
        layers = []
        for i, o in zip(hidden_units[:-1], hidden_units[1:]):
            layers += _dense(i, o, dropout)
        self.mlp = nn.Sequential(*layers)

        self._native = None

    def load_state_dict(self, *args, **kwargs):
        self._native = None
        return super(MLP, self).load_state_dict(*args, **kwargs)

    def forward(self, input_tensor):
        """(bs, num_frames, features) -> (bs, num_frames, 256), eval mode (reference :22-26), via mimamo_mlp_forward.
        Inside Two_Stream_RNN.forward the same layers run as part of mimamo_head_forward."""
        if self.training:
            raise RuntimeError('MLP (B200) is inference only: call .eval() first')
        if self._native is None:
            self._native = _nets.NativeMLP(self.state_dict())
        bs, num_frames, feature_dim = input_tensor.size()
        with torch.no_grad():
            return self._native.forward(input_tensor.reshape(bs * num_frames, feature_dim)).view(bs, num_frames, -1)


class PhaseNet(nn.Module):
    def __init__(self, input_size, num_channels, hidden_units=[256, 256, 1], dropout=0.3, feature=False):
        super(PhaseNet, self).__init__()
        if input_size not in [48, 96, 112]:
            raise ValueError("Incorrect input size")
        if list(hidden_units) != [256, 256, 1]:
            raise NotImplementedError('the CUDA PhaseNet is built for hidden_units=[256, 256, 1]')
        self.num_channels = num_channels
        self.input_size = input_size
        n_blocks = 3 if input_size == 48 else 4                    # reference :33-40
        widths = [64 << b for b in range(n_blocks)]
        ins = [num_channels, num_channels + 64] + widths[1:-1]
        self.conv_net = nn.ModuleList([self._block(i, o) for i, o in zip(ins, widths)])
        self.dropout = nn.Dropout2d(p=0.2)
        last_conv_width = 6 if input_size in (48, 96) else 7
        self.avgpool = nn.AvgPool2d(kernel_size=[last_conv_width, last_conv_width])
        fc, prev = [], widths[-1]
        for h in hidden_units[:-1]:                                   # keys 0,2 / 4,6
            fc += [nn.Linear(prev, h), nn.ReLU(inplace=True), nn.BatchNorm1d(h), nn.Dropout(dropout)]
            prev = h
        self.fc = nn.Sequential(*fc)
        self.classifier = nn.Sequential(nn.Linear(hidden_units[-2], hidden_units[-1]),
                                        nn.BatchNorm1d(1, eps=1e-6, momentum=0.1))     # unused when feature=True
        self.feature = feature

    @staticmethod
    def _block(i, o):                                                # keys 0,1,3,4
        return nn.Sequential(nn.Conv2d(i, o, 3, padding=1), nn.BatchNorm2d(o), nn.ReLU(inplace=True),
                             nn.Conv2d(o, o, 3, padding=1, stride=2), nn.BatchNorm2d(o), nn.ReLU(inplace=True))

    def load_state_dict(self, *args, **kwargs):
        self._native = None
        return super(PhaseNet, self).load_state_dict(*args, **kwargs)

    def forward(self, data_level0, data_level1):
        """(bs, frames, C, S, S), (bs, frames, C, S/2, S/2) -> (bs*frames, 256) when `feature` else (bs*frames, 1), eval
        mode (reference :79-95), via mimamo_phasenet_forward."""
        if self.training:
            raise RuntimeError('PhaseNet (B200) is inference only: call .eval() first')
        if getattr(self, '_native', None) is None:
            self._native = _nets.NativePhaseNet(self.state_dict(), self.num_channels, self.input_size)
        bs, num_frames, num_channel, W0, H0 = data_level0.size()
        bs, num_frames, num_channel, W1, H1 = data_level1.size()
        with torch.no_grad():
            return self._native.forward(data_level0.reshape(bs * num_frames, num_channel, W0, H0),
                                        data_level1.reshape(bs * num_frames, num_channel, W1, H1), self.feature)


class Two_Stream_RNN(nn.Module):
    def __init__(self, mlp_hidden_units=[2048, 256, 256], dropout=0.5, label_name='arousal_valence',
                 num_phase=12):
        super(Two_Stream_RNN, self).__init__()
        if mlp_hidden_units[0] != 2048 or len(label_name.split("_")) != 2:
            raise NotImplementedError('the CUDA head takes 2048 ResNet50 features and predicts two labels '
                                      '(label_name "arousal_valence"); any MLP depth / widths ending in 256 are fine')
        if not 1 <= num_phase <= 32:
            raise NotImplementedError('num_phase must be in [1, 32]')
        self.mlp = MLP(mlp_hidden_units)
        self.num_phase = num_phase
        self.phasenet = PhaseNet(48, 2 * num_phase, hidden_units=[256, 256, 1], dropout=0.3, feature=True)
        self.transform = nn.Sequential(nn.Linear(512, 256), nn.ReLU(inplace=True), nn.BatchNorm1d(256),
                                       nn.Dropout(dropout))
        # no batch_first: the recurrence runs over dim 0 of (bs, frames, 256) -- kept
        self.rnns = nn.GRU(256, 128, bidirectional=True, num_layers=2, dropout=0.3)
        self.classifier = nn.Sequential(nn.Dropout(dropout), nn.Linear(256, 2), nn.BatchNorm1d(2))
        self._native = None

    def load_model_weights(self, model, model_path):
        ckp = torch.load(model_path)
        net_key = [key for key in ckp.keys() if (key != 'epoch') and (key != 'iter')][0]
        model.load_state_dict(ckp[net_key])
        return model

    def load_state_dict(self, *args, **kwargs):
        self._native = None                     # folded device weights are rebuilt lazily
        return super(Two_Stream_RNN, self).load_state_dict(*args, **kwargs)

    def refresh(self):
        """Re-fold the current parameters into the device-side head (call after editing weights)."""
        self._native = _nets.NativeHead(self.state_dict(), self.num_phase)
        return self

    def forward(self, phase_data, rgb_data):
        if self.training:
            raise RuntimeError('Two_Stream_RNN (B200) is inference only: call .eval() first')
        if self._native is None:
            self.refresh()
        phase_0, phase_1 = phase_data
        with torch.no_grad():
            return self._native.forward(phase_0, phase_1, rgb_data)

    def forward_operands(self, phase0_nhwc, cat_nhwc, rgb_data):
        """forward() fed by Phase_Difference_Extractor.phasenet_operands (fp16 NHWC phase differences straight from the
        phase tail); bit-identical to forward() on the fp32 phase tensors."""
        if self.training:
            raise RuntimeError('Two_Stream_RNN (B200) is inference only: call .eval() first')
        if self._native is None:
            self.refresh()
        with torch.no_grad():
            return self._native.forward_nhwc16(phase0_nhwc, cat_nhwc, rgb_data)
