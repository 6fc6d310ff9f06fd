"""The multimodal feature extractor of utils/data_loaders.py: ``HSR_Net``
(152-229), ``Multisensory_module`` (601-674), ``norm_vec`` (703-712) and ``HsrDataset`` (714-731), and
``get_input_size`` (16-29).

Same classes, constructor arguments, parameter names (``conv1r.weight`` ... -- the unused LiDAR / ``conv*m``
layers are kept so ``state_dict`` round-trips) and outputs ``[B, 27, 8, 8]`` (``.view(B, -1)`` -> 1728), but the
per-sample Python loop of tiny convolutions is ONE fused launch of libmmad (``mmad_multisensory_forward``):
one CTA per sample, all intermediates in shared memory, ``norm_vec`` folded into the loads.
File / ROS / librosa loading stays out of scope (unpublished data).
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _lib


def get_input_size(config):
    # utils/data_loaders.py:16-29
    sensor = getattr(config, "sensor", "All")
    sizes = {"All": 1728, "hand_camera": 1024, "head_depth": 512, "LiDAR": 2048, "mic": 128, "force_torque": 64}
    if sensor not in sizes:
        raise ValueError("unknown sensor {}".format(sensor))
    return sizes[sensor]


def norm_vec(v, range_in=None, range_out=None):
    # utils/data_loaders.py:703-712
    a, b = _norm_affine(v, range_in, range_out)
    return v * a + b


def _norm_affine(v, range_in=None, range_out=None):
    """(scale, shift) with norm_vec(v) == v * scale + shift."""
    if range_out is None:
        range_out = [-1, 1]
    if range_in is None:
        range_in = [float(torch.min(v)), float(torch.max(v))]
    r_out = range_out[1] - range_out[0]
    r_in = range_in[1] - range_in[0]
    scale = r_out / r_in
    return scale, range_out[0] - range_in[0] * scale


class _FeatureNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1r = nn.Conv2d(3, 16, kernel_size=2, stride=2)
        self.conv2r = nn.Conv2d(16, 16, kernel_size=3, stride=1, padding=1)
        self.conv3r = nn.Conv2d(16, 16, kernel_size=2, stride=2)
        self.conv1d = nn.Conv2d(1, 8, kernel_size=2, stride=2)
        self.conv2d = nn.Conv2d(8, 8, kernel_size=3, stride=1, padding=1)
        self.conv3d = nn.Conv2d(8, 8, kernel_size=2, stride=2)
        self.conv1l = nn.Conv1d(1, 8, kernel_size=18, stride=9, padding=9)
        self.conv2l = nn.Conv1d(8, 16, kernel_size=2, stride=2)
        self.conv3l = nn.Conv1d(16, 32, kernel_size=2, stride=2)
        self.conv4l = nn.Conv1d(32, 16, kernel_size=3, stride=2, padding=3)
        self.conv5l = nn.Conv1d(16, 32, kernel_size=2, stride=2)
        self.conv1m = nn.Conv1d(1, 12, kernel_size=2, stride=1)
        self.conv2m = nn.Conv1d(12, 8, kernel_size=2, stride=2, padding=2)

    def _weights(self):
        w = _lib.FeatureWeights()
        keep = []
        for name in ("conv1r", "conv2r", "conv3r", "conv1d", "conv2d", "conv3d", "conv1l", "conv2l"):
            conv = getattr(self, name)
            for suffix, t in (("_w", conv.weight), ("_b", conv.bias)):
                t = t.detach().float().contiguous()
                keep.append(t)
                setattr(w, name + suffix, t.data_ptr())
        return w, keep

    def extract(self, r, d, t, m, affine=None):
        """Fused forward for the whole batch: any of r [B,(1,)3,32,32], d [B,(1,)1,32,32], t [B], m [B,(1,1,)13]
        may be None.  Returns [B, C, 8, 8] with C = 16 (r) + 8 (d) + 1 (t) + 2 (m) of the given modalities."""
        given = [x for x in (r, d, t, m) if x is not None]
        if not given:
            raise ValueError("no modality given")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.MmadError("the feature extractor needs CUDA tensors (no CPU path)")
        B = given[0].shape[0]

        def prep(x, per):
            if x is None:
                return None
            x = x.detach().to(dev, torch.float32).reshape(B, -1).contiguous()
            if x.shape[1] != per:
                raise ValueError("expected {} values per sample, got {}".format(per, x.shape[1]))
            return x
        r2, d2, t2, m2 = prep(r, 3 * 32 * 32), prep(d, 32 * 32), prep(t, 1), prep(m, 13)
        width = _lib.lib().mmad_multisensory_width(r2 is not None, d2 is not None, t2 is not None, m2 is not None)
        out = torch.empty(B, width, dtype=torch.float32, device=dev)
        w, keep = self._weights()
        aff = (C.c_float * 8)(*(affine if affine is not None else (1, 0, 1, 0, 1, 0, 1, 0)))
        p = lambda x: None if x is None else x.data_ptr()  # noqa: E731
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mmad_multisensory_forward(p(r2), p(d2), p(t2), p(m2), B, C.byref(w), aff, out.data_ptr(),
                                                            width, torch.cuda.current_stream().cuda_stream))
        del keep
        return out.view(B, width // 64, 8, 8)


class Multisensory_module(_FeatureNet):
    # utils/data_loaders.py:601-674
    def __init__(self, config, unimodal=False):
        super().__init__()
        self.batch_size = config.batch_size
        self.config = config
        self.unimodal = unimodal

    def forward(self, r, d, t, m):
        return self._forward(r, d, t, m)

    def _forward(self, r, d, t, m, affine=None):
        n = self.batch_size
        cut = lambda x: None if x is None else x[:n]  # noqa: E731   (the reference loops over range(batch_size))
        r, d, t, m = cut(r), cut(d), cut(t), cut(m)
        if self.unimodal:   # the reference keeps only the LAST given modality (order r, d, t, m)
            last = [i for i, x in enumerate((r, d, t, m)) if x is not None][-1]
            r, d, t, m = [x if i == last else None for i, x in enumerate((r, d, t, m))]
        elif any(x is None for x in (r, d, t, m)):
            raise ValueError("multimodal fusion needs r, d, t and m (utils/data_loaders.py:667-668)")
        return self.extract(r, d, t, m, affine)


class HSR_Net(Multisensory_module):
    # utils/data_loaders.py:152-229
    def __init__(self, unimodal, config):
        nn.Module.__init__(self)
        _FeatureNet.__init__(self)
        self.batch_size = config.slicing_size
        self.config = config
        self.unimodal = unimodal

    def forward(self, r, d, l, t, m):  # noqa: E741
        if l is not None:
            raise NotImplementedError("the LiDAR branch is not wired in the reference's own callers "
                                      "(utils/data_loaders.py:401: hsr_net(r, d, None, t, m))")
        return self._forward(r, d, t, m)


def HsrDataset(config, force_q, hand_q, depth_q, mic_q):
    # utils/data_loaders.py:714-731 -- normalisation folded into the fused kernel's loads
    t = torch.as_tensor(force_q, dtype=torch.float32)
    r = torch.as_tensor(hand_q, dtype=torch.float32).view(-1, 1, 3, 32, 32)
    d = torch.as_tensor(depth_q, dtype=torch.float32).view(-1, 1, 1, 32, 32)
    m = torch.as_tensor(mic_q, dtype=torch.float32).view(-1, 1, 1, 13)
    affine = (*_norm_affine(r, [0, 255]), *_norm_affine(d, [0, 255]), *_norm_affine(t, [0, 400]), *_norm_affine(m))
    module = Multisensory_module(config).cuda(config.gpu_id)
    fusion = module._forward(r, d, t, m, affine)
    return fusion.view(config.batch_size, -1)
