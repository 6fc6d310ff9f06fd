"""CPU restatement of the reference's multimodal feature extractor.  TEST INFRASTRUCTURE.

utils/data_loaders.py:601-674 (Multisensory_module.forward; HSR_Net.forward 152-229 is the same arithmetic):
per sample, ReLU conv stacks on the RGB and depth images, broadcast force-torque scalar, two 1-d convolutions on
the MFCC vector, channel concatenation.  Written with torch functional ops on the whole batch (the reference loops
over samples; convolutions do not mix samples).  Pinned against tests/golden/features.pt, produced by the
unmodified reference class (tests/golden/make_golden.py)."""
import torch
import torch.nn.functional as F


def multisensory_forward(sd, r=None, d=None, t=None, m=None):
    """sd: state dict with the reference's keys (conv1r.weight ...).  Inputs as the reference takes them:
    r [B,1,3,32,32], d [B,1,1,32,32], t [B], m [B,1,1,13]; None drops the modality.  Returns [B, C, 8, 8]."""
    parts = []
    if r is not None:
        x = r.reshape(-1, 3, 32, 32)
        x = F.relu(F.conv2d(x, sd["conv1r.weight"], sd["conv1r.bias"], stride=2))
        x = F.relu(F.conv2d(x, sd["conv2r.weight"], sd["conv2r.bias"], stride=1, padding=1))
        parts.append(F.relu(F.conv2d(x, sd["conv3r.weight"], sd["conv3r.bias"], stride=2)))
    if d is not None:
        x = d.reshape(-1, 1, 32, 32)
        x = F.relu(F.conv2d(x, sd["conv1d.weight"], sd["conv1d.bias"], stride=2))
        x = F.relu(F.conv2d(x, sd["conv2d.weight"], sd["conv2d.bias"], stride=1, padding=1))
        parts.append(F.relu(F.conv2d(x, sd["conv3d.weight"], sd["conv3d.bias"], stride=2)))
    if t is not None:
        parts.append(t.reshape(-1, 1, 1, 1).repeat(1, 1, 8, 8))           # utils/data_loaders.py:645-648
    if m is not None:
        x = m.reshape(-1, 1, 13)
        x = F.relu(F.conv1d(x, sd["conv1l.weight"], sd["conv1l.bias"], stride=9, padding=9))     # 655
        x = F.relu(F.conv1d(x, sd["conv2l.weight"], sd["conv2l.bias"], stride=2))                # 656
        parts.append(x.reshape(-1, 2, 8, 1).repeat(1, 1, 1, 8))                                  # 657
    return torch.cat(parts, dim=1)


def norm_vec(v, range_in=None, range_out=None):
    """utils/data_loaders.py:703-712."""
    if range_out is None:
        range_out = [-1, 1]
    if range_in is None:
        range_in = [torch.min(v), torch.max(v)]
    return ((range_out[1] - range_out[0]) * (v - range_in[0]) / (range_in[1] - range_in[0])) + range_out[0]
