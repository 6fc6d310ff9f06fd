"""layers/fc_layer.py of the reference: Linear -> LeakyReLU(0.2) -> BatchNorm1d, executed by
one fused sm_100a kernel (``mmad_fc_layer_forward``).  ``nn.Linear`` / ``nn.BatchNorm1d`` are
kept as parameter containers so ``state_dict`` keys (``layer.weight``, ``bn.running_mean`` ...)
and initialisation are those of the reference (layers/fc_layer.py:23-35)."""
import torch
from torch import nn

from .. import _lib
from ..modules.activation import Activation


class FCLayer(nn.Module):
    def __init__(self, input_size, output_size=1, bias=True, act="relu", bn=False, dropout_p=0):
        super().__init__()
        if dropout_p:
            raise NotImplementedError("dropout is never enabled on the reference path (modules/fc_module.py:38-39)")
        if not bias:
            raise NotImplementedError("bias-free layers are not on the reference path")
        self.layer = nn.Linear(input_size, output_size, bias)
        self.bn = nn.BatchNorm1d(output_size) if bn else None
        self.dropout = None
        self.act = Activation(act) if act else None
        if self.bn is not None and self.act is None:
            raise NotImplementedError("BatchNorm without activation is not on the reference path")

    def forward(self, x):
        """Eval-mode forward (layers/fc_layer.py:37-48); >2-D inputs are flattened for the kernel
        exactly as the reference flattens them for BatchNorm (41-43)."""
        if self.training and torch.is_grad_enabled():
            raise RuntimeError("stand-alone FCLayer training is not supported; train through "
                               "AutoEncoder.step / get_loss_value (fused forward+backward kernels)")
        if self.bn is not None and self.training:
            raise RuntimeError("train-mode BatchNorm forward outside AutoEncoder.step is not supported")
        if not x.is_cuda:
            raise _lib.MmadError("FCLayer.forward needs CUDA tensors (no CPU path)")
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).float().contiguous()
        W = self.layer.weight
        n, K, N = x2.shape[0], W.shape[1], W.shape[0]
        y = torch.empty(n, N, dtype=torch.float32, device=x.device)
        p = lambda t: t.detach().contiguous().data_ptr()  # noqa: E731
        bn = self.bn
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mmad_fc_layer_forward(
                x2.data_ptr(), K, n, K, N, p(W), p(self.layer.bias),
                p(bn.weight) if bn else None, p(bn.bias) if bn else None,
                p(bn.running_mean) if bn else None, p(bn.running_var) if bn else None,
                self.act.slope if self.act else 0.0, bn.eps if bn else 1e-5,
                y.data_ptr(), N, torch.cuda.current_stream().cuda_stream))
        return y.reshape(*shp[:-1], N)
