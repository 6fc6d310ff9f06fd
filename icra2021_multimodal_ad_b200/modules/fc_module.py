"""modules/fc_module.py of the reference: a stack of FCLayers; hidden layers carry activation +
BatchNorm, the last one is a bare Linear (42-56).  ``layer_list`` stays a python list because the
RaPP scorer iterates it (reconstruction_aggregation.py:25)."""
import torch.nn as nn

from ..decorators import variational_info_bottleneck as vib
from ..layers import FCLayer


class FCModule(nn.Module):
    def __init__(self, input_size, output_size, hidden_sizes=None, use_batch_norm=True, dropout_p=0,
                 act="leakyrelu", last_act=None):
        super().__init__()
        hidden_sizes = list(hidden_sizes or [])
        if use_batch_norm and dropout_p > 0:
            raise Exception("Either batch_norm or dropout is allowed, not both")
        self.layer_list = []
        sizes = [input_size] + hidden_sizes + [output_size]
        for idx, (k, n) in enumerate(zip(sizes[:-1], sizes[1:])):
            if idx < len(hidden_sizes):
                self.layer_list.append(FCLayer(k, n, act=act, bn=use_batch_norm, dropout_p=dropout_p))
            else:
                self.layer_list.append(FCLayer(k, n, act=last_act))
        self.net = nn.Sequential(*self.layer_list)
        self.widths = sizes

    @vib
    def forward(self, x):
        for layer in self.layer_list:
            x = layer(x)
        return x
