"""modules/activation.py of the reference.  Only the variants the path instantiates are
device-backed: 'leakyrelu' (LeakyReLU(0.2), fused into the layer kernel) and None."""
import torch.nn as nn


class Activation(nn.Module):
    SLOPES = {"leakyrelu": 0.2}

    def __init__(self, act):
        super().__init__()
        if act is not None and act not in self.SLOPES:
            raise NotImplementedError(
                f"activation {act!r} is never instantiated by the reference path (model_builder.py:26,35); "
                "only 'leakyrelu' and None have sm_100a kernels")
        self.name = act
        self.slope = self.SLOPES.get(act)

    def forward(self, x):
        raise RuntimeError("Activation is fused into FCLayer's kernel; call the layer, not the activation")
