"""VIB-decorated autoencoder (BASELINE.json configs[3]).

The reference ships the decorator (decorators/variational_info_bottleneck.py:19-42) but no model class
that uses it and no KL term, so this wrapper is defined here: the encoder emits
``2 * btl`` values per sample, split (mu | logvar) like the decorator does; the decoder consumes
``z = eps * exp(logvar / 2) + mu`` (k = 1); loss = sum (x_hat - x)^2 + beta_kl * KL with
KL = -1/2 sum (1 + logvar - mu^2 - exp(logvar)).  The reparameterisation is pinned by the reference
(tests/golden/vib_D64.pt); the KL term is PARITY-UNPINNED (checked against a torch-autograd restatement).
The noise ``eps`` is an explicit input so CPU and GPU runs see the same sample.
"""
import torch

from ..train import fused_train_loss
from .auto_encoder import AutoEncoder


class VIBAutoEncoder(AutoEncoder):
    def __init__(self, encoder, decoder, recon_loss, beta_kl=1.0, precision="fp32"):
        super().__init__(encoder, decoder, recon_loss, precision=precision)
        if encoder.widths[-1] != 2 * decoder.widths[0]:
            raise ValueError("VIB encoder must emit 2 * btl_size values (mu | logvar)")
        self.beta_kl = beta_kl

    def encode(self, x, k=1, stochastic_inference=True, eps=None):
        # decorators/variational_info_bottleneck.py:29-42: dict of z (k,B,h), mu, logvar
        return self.encoder(x.reshape(x.size(0), -1), distribution="normal", k=k,
                            stochastic_inference=stochastic_inference, eps=eps)

    def forward(self, x, eps=None, stochastic_inference=True):
        if self.training:
            raise RuntimeError("train-mode forward goes through get_loss_value()/step() (fused kernels)")
        r = self.encode(x, k=1, stochastic_inference=stochastic_inference, eps=eps)
        return self.decoder(r["z"][0]).view(x.size(0), -1)

    def engine(self):
        raise NotImplementedError("the fused RaPP scorer takes plain autoencoders (the reference never scores a VIB model)")

    def get_loss_value(self, x, y, eps=None, *args, **kwargs):
        x2 = x.reshape(x.size(0), -1)
        h = self.decoder.widths[0]
        if self.training and torch.is_grad_enabled():
            if eps is None:
                eps = torch.randn(x2.size(0), h, device=x2.device, dtype=torch.float32)
            return fused_train_loss(self, x2, eps=eps, beta_kl=self.beta_kl)
        from ..ops import mse_sum
        return mse_sum(self.forward(x2, eps=eps, stochastic_inference=eps is not None), x2)
