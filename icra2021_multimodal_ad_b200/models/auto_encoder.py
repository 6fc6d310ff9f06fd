"""models/auto_encoder.py of the reference, backed by the fused sm_100a engine.

Same surface: encode / decode / forward / get_loss_value / step / validate / attach,
``model.encoder.layer_list``, reference ``state_dict`` keys.  Eval-mode calls run the fused
layer chain of libmmad; train-mode ``get_loss_value`` returns a loss whose ``backward()``
runs the fused backward kernels (``train.py``), so the reference's ``step`` body works verbatim.
"""
import torch

from .. import _lib
from ..engine import Engine
from .abstract_model import AbstractModel


class AutoEncoder(AbstractModel):
    def __init__(self, encoder, decoder, recon_loss, precision="fp32"):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.recon_loss = recon_loss
        self.precision = precision
        self._eng = None
        self._eng_key = None

    # ---- device engine ---------------------------------------------------------------
    def _state_key(self):
        ts = list(self.parameters()) + list(self.buffers())
        return (tuple(t.data_ptr() for t in ts), tuple(t._version for t in ts), self.precision,
                getattr(self, "_train_steps", 0))

    def handle_engine(self) -> Engine:
        """The device engine (library handle) without (re)packing the eval-mode weights."""
        p = self.encoder.layer_list[0].layer.weight
        if not p.is_cuda:
            raise _lib.MmadError("model is on the CPU; the B200 path has no CPU fallback (use gpu_id >= 0)")
        if self._eng is None or self._eng.device != p.device:
            self._eng = Engine(self.encoder.widths, self.decoder.widths, precision=self.precision, device=p.device)
            self._eng_key = None
        return self._eng

    def engine(self) -> Engine:
        """The packed device engine, re-packed whenever a parameter or BatchNorm buffer changed."""
        self.handle_engine()
        key = self._state_key()
        if key != self._eng_key:
            if self._eng.precision != self.precision:
                self._eng.set_precision(self.precision)
            self._eng.load_state_dict(self.state_dict())
            self._eng_key = key
        return self._eng

    def set_precision(self, precision):
        """'fp32' (CUDA-core fp32, default), 'f16x3' (tcgen05, fp16 hi/lo split), 'f16f8' (tcgen05, fp16 hi*hi + fp8
        cross terms: the bulk-scoring mode), 'f16' (one pass).  Error bars: DESIGN.md section 3."""
        self.precision = precision
        return self

    # ---- reference API ---------------------------------------------------------------
    def encode(self, x):
        # |x| = (batch_size, ...)  models/auto_encoder.py:36-39
        if self.training and torch.is_grad_enabled():
            raise RuntimeError("use get_loss_value()/step() for training (fused forward+backward)")
        x2 = x.reshape(x.size(0), -1)
        _, z = self.engine().forward(x2, want_code=True)
        return z.view(x.size(0), -1)

    def decode(self, z):
        # models/auto_encoder.py:41-44 -- decoder alone, layer by layer
        return self.decoder(z)

    def forward(self, x):
        # models/auto_encoder.py:46-50
        if self.training:
            raise RuntimeError("train-mode forward goes through get_loss_value()/step() (fused kernels); "
                               "call model.eval() for inference")
        x2 = x.reshape(x.size(0), -1)
        return self.engine().forward(x2).view(x.size(0), -1)

    def get_loss_value(self, x, y, *args, **kwargs):
        # models/auto_encoder.py:52-55: sum-reduced MSE between model(x) and x
        x2 = x.reshape(x.size(0), -1)
        if self.training and torch.is_grad_enabled():
            from ..train import fused_train_loss
            return fused_train_loss(self, x2)
        if self.training:
            raise RuntimeError("train-mode loss without grad is not on the reference path")
        return self.engine().recon_loss(x2)

    @staticmethod
    def step(engine, mini_batch):
        # models/auto_encoder.py:57-77.  The step ends with a device->host read of the loss, so everything the host does
        # before the next launch is GPU idle time: Module.train() walks ~40 submodules (25 us) -- skipped when the model
        # is already in train mode.
        if not engine.model.training:
            engine.model.train()
        engine.optimizer.zero_grad()
        x, _ = mini_batch
        if engine.config.gpu_id >= 0:
            x = x.cuda(engine.config.gpu_id)
        x = x.view(x.size(0), -1)
        loss = engine.model.get_loss_value(x, x)
        loss.backward(retain_graph=True)
        engine.optimizer.step()
        from ..train import step_loss
        return (step_loss(engine.model, loss), )     # float(loss) without waiting for backward + Adam (train.step_loss)

    @staticmethod
    def validate(engine, mini_batch):
        # models/auto_encoder.py:79-91
        engine.model.eval()
        with torch.no_grad():
            x, _ = mini_batch
            if engine.config.gpu_id >= 0:
                x = x.cuda(engine.config.gpu_id)
            x = x.view(x.size(0), -1)
            loss = engine.model.get_loss_value(x, x)
        return (float(loss), )

    @staticmethod
    def attach(trainer, evaluator, config):
        """models/auto_encoder.py:93-123 (ignite logging).  ignite is optional: without it the
        trainer from ``novelty_detection.py`` of this package keeps its own running average."""
        try:
            from ignite.engine import Events
            from ignite.metrics import RunningAverage
        except ImportError:
            return
        RunningAverage(output_transform=lambda x: x[0]).attach(trainer, "recon")
        RunningAverage(output_transform=lambda x: x[0]).attach(evaluator, "recon")
        if config.verbose >= 1:
            @trainer.on(Events.EPOCH_COMPLETED)
            def print_train_logs(engine):
                print("Epoch {} - loss={:.4e}".format(engine.state.epoch, engine.state.metrics["recon"]))

            @evaluator.on(Events.EPOCH_COMPLETED)
            def print_valid_logs(engine):
                print("Validation - recon={:.4e} lowest_recon={:.4e}".format(engine.state.metrics["recon"],
                                                                             engine.lowest_loss))
