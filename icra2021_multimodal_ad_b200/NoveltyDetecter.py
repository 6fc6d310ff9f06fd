"""The realtime caller's detector: ``from NoveltyDetecter import NoveltyDetecter`` at test_file/realtime_tester.py:262 names a
module that is NOT in the reference repository; its one call site (291-304) fixes the contract::

    detecter = NoveltyDetecter(config)
    score = detecter.test(model, fusion_representation, config, nap=False)     # one value per window of the batch
    new_val = np.array(score)

``fusion_representation`` is the ``(batch_size = 10, 1728)`` matrix of the latest windows.  Here ``test`` scores it with the
SAP score (``nap=False``, utils/metric.py:145-171) or the NAP score (``nap=True``; the fit comes from the checkpoint
``config.nap_fit`` written by ``novelty_detection.NoveltyDetecter.score_fast``).  Host matrices of up to 64 windows run in ONE
kernel launch (``mmad_score_host`` -> stream kernel); ``window_buffer`` exposes the pinned, device-mapped input buffer of that
kernel, so a caller that writes its windows there (the ROS callbacks' deques, realtime_tester.py:34-203) pays no copy at all.
"""
import numpy as np
import torch

from .engine import clamp_layer_range


class NoveltyDetecter():
    def __init__(self, config):
        self.config = config

    def _range(self, eng, cfg):
        end = cfg.n_layers + 1 - getattr(cfg, "end_layer_index", -1)      # novelty_detection.py:56-57
        return clamp_layer_range(eng.n_diffs, getattr(cfg, "start_layer_index", 0), end)

    def window_buffer(self, model, config=None):
        """Pinned ``[64, D]`` float32 array owned by the engine: rows written here are scored in place by ``test``."""
        eng = model.eval().engine()
        lo, hi = self._range(eng, config or self.config)
        return eng.stream_input(lo, hi)

    def test(self, model, x, config=None, nap=False):
        cfg = config or self.config
        eng = model.eval().engine()
        lo, hi = self._range(eng, cfg)
        if nap and eng.nap_range != (lo, hi):
            path = getattr(cfg, "nap_fit", None)
            if not path:
                raise ValueError("nap=True needs a NAP fit: set config.nap_fit to a checkpoint written by score_fast")
            eng.load_nap_state_dict(torch.load(path, weights_only=False))
        if isinstance(x, torch.Tensor) and x.is_cuda:                 # get_realtime_dataloader builds the batch on the GPU
            out = eng.score(x.float().reshape(x.shape[0], -1), lo, hi, base=False, sap=not nap, nap=nap)
            return out["nap" if nap else "sap"].cpu().numpy().tolist()
        if isinstance(x, torch.Tensor):
            x = x.detach().numpy()
        x = np.asarray(x, dtype=np.float32)
        x = x.reshape(x.shape[0], -1)
        out = eng.score_host(x, lo, hi, base=False, sap=not nap, nap=nap)
        return out["nap" if nap else "sap"].tolist()
