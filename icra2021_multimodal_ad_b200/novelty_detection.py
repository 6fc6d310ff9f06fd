"""novelty_detection.py of the reference (NoveltyDetecter.train / .test) without ignite.

Same call surface: ``NoveltyDetecter(config).train(model, train_loader, valid_loader)`` returns
``(train_history, valid_history, test_history, model)`` with the best-on-validation weights loaded
(novelty_detection.py:88-127); ``.test(model, dset_manager, train_loader, valid_loader, test_loader, df_test)``
returns ``(base_auroc, base_aupr), (sap_auroc, sap_aupr), (nap_auroc, nap_aupr), df_test``
(novelty_detection.py:15-85).  ``.test_arrays`` takes the feature matrices directly and ``.score_fast`` is
the fused path (no diff materialisation, NAP fit from device statistics, metrics on device).
Loaders are any iterables of ``(x, y)`` mini-batches.
"""
from copy import deepcopy

import numpy as np
import torch

from .optim import Adam
from .reconstruction_aggregation import get_diffs, get_scores
from .utils import metric as M


class _RunningAverage:
    """ignite.metrics.RunningAverage(alpha=0.98), reset every epoch (models/auto_encoder.py:101-104)."""

    def __init__(self, alpha=0.98):
        self.alpha, self.value = alpha, None

    def update(self, x):
        self.value = x if self.value is None else self.value * self.alpha + (1.0 - self.alpha) * x


class _Engine:
    pass


class NoveltyDetecter():
    def __init__(self, config):
        self.config = config

    # ---- novelty_detection.py:88-127 ---------------------------------------------------------------
    def train(self, model, train_loader, valid_loader, optimizer=None):
        cfg = self.config
        optimizer = optimizer if optimizer is not None else Adam(model.parameters(), lr=1e-3)
        trainer, evaluator = _Engine(), _Engine()
        trainer.model, trainer.optimizer, trainer.config = model, optimizer, cfg
        evaluator.model, evaluator.config, evaluator.lowest_loss = model, cfg, np.inf
        train_history, valid_history, test_history = [], [], []
        best = None
        for epoch in range(getattr(cfg, "n_epochs", 20)):
            ra = _RunningAverage()
            for mini_batch in train_loader:
                ra.update(model.step(trainer, mini_batch)[0])
            train_history.append(float(ra.value))
            rv = _RunningAverage()
            for mini_batch in valid_loader:
                rv.update(model.validate(evaluator, mini_batch)[0])
            loss = float(rv.value)
            if loss < evaluator.lowest_loss:
                evaluator.lowest_loss = loss
                best = deepcopy(model.state_dict())
            valid_history.append(loss)
            if getattr(cfg, "verbose", 0) >= 1:
                print("Epoch {} - loss={:.4e}  validation recon={:.4e} lowest_recon={:.4e}".format(
                    epoch + 1, train_history[-1], loss, evaluator.lowest_loss))
        if best is not None:
            model.load_state_dict(best)
        return train_history, valid_history, test_history, model

    # ---- novelty_detection.py:15-85 ----------------------------------------------------------------
    def _labels(self, y):
        cfg = self.config
        hit = np.isin(np.asarray(y), [getattr(cfg, "target_class", 1)])
        return np.where(hit, False, True) if getattr(cfg, "unimodal_normal", False) else np.where(hit, True, False)

    def test(self, model, dset_manager, train_loader, valid_loader, test_loader, df_test=None):
        with torch.no_grad():
            train_x, _ = dset_manager.get_transformed_data(train_loader)
            valid_x, _ = dset_manager.get_transformed_data(valid_loader)
            test_x, test_y = dset_manager.get_transformed_data(test_loader)
        return self.test_arrays(model, train_x, valid_x, test_x, test_y, df_test)

    def test_arrays(self, model, train_x, valid_x, test_x, test_y, df_test=None):
        cfg = self.config
        model.eval()
        y = self._labels(test_y)
        with torch.no_grad():
            train_d = get_diffs(train_x, model, batch_size=getattr(cfg, "batch_size", 698))
            valid_d = get_diffs(valid_x, model)
            test_d = get_diffs(test_x, model)
        end = cfg.n_layers + 1 - getattr(cfg, "end_layer_index", -1)
        start = getattr(cfg, "start_layer_index", 0)
        _, b_auroc, b_aupr, b_f1, b_p, b_r = M.get_recon_loss(valid_d[0], test_d[0], y, f1_quantiles=[.90])
        _, s_auroc, s_aupr, s_f1, s_p, s_r = M.get_d_loss(train_d, valid_d, test_d, y, gpu_id=cfg.gpu_id,
                                                          start_layer_index=start, end_layer_index=end, norm_type=2,
                                                          f1_quantiles=[.90])
        _, n_auroc, n_aupr, n_f1, n_p, n_r = M.get_d_norm_loss(train_d, valid_d, test_d, y, cfg, gpu_id=cfg.gpu_id,
                                                               start_layer_index=start, end_layer_index=end, norm_type=2,
                                                               f1_quantiles=[.90])
        row = {'base_auroc': b_auroc, 'sap_auroc': s_auroc, 'nap_auroc': n_auroc,
               'base_f1score': b_f1, 'sap_f1score': s_f1, 'nap_f1score': n_f1,
               'base_precision': b_p, 'sap_precision': s_p, 'nap_precision': n_p,
               'base_recalls': b_r, 'sap_recalls': s_r, 'nap_recalls': n_r,
               'base_aupr': b_aupr, 'sap_aupr': s_aupr, 'nap_aupr': n_aupr}
        df_test = _append_row(df_test, row)
        return (b_auroc, b_aupr), (s_auroc, s_aupr), (n_auroc, n_aupr), df_test

    def score_fast(self, model, train_x, valid_x, test_x, test_y, nap=True):
        """Fused path: per-sample scores never leave the device and diffs are never materialised.

        NAP-fit checkpoint: with ``config.nap_fit`` set to a path, the fit made from ``train_x`` is saved
        there as ``(mu, factor rows, var, mu2, N)`` (``Engine.nap_state_dict``) -- the compact replacement of the raw
        ``train_diffs`` array the reference saves (utils/metric.py:205) and re-SVDs on every offline call
        (test_file/FullTest.py:33-44).  ``train_x=None`` loads that checkpoint instead of refitting."""
        cfg = self.config
        y = self._labels(test_y)
        end = cfg.n_layers + 1 - getattr(cfg, "end_layer_index", -1)
        start = getattr(cfg, "start_layer_index", 0)
        eng = model.eval().engine()
        from .engine import clamp_layer_range
        lo, hi = clamp_layer_range(eng.n_diffs, start, end)
        ckpt = getattr(cfg, "nap_fit", None)
        if nap and train_x is None:
            if not ckpt:
                raise ValueError("score_fast(train_x=None) needs config.nap_fit (a saved NAP-fit checkpoint)")
            eng.load_nap_state_dict(torch.load(ckpt, weights_only=False))
            if eng.nap_range != (lo, hi):
                raise ValueError("the NAP-fit checkpoint covers layers %s, the config selects %s" % (eng.nap_range, (lo, hi)))
        elif nap:
            xt = train_x if isinstance(train_x, torch.Tensor) else torch.from_numpy(np.asarray(train_x))
            eng.nap_fit(xt.to(eng.device).float().reshape(len(xt), -1), lo, hi)
            if ckpt:
                torch.save(eng.nap_state_dict(), ckpt)
        with torch.no_grad():
            sv = get_scores(valid_x, model, start, end, nap=nap)
            st = get_scores(test_x, model, start, end, nap=nap)
        out = {}
        for k in st:
            f1, thr = M.get_f1_score(sv[k], st[k], y, f1_quantiles=[.90])
            out[k] = dict(score=st[k], auroc=M.get_auc_roc(st[k], y), aupr=M.get_auc_prc(st[k], y), f1=f1, threshold=thr)
        return out


def _append_row(df, row):
    try:
        import pandas as pd
    except ImportError:
        return (df or []) + [row]
    new = pd.DataFrame([row])
    return new if df is None else pd.concat([df, new], ignore_index=True)
