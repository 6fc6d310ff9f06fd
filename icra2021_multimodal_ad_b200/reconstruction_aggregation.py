"""reconstruction_aggregation.py of the reference: ``get_diffs(x, model, batch_size)``.

Returns the same list of ``n_layers + 1`` float32 ndarrays ``[N, w_l]``.  One fused device pass
per chunk produces every diff (encoder(x) evaluated once, diffs formed in the GEMM epilogues);
``batch_size`` only bounds the chunk, results do not depend on it (eval mode)."""
import numpy as np
import torch


def get_diffs(x, model, batch_size=698):
    model.eval()
    if isinstance(x, np.ndarray):
        x = torch.tensor(x)
    eng = model.engine()
    widths = eng.enc_widths
    chunk = max(int(batch_size), 16384)
    outs = []
    for xb in x.split(chunk):
        xb = xb.to(eng.device).float()
        xb = xb.reshape(xb.size(0), -1)
        d = eng.score(xb, 0, eng.n_diffs, base=False, sap=False, diffs=True)["diffs"]
        outs.append(d.cpu())
    cat = torch.cat(outs, dim=0) if outs else torch.empty(0, sum(widths))
    return [np.ascontiguousarray(t.numpy()) for t in cat.split(widths, dim=1)]


def get_scores(x, model, start_layer_index=0, end_layer_index=None, nap=False):
    """Fused fast path (no diff materialisation): base and SAP (and NAP when a fit is installed)
    scores of ``x`` as device tensors.  Layer range follows utils/metric.py:155-162."""
    from .engine import clamp_layer_range
    model.eval()
    eng = model.engine()
    lo, hi = clamp_layer_range(eng.n_diffs, start_layer_index, end_layer_index)
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    x = x.to(eng.device).float().reshape(x.shape[0], -1)
    return eng.score(x, lo, hi, base=True, sap=True, nap=nap)
