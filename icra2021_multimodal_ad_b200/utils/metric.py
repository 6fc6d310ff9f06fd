"""utils/metric.py of the reference on device.

Same functions, signatures and return tuples: ``get_norm``, ``get_auc_roc``, ``get_nap_auc_roc``,
``get_threshold``, ``get_confusion_matrix``, ``get_auc_prc``, ``get_f1_score``, ``get_recon_loss``,
``get_d_loss`` (SAP), ``get_d_norm_loss`` (NAP).  Score arrays may be numpy (reference behaviour) or
CUDA tensors; curves, quantile and confusion counts are computed by libmmad (bit-identical to
scikit-learn / NumPy given identical fp32 scores).  Like the reference, metric wrappers swallow
exceptions and return ``.0`` (utils/metric.py:43-44,62-63,115-116).
"""
import ctypes as C
import time

import numpy as np
import torch

from .. import _lib
from ..ops import row_mean_sq
from .normalize import Rotater, Standardizer, Truncater  # noqa: F401  (re-exported like the reference)

_ws_cache = {}


def _dev():
    if not torch.cuda.is_available():
        raise _lib.MmadError("metrics need a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev_f32(a):
    if isinstance(a, torch.Tensor):
        return a.detach().to(_dev() if not a.is_cuda else a.device, torch.float32).contiguous().reshape(-1)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)).reshape(-1)).to(_dev())


def _to_dev_label(y, device):
    if isinstance(y, torch.Tensor):
        return (y.detach().to(device) != 0).to(torch.uint8).contiguous().reshape(-1)
    return torch.from_numpy(np.ascontiguousarray((np.asarray(y) == 1) | (np.asarray(y) == True)).astype(np.uint8).reshape(-1)).to(device)  # noqa: E712


def _ws(n, device):
    need = _lib.lib().mmad_metric_workspace_bytes(int(n))
    key = str(device)
    t = _ws_cache.get(key)
    if t is None or t.numel() < need:
        t = torch.empty(need, dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _curve(fn_name, score, label):
    s = _to_dev_f32(score)
    y = _to_dev_label(label, s.device)
    if s.numel() != y.numel():
        raise ValueError("inconsistent lengths")
    if s.numel() == 0:
        raise ValueError("empty input")
    out = C.c_double()
    ws = _ws(s.numel(), s.device)
    with torch.cuda.device(s.device):
        _lib.check(getattr(_lib.lib(), fn_name)(s.data_ptr(), y.data_ptr(), s.numel(), C.byref(out), ws.data_ptr(),
                                                ws.numel(), _stream()))
    return float(out.value)


def get_norm(x, norm_type=2):
    return abs(x) ** norm_type


def get_auc_roc(score, test_label, nap=False):
    try:
        auc = _curve("mmad_auc_roc", score, test_label)
        if nap is True:
            print('auroc', auc)
        return auc
    except Exception:
        return .0


def get_nap_auc_roc(score, test_label, nap=False):
    try:
        return _curve("mmad_auc_roc", score, test_label)
    except Exception:
        return .0


def get_threshold(precisions, recalls, threshold):
    # utils/metric.py:64-81 (unused by the reference: its call site is commented out, line 100)
    index_cnt = [cnt for cnt, (p, r) in enumerate(zip(precisions, recalls)) if p == r][0]
    print('precision: ', precisions[index_cnt], ', recall: ', recalls[index_cnt])
    threshold_fixed = threshold[index_cnt]
    print('threshold: ', threshold_fixed)
    return threshold_fixed


def confusion_counts(score, test_label, threshold, strict):
    """(tp, fp, fn, tn) with pred = score > thr (strict) or score >= thr."""
    s = _to_dev_f32(score)
    y = _to_dev_label(test_label, s.device)
    counts = (C.c_longlong * 4)()
    ws = _ws(s.numel(), s.device)
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib().mmad_confusion(s.data_ptr(), y.data_ptr(), s.numel(), float(threshold), int(strict), counts,
                                             ws.data_ptr(), ws.numel(), _stream()))
    return tuple(int(c) for c in counts)


def get_confusion_matrix(score, test_label, threshold):
    # utils/metric.py:83-95: pred = score >= threshold (note >=), sklearn confusion_matrix, prints
    tp, fp, fn, tn = confusion_counts(score, test_label, threshold, strict=False)
    print('Tn, Fp : ' + str(tn) + ', ' + str(fp) + '\nFn, Tp : ' + str(fn) + ', ' + str(tp))
    with np.errstate(all="ignore"):
        precision = np.float64(tp) / np.float64(tp + fp)
        recall = np.float64(tp) / np.float64(tp + fn)
    return precision, recall


def get_auc_prc(score, test_label):
    try:
        return _curve("mmad_auc_prc", score, test_label)
    except Exception:
        return .0


def quantile(valid_score, q):
    """np.quantile(valid_score, q) for an fp32 array (NumPy-2 fp32 arithmetic) on device -> np.float32."""
    v = _to_dev_f32(valid_score)
    out = C.c_float()
    ws = _ws(v.numel(), v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().mmad_quantile(v.data_ptr(), v.numel(), float(q), C.byref(out), ws.data_ptr(), ws.numel(),
                                            _stream()))
    return np.float32(out.value)


def get_f1_score(valid_score, test_score, test_label, f1_quantiles=[.99]):
    f1_quantiles = 0.90  # added (utils/metric.py:120: the argument is overwritten)
    threshold = quantile(valid_score, f1_quantiles)
    tp, fp, fn, tn = confusion_counts(test_score, test_label, threshold, strict=True)
    with np.errstate(all="ignore"):
        p = np.float64(tp) / float(tp + fp)
        r = np.float64(tp) / float(tp + fn)
        f1s = p * r * 2 / (p + r)
    return f1s, threshold


def _mean_sq(d):
    """(d**2).mean(axis=1) -> numpy fp32 when given numpy (reference behaviour), tensor when given a tensor."""
    if isinstance(d, torch.Tensor):
        return row_mean_sq(d if d.is_cuda else d.to(_dev()))
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(d, dtype=np.float32))).to(_dev())
    return row_mean_sq(t).cpu().numpy()


def get_recon_loss(valid_diff, test_diff, test_label, f1_quantiles=[.99]):
    # utils/metric.py:132-143
    loss = _mean_sq(test_diff)
    loss_auc_roc = get_auc_roc(loss, test_label)
    loss_auc_prc = get_auc_prc(loss, test_label)
    loss_f1s, threshold = get_f1_score(_mean_sq(valid_diff), loss, test_label, f1_quantiles=f1_quantiles)
    precision, recall = get_confusion_matrix(loss, test_label, threshold)
    print('base threshold', threshold)
    return loss, loss_auc_roc, loss_auc_prc, loss_f1s, precision, recall


def _clamp(n_diffs, start_layer_index, end_layer_index):
    # utils/metric.py:155-162 / 195-202
    if end_layer_index is None:
        end_layer_index = n_diffs + 1
    if start_layer_index > n_diffs - 1:
        start_layer_index = n_diffs - 1
    if end_layer_index - start_layer_index < 1:
        end_layer_index = start_layer_index + 1
    return start_layer_index, end_layer_index


def _concat(diffs, start, end):
    sel = diffs[start:end]
    if isinstance(sel[0], torch.Tensor):
        return torch.cat([d if d.is_cuda else d.to(_dev()) for d in sel], dim=-1)
    return torch.from_numpy(np.concatenate([np.asarray(d, dtype=np.float32) for d in sel], axis=-1)).to(_dev())


def get_d_loss(train_diffs, valid_diffs, test_diffs, test_label, start_layer_index=0, end_layer_index=None, gpu_id=-1,
               norm_type=2, f1_quantiles=[.99]):
    """SAP, utils/metric.py:145-181."""
    start_layer_index, end_layer_index = _clamp(len(test_diffs), start_layer_index, end_layer_index)
    as_np = not isinstance(test_diffs[0], torch.Tensor)
    valid_cat = _concat(valid_diffs, start_layer_index, end_layer_index)
    test_cat = _concat(test_diffs, start_layer_index, end_layer_index)
    d_loss = row_mean_sq(test_cat)
    d_loss_auc_roc = get_auc_roc(d_loss, test_label)
    d_loss_auc_prc = get_auc_prc(d_loss, test_label)
    d_loss_f1s, threshold = get_f1_score(row_mean_sq(valid_cat), d_loss, test_label, f1_quantiles=f1_quantiles)
    print()
    precision, recall = get_confusion_matrix(d_loss, test_label, threshold)
    if as_np:
        d_loss = d_loss.cpu().numpy()
    return d_loss, d_loss_auc_roc, d_loss_auc_prc, d_loss_f1s, precision, recall


def get_d_norm_loss(train_diffs, valid_diffs, test_diffs, test_label, config, start_layer_index=0, end_layer_index=None,
                    gpu_id=-1, norm_type=2, f1_quantiles=[.99]):
    """NAP, utils/metric.py:183-238: Rotater/Standardizer fitted on the train diffs, score =
    mean_j ((x-mu)V - mu2)_j^2 / var_j.  ``config.train_diffs`` (if set) receives the concatenated
    train diffs like the reference (line 205)."""
    start_layer_index, end_layer_index = _clamp(len(test_diffs), start_layer_index, end_layer_index)
    as_np = not isinstance(test_diffs[0], torch.Tensor)
    train_cat = _concat(train_diffs, start_layer_index, end_layer_index)
    if getattr(config, "train_diffs", None):
        torch.save(train_cat.cpu().numpy(), config.train_diffs)
    valid_cat = _concat(valid_diffs, start_layer_index, end_layer_index)
    start_data = time.time()
    test_cat = _concat(test_diffs, start_layer_index, end_layer_index)
    total = time.time() - start_data

    rotater = Rotater()
    stndzer = Standardizer()
    rotater.fit(train_cat, gpu_id=gpu_id)
    stndzer.fit(rotater.run(train_cat, gpu_id=gpu_id, as_tensor=True))

    valid_rot = stndzer.run(rotater.run(valid_cat, gpu_id=gpu_id, as_tensor=True), as_tensor=True)
    start_data = time.time()
    test_rot = stndzer.run(rotater.run(test_cat, gpu_id=gpu_id, as_tensor=True), as_tensor=True)
    score = row_mean_sq(test_rot)
    total += time.time() - start_data
    print('nap cal', total)
    auc_roc = get_auc_roc(score, test_label, nap=True)
    auc_prc = get_auc_prc(score, test_label)
    f1_scores, threshold = get_f1_score(row_mean_sq(valid_rot), score, test_label, f1_quantiles=f1_quantiles)
    precision, recall = get_confusion_matrix(score, test_label, threshold)
    print('nap threshold', threshold)
    if as_np:
        score = score.cpu().numpy()
    return score, auc_roc, auc_prc, f1_scores, precision, recall
