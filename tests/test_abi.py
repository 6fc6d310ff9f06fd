"""CPU checks of the boundary: the library loads, exports every symbol include/mmad.h declares,
and the host layer mirrors the reference's API surface (no compute calls here)."""
import argparse
import ctypes
import os
import re

import pytest
import torch

from icra2021_multimodal_ad_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "mmad.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(mmad_[a-z0-9_]+)\s*\(", txt))
    names -= {"mmad_allreduce_fn"}
    return names


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mmad.h but not exported"
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    assert _lib.lib().mmad_version() >= 100


def test_create_rejects_bad_descriptors_without_gpu():
    L = _lib.lib()
    h = ctypes.c_void_p()
    d = _lib.Desc()
    d.n_enc, d.n_dec = 0, 1
    assert L.mmad_create(ctypes.byref(d), ctypes.byref(h)) == -1
    assert b"layer" in L.mmad_last_error()
    d.n_enc, d.n_dec = 1, 1
    d.enc_widths[0], d.enc_widths[1] = 8, 4
    d.dec_widths[0], d.dec_widths[1] = 4, 9
    assert L.mmad_create(ctypes.byref(d), ctypes.byref(h)) == -1


def test_model_surface_matches_reference():
    from icra2021_multimodal_ad_b200.model_builder import ae_wrapper, get_model
    cfg = argparse.Namespace(input_size=1728, btl_size=100, n_layers=5, gpu_id=-1)
    m = get_model(cfg)
    keys = list(m.state_dict().keys())
    assert keys[:7] == ["encoder.net.0.layer.weight", "encoder.net.0.layer.bias", "encoder.net.0.bn.weight",
                        "encoder.net.0.bn.bias", "encoder.net.0.bn.running_mean", "encoder.net.0.bn.running_var",
                        "encoder.net.0.bn.num_batches_tracked"]
    assert "encoder.net.4.layer.bias" in keys and "encoder.net.4.bn.weight" not in keys
    assert sum(p.numel() for p in m.parameters()) == 10225670
    assert [l.layer.out_features for l in m.encoder.layer_list] == [1402, 1076, 751, 425, 100]
    assert [l.layer.out_features for l in m.decoder.layer_list] == [425, 751, 1076, 1402, 1728]
    assert isinstance(m.encoder.layer_list, list) and len(list(m.parameters())) == 36
    for name in ("encode", "decode", "forward", "get_loss_value", "step", "validate", "attach",
                 "get_all_optimizers_state_dicts"):
        assert hasattr(m, name)
    assert m.optimizer_list == []
    m2 = ae_wrapper(argparse.Namespace(input_size=(3, 8, 8), btl_size=10, n_layers=3, gpu_id=-1))
    assert m2.encoder.widths == [192, 131, 70, 10]


def test_no_cpu_fallback():
    from icra2021_multimodal_ad_b200.model_builder import get_model
    m = get_model(argparse.Namespace(input_size=64, btl_size=100, n_layers=5, gpu_id=-1)).eval()
    with pytest.raises(_lib.MmadError):
        m(torch.rand(4, 64))


def test_reference_error_behaviour():
    from icra2021_multimodal_ad_b200.modules import FCModule
    with pytest.raises(Exception, match="Either batch_norm or dropout"):
        FCModule(8, 4, [6], use_batch_norm=True, dropout_p=0.5)


def test_layer_range_clamp_matches_reference():
    from icra2021_multimodal_ad_b200.engine import clamp_layer_range
    from oracle.rapp_oracle import clamp_layer_range as ref
    for n in (4, 6):
        for lo in range(0, 10):
            for hi in list(range(-2, 10)) + [None]:
                a, b = ref(n, lo, hi)
                exp = list(range(n))[a:b]
                l2, h2 = clamp_layer_range(n, lo, hi)
                assert list(range(l2, h2)) == exp or (exp == [] and h2 <= l2)


def test_communicator_entry_points_without_gpu():
    """The library resolves libnccl from the process (torch ships it): a unique id can be produced on a CPU-only
    machine; argument errors are reported without touching a device."""
    L = _lib.lib()
    buf = (ctypes.c_ubyte * 128)()
    assert L.mmad_comm_unique_id(buf) == 0
    assert any(bytes(buf))
    assert L.mmad_comm_unique_id(None) == -1
    assert L.mmad_comm_world(None) == 1
    assert L.mmad_comm_init(None, buf, 0, 1) == -1
    assert L.mmad_comm_allreduce_f32(None, None, 4, None) == -1


def test_precision_modes_match_the_header():
    """The host layer's precision names map to the MMAD_PREC_* values of include/mmad.h; a precision outside the
    enum is rejected by mmad_create before anything touches a GPU."""
    txt = open(os.path.join(ROOT, "include", "mmad.h")).read()
    enum = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"#define MMAD_PREC_([A-Z0-9]+)\s+(\d+)", txt)}
    assert enum == _lib.PREC == {"fp32": 0, "f16x3": 1, "f16": 2, "f16f8": 3}
    L = _lib.lib()
    h = ctypes.c_void_p()
    d = _lib.Desc()
    d.n_enc = d.n_dec = 1
    d.enc_widths[0], d.enc_widths[1] = 8, 4
    d.dec_widths[0], d.dec_widths[1] = 4, 8
    d.precision = max(enum.values()) + 1
    assert L.mmad_create(ctypes.byref(d), ctypes.byref(h)) == -1
    assert b"precision" in L.mmad_last_error()


def test_round2_entry_points_reject_bad_arguments_without_gpu():
    """Argument checking of the entry points added in round 2 happens before any device call."""
    L = _lib.lib()
    out = ctypes.c_float()
    assert L.mmad_train_loss(None, ctypes.byref(out)) == -1
    hb = (ctypes.c_ubyte * 64)()
    ptr = ctypes.c_void_p()
    assert L.mmad_peer_grad_alloc(None, 16, ctypes.byref(ptr), hb) == -1
    assert L.mmad_peer_grad_open(None, hb) == -1
    assert L.mmad_peer_close(None) == 0
    assert L.mmad_set_option(None, b"nap_passes", 4.0) == -1


def test_step_loss_falls_back_to_float_for_foreign_tensors():
    """train.step_loss reads the doorbell only for the tensor the fused step itself returned; any other tensor (a scaled
    loss, a validation loss, a CPU tensor) is read the ordinary way."""
    import types
    import torch
    from icra2021_multimodal_ad_b200 import train as T
    model = types.SimpleNamespace()                       # no _train_state at all
    assert T.step_loss(model, torch.tensor(2.5)) == 2.5
    model._train_state = types.SimpleNamespace(last_loss=torch.tensor(1.0))
    assert T.step_loss(model, torch.tensor(3.5)) == 3.5   # not the tensor of the last step


def test_device_pointer_wrapper_describes_a_flat_fp32_buffer():
    """The library-owned, peer-mapped gradient buffer reaches torch through __cuda_array_interface__ (zero copy)."""
    from icra2021_multimodal_ad_b200.train import _DevicePtr
    cai = _DevicePtr(0x7f0000000000, 1024).__cuda_array_interface__
    assert cai["shape"] == (1024,) and cai["typestr"] == "<f4" and cai["data"] == (0x7f0000000000, False) and cai["version"] == 2
