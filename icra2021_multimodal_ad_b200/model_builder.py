"""model_builder.py of the reference (ae_wrapper / get_model), same config fields:
input_size, btl_size, n_layers, gpu_id (+ optional ``precision``)."""
from .modules import FCModule, Loss
from .utils.common_utils import get_hidden_layer_sizes


def ae_wrapper(config):
    from .models.auto_encoder import AutoEncoder
    input_size, btl_size, n_layers = config.input_size, config.btl_size, config.n_layers
    if type(input_size) != int:          # (C, H, W) -> flatten, model_builder.py:15-19
        C, H, W = input_size
        input_size = C * H * W
    encoder = FCModule(input_size=input_size, output_size=btl_size,
                       hidden_sizes=get_hidden_layer_sizes(input_size, btl_size, n_hidden_layers=n_layers - 1),
                       use_batch_norm=True, act="leakyrelu", last_act=None)
    decoder = FCModule(input_size=btl_size, output_size=input_size,
                       hidden_sizes=get_hidden_layer_sizes(btl_size, input_size, n_hidden_layers=n_layers - 1),
                       use_batch_norm=True, act="leakyrelu", last_act=None)
    return AutoEncoder(encoder=encoder, decoder=decoder, recon_loss=Loss("mse", reduction="sum"),
                       precision=getattr(config, "precision", "fp32"))


def vib_ae_wrapper(config):
    """VIB variant (BASELINE configs[3]): encoder output 2 * btl_size = (mu | logvar)."""
    from .models.vib_auto_encoder import VIBAutoEncoder
    input_size, btl_size, n_layers = config.input_size, config.btl_size, config.n_layers
    if type(input_size) != int:
        C, H, W = input_size
        input_size = C * H * W
    encoder = FCModule(input_size=input_size, output_size=2 * btl_size,
                       hidden_sizes=get_hidden_layer_sizes(input_size, 2 * btl_size, n_hidden_layers=n_layers - 1),
                       use_batch_norm=True, act="leakyrelu", last_act=None)
    decoder = FCModule(input_size=btl_size, output_size=input_size,
                       hidden_sizes=get_hidden_layer_sizes(btl_size, input_size, n_hidden_layers=n_layers - 1),
                       use_batch_norm=True, act="leakyrelu", last_act=None)
    return VIBAutoEncoder(encoder=encoder, decoder=decoder, recon_loss=Loss("mse", reduction="sum"),
                          beta_kl=getattr(config, "beta_kl", 1.0), precision=getattr(config, "precision", "fp32"))


def get_model(config):
    model = vib_ae_wrapper(config) if getattr(config, "vib", False) else ae_wrapper(config)
    if config.gpu_id >= 0:
        model = model.cuda(config.gpu_id)
    return model
