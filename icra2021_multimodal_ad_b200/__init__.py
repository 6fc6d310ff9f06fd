"""B200-native (sm_100a) implementation of the autoencoder + RaPP hot path of
Yoo-Youngjae/ICRA2021_multimodal_ad behind the reference's own Python API:

    from icra2021_multimodal_ad_b200.model_builder import get_model
    from icra2021_multimodal_ad_b200.reconstruction_aggregation import get_diffs
    from icra2021_multimodal_ad_b200.utils.metric import get_recon_loss, get_d_loss, get_d_norm_loss

All arithmetic runs in libmmad.so (include/mmad.h); there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
