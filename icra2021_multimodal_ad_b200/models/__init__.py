from .auto_encoder import AutoEncoder

__all__ = ["AutoEncoder"]
