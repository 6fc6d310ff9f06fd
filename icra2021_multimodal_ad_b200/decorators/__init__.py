from .variational_info_bottleneck import variational_info_bottleneck

__all__ = ["variational_info_bottleneck"]
