from .fc_layer import FCLayer

__all__ = ["FCLayer"]
