from .activation import Activation
from .loss import Loss
from .fc_module import FCModule

__all__ = ["Activation", "Loss", "FCModule"]
