"""models/abstract_model.py of the reference."""
from collections.abc import Iterable

import torch.nn as nn


class AbstractModel(nn.Module):
    def __init__(self, *model_and_opts):
        super().__init__()
        for model_and_opt in model_and_opts:
            if not (model_and_opt is None or isinstance(model_and_opt, Iterable)):
                raise Exception("model_and_opt arg should be None or iterable objects")
        self.optimizer_list = []   # for saving best models' optimizers

    def forward(self):
        raise NotImplementedError

    def get_loss_value(self, x, y, *args, **kwargs):
        raise NotImplementedError

    def get_all_optimizers_state_dicts(self):
        return [opt.state_dict() for opt in self.optimizer_list]
