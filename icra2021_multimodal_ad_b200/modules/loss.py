"""modules/loss.py of the reference.  Only Loss('mse', reduction='sum') is used by the path
(model_builder.py:42); it is evaluated inside the fused kernels."""
import torch.nn as nn


class Loss(nn.Module):
    def __init__(self, loss, weight=None, reduction="sum"):
        super().__init__()
        if loss != "mse" or reduction != "sum" or weight is not None:
            raise NotImplementedError("only Loss('mse', reduction='sum') is on the accelerated path "
                                      "(model_builder.py:42)")
        self.loss_name, self.reduction = loss, reduction

    def is_classification_task(self):
        return False

    def forward(self, y_hat, y):
        from ..ops import mse_sum
        return mse_sum(y_hat, y)
