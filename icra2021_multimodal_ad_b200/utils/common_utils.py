"""utils/common_utils.py of the reference (shape helper only)."""


def get_hidden_layer_sizes(start_size, end_size, n_hidden_layers):
    """Reference utils/common_utils.py:22-31: linear interpolation of layer widths with
    truncating ``int()``; handles growing and shrinking sizes."""
    step = (start_size - end_size) / (n_hidden_layers + 1)
    return [int(start_size - step * (k + 1)) for k in range(n_hidden_layers)]
