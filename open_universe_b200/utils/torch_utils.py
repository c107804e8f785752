"""Reference ``utils/torch_utils.py`` helpers used around the hot path."""


def count_parameters(module):
    return sum(p.numel() for p in module.parameters())


def pad_dim_right(a, x):
    """View ``a`` (batch-leading) so that it broadcasts over the trailing dims of ``x``
    (reference utils/torch_utils.py:67-72)."""
    if a.shape != x.shape[: a.ndim]:
        raise ValueError("All left dimensions of a and x should be matching")
    return a[(...,) + (None,) * (x.ndim - a.ndim)]
