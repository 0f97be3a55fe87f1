"""The cheap sigmoid the reference constrains parameters with (/root/reference/src/jaxns/internals/constraint_bijections.py:11-40)."""
import torch

__all__ = ["quick_unit", "quick_unit_inverse"]


def quick_unit(x):
    """R -> (0, 1): 0.5 (x / (1 + |x|) + 1)."""
    return 0.5 * (x / (1 + torch.abs(x)) + 1)


def quick_unit_inverse(y):
    twoy = y + y
    return torch.where(y >= 0.5, (1 - twoy) / (twoy - 2), 1 - 1 / twoy)
