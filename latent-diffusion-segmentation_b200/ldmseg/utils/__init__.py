"""`ldmseg.utils` surface needed by the sampling hot path: only `OutputDict`
(reference: ldmseg/utils/utils.py:26-31).  Logging / visualisation / config helpers of the
reference are out of scope."""
from collections import OrderedDict


class OutputDict(OrderedDict):
    """OrderedDict whose item assignment is mirrored to attributes (`out.sample` and `out['sample']`)."""

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        super().__setattr__(key, value)


__all__ = ["OutputDict"]
