"""`ldmseg.models.Upscaler` -- name kept for import compatibility with
/root/reference/ldmseg/models/__init__.py:4.  The reference never instantiates it from any
tools/*.py entry point, so it is out of the sampling hot path and not built."""
import torch.nn as nn


class Upscaler(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("Upscaler is unused by the LDMSeg entry points and out of scope of the B200 hot path")
