"""`from monai.transforms import ToTensor` (src/py/utils.py:16-18): unused by the executed code."""


class ToTensor(object):
    def __init__(self, dtype=None, device=None):
        self.dtype, self.device = dtype, device

    def __call__(self, x):
        import torch
        return torch.as_tensor(x, dtype=self.dtype, device=self.device)
