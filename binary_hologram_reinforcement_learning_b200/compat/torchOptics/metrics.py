"""tm.get_PSNR (reference env.py:132,174): 10 log10(1 / mse), peak 1.0, returned as a python float."""
import math

import torch


def get_PSNR(a, b) -> float:
    a = torch.as_tensor(a)
    b = torch.as_tensor(b, device=a.device)
    mse = torch.mean((a.float() - b.float()) ** 2).item()
    return float(10.0 * math.log10(1.0 / mse))
