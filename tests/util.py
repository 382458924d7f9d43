import torch


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(x):  # (N,C,H,W) -> contiguous (N,H,W,C)
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):  # (N,H,W,C) -> contiguous (N,C,H,W)
    return x.permute(0, 3, 1, 2).contiguous()


def bf16_round(x):
    return x.to(torch.bfloat16).float()
