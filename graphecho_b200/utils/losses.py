"""Mirror of utils/losses.py (make_one_hot, BinaryDiceLoss, DiceLoss) on device tensors."""
import torch
import torch.nn.functional as F
from torch import nn


def make_one_hot(input, num_classes):
    shape = list(input.shape)
    shape[1] = num_classes
    return torch.zeros(shape, device=input.device).scatter_(1, input.long(), 1)


class BinaryDiceLoss(nn.Module):
    def __init__(self, smooth=1, p=2, reduction="mean"):
        super().__init__()
        self.smooth, self.p, self.reduction = smooth, p, reduction

    def forward(self, predict, target):
        assert predict.shape[0] == target.shape[0], "predict & target batch size don't match"
        predict = predict.contiguous().view(predict.shape[0], -1)
        target = target.contiguous().view(target.shape[0], -1)
        num = (predict * target).sum(1) + self.smooth
        den = (predict.pow(self.p) + target.pow(self.p)).sum(1) + self.smooth
        loss = 1 - num / den
        if self.reduction == "mean":
            return loss.mean()
        if self.reduction == "sum":
            return loss.sum()
        if self.reduction == "none":
            return loss
        raise Exception(f"Unexpected reduction {self.reduction}")


class DiceLoss(nn.Module):
    def __init__(self, weight=None, ignore_index=None, **kwargs):
        super().__init__()
        self.kwargs, self.weight, self.ignore_index = kwargs, weight, ignore_index

    def forward(self, predict, target):
        assert predict.shape == target.shape, "predict & target shape do not match"
        dice = BinaryDiceLoss(**self.kwargs)
        prob = F.softmax(predict.float(), dim=1)
        total = 0
        for i in range(target.shape[1]):
            if i != self.ignore_index:
                term = dice(prob[:, i], target[:, i])
                if self.weight is not None:
                    term = term * self.weight[i]
                total = total + term
        return total / target.shape[1]
