"""Scalar loss terms of the textual and duration training stages (stage_type.py:230-256, 494-556,
losses.py:430-446).  They act on (B,F) / (B,T) curves and (B,T,16) class scores — a few thousand values — so they
are plain tensor formulas on whatever device the inputs live on (no kernels of ours, nothing on the hot path);
gradients flow into the CUDA graphs of the duration / pitch-energy predictors through their outputs.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def curve_loss(target, prediction):
    """smooth-L1 on the curve + smooth-L1 on its first difference (stage_type.py:238-256: `pitch`, `energy`)"""
    return F.smooth_l1_loss(target, prediction) + F.smooth_l1_loss(torch.diff(target), torch.diff(prediction))


def duration_class_weights(durations_per_class):
    """dataloader.py:51: weight_c = sum(n) / (n_c * classes); DurationLoss takes its square root"""
    d = durations_per_class.to(torch.float32)
    return d.sum() / (d * d.shape[0])


def duration_losses(duration_raw, duration, target_dur, target_class, text_length, class_weight):
    """train_duration (stage_type.py:507-522) + DurationLoss.forward (losses.py:436-446):
    -> (smooth-L1 of the soft durations over each utterance's tokens, averaged over utterances;
        cross entropy of the class scores with weights sqrt(class_weight), averaged over utterances)"""
    B = duration.shape[0]
    w = torch.sqrt(class_weight.to(duration_raw.dtype))
    l1 = duration_raw.new_zeros(())
    ce = duration_raw.new_zeros(())
    for i in range(B):
        n = int(text_length[i])
        l1 = l1 + F.smooth_l1_loss(duration[i, :n], target_dur[i, :n].to(duration.dtype))
        ce = ce + F.cross_entropy(duration_raw[i, :n], target_class[i, :n].long(), weight=w)
    return l1 / B, ce / B
