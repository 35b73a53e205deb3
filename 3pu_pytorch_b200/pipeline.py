"""Whole-shape inference: the reference's main.pc_prediction + the resampling of main.test (main.py:214-246,
375-380), with all patches of the shape upsampled in one batched Net.forward instead of a Python loop of
B=1 calls.  BASELINE config 5: FPS seed extraction + kNN patch extraction + 16x upsample + final FPS.
"""
import torch

from . import operations
from .dist import upsample_sharded


def pc_prediction(net, input_pc, num_point=312, patch_num_ratio=3, up_ratio=16, sharded=False):
    """main.py:214-246.  input_pc 1x3xN -> (patches (P,3,num_point) normalised, upsampled (P,3,num_point*up_ratio)
    in the frame of the input cloud).  sharded=True splits the patches over the ranks of torch.distributed."""
    assert input_pc.dim() == 3 and input_pc.size(0) == 1
    num_patches = int(input_pc.shape[2] / num_point * patch_num_ratio)
    _, seeds = operations.furthest_point_sample(input_pc, num_patches, NCHW=True)
    patches, _, _ = operations.group_knn(num_point, seeds, input_pc, NCHW=True)        # 1,3,P,num_point (unique=True default)
    patches = patches[0].permute(1, 0, 2).contiguous()                                  # P,3,num_point
    patches, centroid, radius = operations.normalize_point_batch(patches, NCHW=True)
    if sharded:
        _, _, up = upsample_sharded(net, patches, ratio=up_ratio, gather=True)
    else:
        with torch.no_grad():
            up = net(patches, ratio=up_ratio)
    return patches, up * radius + centroid


def upsample_shape(net, input_pc, num_point=312, patch_num_ratio=3, up_ratio=16, num_out=None, sharded=False):
    """main.py:362-380: upsample every patch, concatenate, resample to N*up_ratio points by FPS.  Returns 1x3xM."""
    _, up = pc_prediction(net, input_pc, num_point, patch_num_ratio, up_ratio, sharded)
    pred = up.permute(1, 0, 2).reshape(1, 3, -1)                                        # torch.cat(pred_pc_list, dim=-1)
    num_out = num_out or int(input_pc.shape[2]) * up_ratio
    _, pred = operations.furthest_point_sample(pred.contiguous(), num_out, NCHW=True)
    return pred
