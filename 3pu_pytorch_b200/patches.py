"""Input side of the train step on the GPU: the reference's H5Dataset.shape_to_patch / augment (data.py:119-172)
and the numpy helpers they call (utils/pc_utils.py:11-79), re-expressed on device tensors so that patch
extraction uses the same hand-written group_knn kernel as the network.

The reference runs group_knn on CPU tensors per item (k = num_patch_point * ratio up to 4992 neighbours over a
5000..80000-point shape, data.py:135-139) and rotates in numpy; with a millisecond-scale train step that loader
becomes the bottleneck.  Semantics are kept: patches are the k nearest neighbours of randomly drawn input points,
input and label share the label's centroid/radius, one random rotation (Rz.Ry.Rx, row-vector convention) per patch.
HDF5 parsing itself (h5py is not in this image) stays out of scope: these functions take arrays.
"""
import math

import torch

from . import operations


def shape_to_patch(input_pc, label_pc, ratio, num_patch_point, batch_size, seed_idx=None, generator=None):
    """data.py:119-142.  input_pc (1,N,3), label_pc (1,r*N,3) CUDA float32 -> (input_patches (B,M,3),
    label_patches (B,r*M,3)), M = num_patch_point.  seed_idx (B,) int64 replaces the random draw (tests)."""
    assert input_pc.dim() == 3 and input_pc.size(0) == 1 and input_pc.size(2) == 3, "input_pc must be (1,N,3)"
    assert label_pc.dim() == 3 and label_pc.size(0) == 1 and label_pc.size(2) == 3, "label_pc must be (1,rN,3)"
    if seed_idx is None:
        seed_idx = torch.randint(0, input_pc.shape[1], (batch_size,), generator=generator)      # np.random.randint, :130
    seed_idx = seed_idx.to(input_pc.device)
    rnd_pts = input_pc[:, seed_idx, :]                                                            # (1,B,3), :131
    # group_knn(..., NCHW=False)[0][0] -> (B,K,3); unique=True is the reference's default (operations.py:165)
    label_patches = operations.group_knn(num_patch_point * ratio, rnd_pts, label_pc, NCHW=False)[0][0]
    input_patches = operations.group_knn(num_patch_point, rnd_pts, input_pc, NCHW=False)[0][0]
    return input_patches, label_patches


def normalize_point_cloud(pc):
    """utils/pc_utils.py:11-25 on tensors: (B,P,3) or (P,3) -> (pc, centroid, furthest_distance)."""
    axis = 0 if pc.dim() == 2 else 1
    centroid = pc.mean(dim=axis, keepdim=True)
    pc = pc - centroid
    furthest = pc.square().sum(dim=-1, keepdim=True).sqrt().amax(dim=axis, keepdim=True)
    return pc / furthest, centroid, furthest


def rotation_matrices(angles):
    """(B,3) Euler angles -> (B,3,3) Rz.Ry.Rx as utils/pc_utils.py:54-64 (points are multiplied from the left: p @ R)."""
    cx, cy, cz = torch.cos(angles).unbind(-1)
    sx, sy, sz = torch.sin(angles).unbind(-1)
    one, zero = torch.ones_like(cx), torch.zeros_like(cx)
    Rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], -1).view(-1, 3, 3)
    Ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], -1).view(-1, 3, 3)
    Rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], -1).view(-1, 3, 3)
    return Rz @ Ry @ Rx


def rotate_point_cloud_and_gt(batch_data, batch_gt=None, angles=None, generator=None):
    """utils/pc_utils.py:45-79: one random rotation per patch applied to data and ground truth."""
    B = batch_data.shape[0]
    if angles is None:
        angles = torch.rand(B, 3, generator=generator) * (2 * math.pi)
    R = rotation_matrices(angles.to(batch_data.device, batch_data.dtype))
    out = torch.matmul(batch_data[..., :3], R)
    gt = torch.matmul(batch_gt[..., :3], R) if batch_gt is not None else None
    return out, gt


def jitter_perturbation_point_cloud(batch_data, sigma=0.005, clip=0.02, generator=None):
    """utils/pc_utils.py:28-42: clipped Gaussian noise per point."""
    noise = torch.randn(batch_data.shape, generator=generator).to(batch_data.device, batch_data.dtype)
    return batch_data + (sigma * noise).clamp_(-clip, clip)


def augment(input_patches, label_patches, jitter=False, jitter_sigma=0.005, jitter_max=0.02, angles=None, generator=None):
    """data.py:144-172 (the live part: optional jitter, shared normalisation by the label, random rotation; the
    reference's drop-out branch is unreachable -- it calls .value on an int, data.py:166 -- and is not mirrored)."""
    if jitter:
        input_patches = jitter_perturbation_point_cloud(input_patches, jitter_sigma, jitter_max, generator)
    label_patches, centroid, furthest = normalize_point_cloud(label_patches)
    input_patches = (input_patches - centroid) / furthest
    return rotate_point_cloud_and_gt(input_patches, label_patches, angles=angles, generator=generator)
