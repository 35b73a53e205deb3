"""Chamfer loss: same classes as the reference's network/model_loss.py (NmDistanceFunction :5,
nndistance :30, ChamferLoss :33) on the sm_100a nmdistance kernels.

The reference's backward cannot run as shipped (NameError on d_dist1/d_dist2, model_loss.py:22-23,
and the removed ctx.saved_variables :24); this backward is what those lines evidently meant.
"""
import torch

from . import losses


class NmDistanceFunction(torch.autograd.Function):
    """3D point set to 3D point set distance: for every point the squared distance to, and the index of,
    its nearest neighbour in the other set.  xyz1 (B,N,3), xyz2 (B,M,3) float32."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
        B, N, _ = xyz1.size()
        B, M, _ = xyz2.size()
        result = torch.empty(B, N, dtype=xyz1.dtype, device=xyz1.device)
        result_i = torch.empty(B, N, dtype=torch.int32, device=xyz1.device)
        result2 = torch.empty(B, M, dtype=xyz2.dtype, device=xyz2.device)
        result2_i = torch.empty(B, M, dtype=torch.int32, device=xyz2.device)
        losses.nmdistance_forward(xyz1, xyz2, result, result2, result_i, result2_i)
        ctx.save_for_backward(xyz1, xyz2, result_i, result2_i)
        ctx.mark_non_differentiable(result_i, result2_i)
        return result, result_i, result2, result2_i

    @staticmethod
    def backward(ctx, graddist1, gradNone1, graddist2, gradNone2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        gradxyz1 = torch.zeros_like(xyz1)
        gradxyz2 = torch.zeros_like(xyz2)
        losses.nmdistance_backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2)
        return gradxyz1, gradxyz2


nndistance = NmDistanceFunction.apply


class ChamferLoss(torch.nn.Module):
    """Bidirectional nearest-neighbour distance of two point sets (model_loss.py:33-85)."""

    def __init__(self, threshold=None, forward_weight=1.0):
        super(ChamferLoss, self).__init__()
        # only distances below threshold * mean(distance) count (outlier rejection)
        self.__threshold = threshold
        self.forward_weight = forward_weight

    def set_threshold(self, value):
        self.__threshold = value

    def unset_threshold(self):
        self.__threshold = None

    def forward(self, pred, gt):
        assert(pred.dim() == 3 and gt.dim() == 3), \
            "input for ChamferLoss must be a 3D-tensor, but pred.size() is {} gt.size() is {}".format(pred.size(), gt.size())
        if pred.size(2) != 3:
            assert(pred.size(1) == 3), "ChamferLoss is implemented for 3D points"
            pred = pred.transpose(2, 1).contiguous()
        if gt.size(2) != 3:
            assert(gt.size(1) == 3), "ChamferLoss is implemented for 3D points"
            gt = gt.transpose(2, 1).contiguous()
        assert(pred.size(2) == 3 and gt.size(2) == 3), "ChamferLoss is implemented for 3D points"
        pred2gt, _, gt2pred, _ = NmDistanceFunction.apply(pred, gt)
        if self.__threshold is not None:
            fwd_limit = torch.mean(pred2gt, dim=1, keepdim=True) * self.__threshold
            bwd_limit = torch.mean(gt2pred, dim=1, keepdim=True) * self.__threshold
            pred2gt = torch.where(pred2gt < fwd_limit, pred2gt, torch.zeros_like(pred2gt))
            gt2pred = torch.where(gt2pred < bwd_limit, gt2pred, torch.zeros_like(gt2pred))
        pred2gt = torch.mean(pred2gt, dim=1)
        gt2pred = torch.mean(gt2pred, dim=1)
        CD_dist = self.forward_weight * pred2gt + gt2pred
        return torch.mean(CD_dist)
