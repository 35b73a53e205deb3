"""Training wrapper with the interface of the reference's model.py (Model :11-81): set_input, forward, optimize,
compute_chamfer_loss, test_model, error_log -- on the B200 kernels, the flat-buffer optimizer of dist.py and,
when torch.distributed is initialised, one gradient all-reduce per step.

Kept from the reference: Adam(lr_init, betas (0.9, 0.999)), the step order zero_grad -> train() -> forward ->
Chamfer x log_step_ratio(max_up_ratio / up_ratio) -> backward -> clip_grad_value_(1) -> step (model.py:53-66),
including the fact that the weight is 0 at the full ratio (model.py:72) unless
`weight_full_ratio` is given.  Not kept: the blocking loss.item() of every step (model.py:76) -- the running mean
is accumulated on the device and read when error_log is accessed.
"""
from collections import defaultdict
from math import log

import torch

from . import _lib, level_train
from .dist import FlatAdam
from .model_loss import ChamferLoss


class Model(object):
    def __init__(self, net, phase, opt=None, lr_init=None, ckpt_loader=None, weight_full_ratio=None):
        self.net = net
        self.phase = phase
        self.weight_full_ratio = weight_full_ratio
        lr = lr_init if lr_init is not None else getattr(opt, "lr_init", 5e-4)
        if phase == 'train':
            self._err_sum = defaultdict(lambda: None)
            self._err_cnt = defaultdict(int)
            self.chamfer_criteria = ChamferLoss()
            self.old_lr = lr
            self.lr = lr
            self.optimizer = FlatAdam(self.net, lr=lr, betas=(0.9, 0.999), clip_value=1.0)
        # model.py:25-28: a given --ckpt is ALWAYS loaded (utils.pytorch_utils.load_network); ckpt_loader overrides the loader
        ckpt = getattr(opt, "ckpt", None)
        if ckpt is not None:
            if ckpt_loader is None:
                from .formats import load_network as ckpt_loader
            step = ckpt_loader(self.net, ckpt)
            # the reference stores str(step) (utils/pytorch_utils.py:12) and then divides it (main.py:141): keep an int
            self.step = int(step) if isinstance(step, str) and step.strip().lstrip("-").isdigit() else step
            if phase == 'train':
                _lib.bump_weight_generation()
        else:
            self.step = 0

    @property
    def error_log(self):
        """{"cd_loss_x<ratio>": running mean} like model.py:74-76 (one device->host read per access, not per step)."""
        return {k: float(s / self._err_cnt[k]) for k, s in self._err_sum.items() if s is not None}

    def set_input(self, input_pc, up_ratio, label_pc=None):
        """input_pc Bx3xN, up_ratio int, label_pc Bx3xN'"""
        self.input = input_pc.detach()
        self.up_ratio = up_ratio
        self.gt = label_pc.detach() if label_pc is not None else None

    def forward(self, **kwargs):
        if self.gt is not None:
            self.predicted, self.gt = self.net(self.input, ratio=self.up_ratio, gt=self.gt, **kwargs)
        else:
            self.predicted = self.net(self.input, ratio=self.up_ratio, **kwargs)

    def optimize(self, epoch=None, **kwargs):
        """run forward and backward, apply gradients (model.py:53-66)"""
        self.optimizer.zero_grad()
        self.net.train()
        self.forward(**kwargs)
        loss = self.compute_chamfer_loss(self.predicted, self.gt)
        # the flat optimizer pre-zeroes one gradient buffer that every .grad is a view of: the native backward kernels
        # accumulate straight into it
        prev = level_train.accumulate_into_param_grads
        level_train.accumulate_into_param_grads = True
        try:
            loss.backward()
        finally:
            level_train.accumulate_into_param_grads = prev
        self.optimizer.step()          # [all-reduce] + clip_grad_value_(1) + Adam, one kernel
        self.step += 1
        return loss

    def compute_chamfer_loss(self, pc, pc_label):
        loss_chamfer = self.chamfer_criteria(pc.transpose(1, 2).contiguous(), pc_label.transpose(1, 2).contiguous())
        weight = log(self.net.max_up_ratio / self.up_ratio, self.net.step_ratio)
        if weight == 0 and self.weight_full_ratio is not None:
            weight = self.weight_full_ratio
        loss_chamfer = loss_chamfer * weight
        key = "cd_loss_x{}".format(self.up_ratio)
        d = loss_chamfer.detach()
        self._err_sum[key] = d if self._err_sum[key] is None else self._err_sum[key] + d
        self._err_cnt[key] += 1
        return loss_chamfer

    def test_model(self, **kwargs):
        self.net.eval()
        with torch.no_grad():
            self.forward(**kwargs)
