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

    # zero_grad -> forward -> Chamfer -> backward is a fixed sequence of ~500 launches on fixed shapes: after two eager steps it
    # is captured in a CUDA graph and replayed (the step was bound by the host issuing those launches); the gradient
    # all-reduce and the fused clip + Adam stay outside the graph.  Any keyword argument (test hooks) selects the eager path.
    use_cuda_graph = True
    _GRAPH_WARMUP = 2

    def _forward_backward(self, **kwargs):
        self.optimizer.zero_grad()
        self.net.train()
        self.forward(**kwargs)
        loss = self._weighted_chamfer(self.predicted, self.gt)
        # the flat optimizer pre-zeroes one gradient buffer that every .grad is a view of: the native backward kernels
        # accumulate straight into it
        prev = level_train.accumulate_into_param_grads
        level_train.accumulate_into_param_grads = True
        try:
            loss.backward()
        finally:
            level_train.accumulate_into_param_grads = prev
        return loss

    def _graph_key(self):
        ptrs = hash(tuple(p.data_ptr() for p in self.optimizer.params))
        thr = getattr(self.chamfer_criteria, "_ChamferLoss__threshold", None)
        return (tuple(self.input.shape), tuple(self.gt.shape), int(self.up_ratio), str(self.input.device), ptrs, thr,
                self.weight_full_ratio)

    def _forward_backward_graphed(self):
        cache = self.__dict__.setdefault("_graphs", {})
        key = self._graph_key()
        entry = cache.get(key)
        if entry is None:
            entry = cache[key] = {"seen": 0}
        if "graph" not in entry:
            entry["seen"] += 1
            if entry["seen"] <= self._GRAPH_WARMUP:
                return self._forward_backward()              # eager steps: lazy initialisation happens here
            if len(cache) > 4:
                for k in [k for k in cache if k != key][:len(cache) - 4]:
                    del cache[k]
            dev = self.input.device
            s_in, s_gt = self.input.clone(), self.gt.clone()
            user_in, user_gt = self.input, self.gt
            side = torch.cuda.current_stream(dev)          # optimize() made the side stream current
            self.predicted = None
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            counter = _lib.Profiler(timing=False)
            prev = _lib.set_profiler(counter)
            try:
                self.input, self.gt = s_in, s_gt
                with torch.cuda.graph(graph, stream=side):
                    loss = self._forward_backward()
                entry.update(graph=graph, s_in=s_in, s_gt=s_gt, loss=loss, predicted=self.predicted, gt_out=self.gt,
                             launches=counter.launches)
            finally:
                _lib.set_profiler(prev)
                self.input, self.gt = user_in, user_gt
        entry["s_in"].copy_(self.input)
        entry["s_gt"].copy_(self.gt)
        entry["graph"].replay()
        prof = _lib._profiler
        if prof is not None:
            prof.launches += entry["launches"]
        self.predicted, self.gt = entry["predicted"], entry["gt_out"]
        return entry["loss"]

    def optimize(self, epoch=None, **kwargs):
        """run forward and backward, apply gradients (model.py:53-66)"""
        prof = _lib._profiler
        graphable = (self.use_cuda_graph and not kwargs and self.gt is not None and self.input.is_cuda
                     and not (prof is not None and prof.timing))
        if graphable:
            # warm-up steps, capture and replays all run on one side stream (the autograd nodes of the parameters remember the
            # stream they were created on: an eager step on the caller's stream would tie the captured backward to it)
            dev = self.input.device
            side = self.__dict__.get("_side_stream")
            if side is None or side.device != dev:
                side = self.__dict__["_side_stream"] = torch.cuda.Stream(device=dev)
            cur = torch.cuda.current_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                loss = self._forward_backward_graphed().detach()
                self.predicted = self.predicted.detach()      # nothing keeps the step's autograd graph alive
            cur.wait_stream(side)
        else:
            loss = self._forward_backward(**kwargs)
        self._log_loss(loss)
        self.optimizer.step()          # [all-reduce] + clip_grad_value_(1) + Adam, one kernel
        self.step += 1
        return loss

    def _weighted_chamfer(self, pc, pc_label):
        loss_chamfer = self.chamfer_criteria(pc.transpose(1, 2).contiguous(), pc_label.transpose(1, 2).contiguous())
        weight = log(self.net.max_up_ratio / self.up_ratio, self.net.step_ratio)
        if weight == 0 and self.weight_full_ratio is not None:
            weight = self.weight_full_ratio
        return loss_chamfer * weight

    def _log_loss(self, loss_chamfer):
        key = "cd_loss_x{}".format(self.up_ratio)
        d = loss_chamfer.detach()
        self._err_sum[key] = d.clone() if self._err_sum[key] is None else self._err_sum[key] + d
        self._err_cnt[key] += 1

    def compute_chamfer_loss(self, pc, pc_label):
        """model.py:68-77: Chamfer x log_step_ratio(max_up_ratio / up_ratio), logged under cd_loss_x<ratio>"""
        loss_chamfer = self._weighted_chamfer(pc, pc_label)
        self._log_loss(loss_chamfer)
        return loss_chamfer

    def test_model(self, **kwargs):
        self.net.eval()
        with torch.no_grad():
            self.forward(**kwargs)
