"""Multi-GPU plumbing of the path (one process per GPU, torch.distributed) and the
train-step tail of the reference's model.py:53-66.

The reference has no distributed code at all.  The path shards by independent units:
  * inference: input patches are independent requests (main.py:237-244) -> contiguous shards, no data-path
    collective; an all_gather only when one rank needs every result (whole-shape FPS, main.py:375-380).
  * training: patches shard along the batch, the model is replicated, and the ONLY exchange is the gradient:
    one all-reduce of one flat 1.2 MB fp32 buffer per step, then clip_grad_value_ + Adam fused in one kernel
    (csrc/optim.cu).  No BatchNorm is ever enabled (upsampler.py:209-230) and ChamferLoss is a per-sample mean
    followed by a batch mean (model_loss.py:80-84), so the mean of equal-shard gradients IS the big-batch gradient.
Works with any backend for the collectives (tests run world_size 2 on gloo/CPU); the fused optimizer kernel needs
CUDA and raises otherwise.
"""
import torch
import torch.distributed as dist

from . import _lib


def shard_range(total, rank, world):
    """Contiguous balanced shard [lo, hi) of `total` units for `rank` of `world` (first total % world get one more)."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def upsample_sharded(net, patches, ratio=None, gather=False):
    """Eval: every rank upsamples its shard of `patches` (P,3,N) (the full list is given on every rank).
    gather=False -> (lo, hi, result of the shard); gather=True -> the full (P,3,N*ratio) on every rank."""
    rank, world = _world()
    lo, hi = shard_range(patches.shape[0], rank, world)
    with torch.no_grad():
        out = net(patches[lo:hi], ratio=ratio) if hi > lo else patches.new_zeros(0, 3, 0)
    if not gather or world == 1:
        return lo, hi, out
    npts = torch.tensor([out.shape[2]], device=patches.device, dtype=torch.int64)
    dist.all_reduce(npts, op=dist.ReduceOp.MAX)
    chunks = []
    for r in range(world):
        rlo, rhi = shard_range(patches.shape[0], r, world)
        chunks.append(patches.new_empty(rhi - rlo, 3, int(npts.item())))
    mine = out if out.shape[0] else patches.new_empty(0, 3, int(npts.item()))
    for r in range(world):  # shards may differ by one patch: broadcast per rank instead of all_gather
        buf = mine.contiguous() if r == rank else chunks[r]
        if buf.numel():
            dist.broadcast(buf, src=r)
        chunks[r] = buf
    return 0, patches.shape[0], torch.cat(chunks, dim=0)


class FlatAdam:
    """Parameters, gradients and Adam state of a module in ONE flat fp32 buffer each.

    step() = [all-reduce of the flat gradient] -> clip_grad_value_ -> Adam, the order of model.py:63-65 (clipping
    happens AFTER averaging).  Parameters and .grad of the module become views into the flat buffers, so autograd
    accumulates straight into the all-reduce / optimizer input: no bucketing, no copies."""

    def __init__(self, module, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, clip_value=1.0, process_group=None,
                 data_parallel=True):
        """process_group: the group the gradient is averaged over (None = the default group, looked up at every
        step, so a group that has been destroyed means single-rank).  data_parallel=False: never all-reduce, even
        inside an initialised job (a rank-local model, e.g. a side measurement on rank 0 only)."""
        self.process_group, self.data_parallel = process_group, bool(data_parallel)
        self.reduce_events = None        # set to [] to collect (start, end) CUDA events around every all-reduce
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if dt != torch.float32:
            raise ValueError("FlatAdam expects float32 parameters")
        ALIGN = 32                                   # every parameter starts on a 128-byte boundary of the flat buffer
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.flat_param = torch.zeros(n, dtype=dt, device=dev)
        self.flat_grad = torch.zeros(n, dtype=dt, device=dev)
        self.exp_avg = torch.zeros(n, dtype=dt, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=dt, device=dev)
        for p, off in zip(self.params, offs):
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view_as(p)
            p.grad = self.flat_grad[off:off + k].view_as(p)
        self.lr, self.betas, self.eps, self.clip_value = lr, betas, eps, clip_value
        self.step_count = 0

    def zero_grad(self):
        self.flat_grad.zero_()           # one memset instead of one per tensor

    def reduce_gradients(self):
        """DDP mean of the gradients: one all-reduce of the flat buffer.  Returns the scale still to apply."""
        if not self.data_parallel or not (dist.is_available() and dist.is_initialized()):
            return 1.0
        world = dist.get_world_size(self.process_group)
        if world == 1:
            return 1.0
        timed = self.reduce_events is not None and self.flat_grad.is_cuda
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.process_group)
        if timed:
            e1.record()
            self.reduce_events.append((e0, e1))
        return 1.0 / world

    def step(self):
        scale = self.reduce_gradients()
        self.step_count += 1
        if not self.flat_param.is_cuda:
            raise RuntimeError("FlatAdam.step: the fused clip+Adam kernel needs CUDA tensors (no CPU fallback)")
        _lib.launch("pu3_clip_adam_f32", self.flat_param, self.flat_param.numel(), self.flat_param.data_ptr(),
                    self.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), float(scale),
                    float(self.clip_value or 0.0), float(self.lr), float(self.betas[0]), float(self.betas[1]),
                    float(self.eps), int(self.step_count))
        _lib.bump_weight_generation()     # raw-pointer update: invalidate caches of weight-derived images
