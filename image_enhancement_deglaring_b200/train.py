"""Training path of the drop-in module: autograd bridge to dg_lw_forward / dg_lw_backward and the fused optimizer tail.

Reference call sequence (optimized_train.py:197-236, fp32 branch :220-233):
    optimizer.zero_grad(set_to_none=True); outputs = model(inputs); loss = criterion(outputs, targets)
    loss.backward(); clip_grad_norm_(model.parameters(), 1.0); optimizer.step()
`model(inputs)` lands in `lightweight_forward_train` when autograd is recording; gradients of every parameter come back
through `torch.autograd.Function.backward`, i.e. they land in `param.grad` the normal way (wandb.watch hooks keep working).
`FusedAdamW` is a `torch.optim.Optimizer` with AdamW's constructor that keeps parameters, gradients and both moments in flat
buffers and runs clip + update as two kernels (dg_adamw_step).

Data parallel (one process per GPU): the ONE collective of a training step -- the mean all-reduce of the
flat 486,409-float gradient -- is issued at the END OF BACKWARD (`_LightweightUNetFn.backward`), i.e. between `loss.backward()`
and everything the reference does next (optimized_train.py:210-219 / :226-233): `scaler.unscale_` + its inf check,
`clip_grad_norm_`, `optimizer.step()` all see the already-averaged gradient, every rank takes the same skip / clip decision,
and no rank can miss the collective.  `module.ddp_sync = False` turns it off (gradient accumulation over micro-batches).
"""
import ctypes as C

import torch

from . import _lib


class _LightweightUNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        lib = _lib.load()
        x = x.detach().float().contiguous()
        N, _, H, W = x.shape
        pc = module._refresh(train=True)
        ws = torch.empty(module.workspace_bytes(N, H, W), dtype=torch.uint8, device=x.device)  # kept for backward
        y = torch.empty((N, module.out_channels, H, W), dtype=torch.float32, device=x.device)
        _lib.check(lib.dg_lw_forward(C.byref(pc), x.data_ptr(), y.data_ptr(), N, H, W, ws.data_ptr(), ws.numel(),
                                     None, None, torch.cuda.current_stream().cuda_stream))
        ctx.module, ctx.ws, ctx.x = module, ws, x
        return y

    @staticmethod
    def backward(ctx, grad_y):
        lib = _lib.load()
        module, x = ctx.module, ctx.x
        N, _, H, W = x.shape
        grad_y = grad_y.detach().float().contiguous()
        pc = module._refresh(train=True)
        nb = C.c_size_t(0)
        _lib.check(lib.dg_lw_backward_workspace_bytes(C.byref(pc), N, H, W, C.byref(nb)))
        bws = torch.empty(nb.value, dtype=torch.uint8, device=x.device)
        params = list(module.parameters())
        total = sum(p.numel() for p in params)
        cnt = C.c_size_t(0)
        _lib.check(lib.dg_lw_num_params(C.byref(pc), C.byref(cnt)))
        if cnt.value != total:
            raise RuntimeError(f"gradient layout mismatch: library {cnt.value} vs module {total}")
        flat = torch.empty(total, dtype=torch.float32, device=x.device)
        _lib.check(lib.dg_lw_backward(C.byref(pc), x.data_ptr(), grad_y.data_ptr(), N, H, W, ctx.ws.data_ptr(),
                                      ctx.ws.numel(), bws.data_ptr(), bws.numel(), flat.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream))
        ctx.ws = None
        sync_gradients(flat, getattr(module, "ddp_sync", True))
        grads, off = [], 0
        for p in params:
            n = p.numel()
            grads.append(flat[off:off + n].view_as(p) if p.requires_grad else None)
            off += n
        return (None, None, *grads)


def sync_gradients(flat, enabled=True, group=None):
    """Mean all-reduce of the flat gradient over the data-parallel ranks (no-op without an initialised process group).
    NCCL averages inside the collective; backends without AVG (gloo) sum and scale."""
    import torch.distributed as dist
    if not enabled or not (dist.is_available() and dist.is_initialized()):
        return flat
    world = dist.get_world_size(group)
    if world <= 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)
    return flat


def lightweight_forward_train(module, x):
    if x.requires_grad:
        raise NotImplementedError("gradient w.r.t. the input image is not implemented (the reference never needs it)")
    return _LightweightUNetFn.apply(module, x, *module.parameters())


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW (optimized_train.py:440-446) with flat storage and a fused clip + update (dg_adamw_step).

    max_grad_norm > 0 folds `torch.nn.utils.clip_grad_norm_(params, max_grad_norm)` (optimized_train.py:215,230) into the
    step; leave it 0 if the caller clips itself.  All parameters must be fp32 on one CUDA device, in ONE param group.

    Checkpoints: `state_dict()` / `load_state_dict()` use torch.optim.AdamW's own layout (per-parameter `step`, `exp_avg`,
    `exp_avg_sq`), so the `optimizer_state_dict` entry the reference writes into every checkpoint (optimized_train.py:69)
    round-trips with a plain AdamW in either direction.  The data-parallel gradient exchange is NOT here: it happens at the
    end of backward (see the module docstring), before any clipping or inf check."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0):
        params = [p for p in params]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdamW keeps one flat buffer: pass a single param group (per-group hyper-parameters are not supported)")
        ps = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
        if not ps or any(not p.is_cuda or p.dtype != torch.float32 for p in ps):
            raise RuntimeError("FusedAdamW needs fp32 CUDA parameters")
        self._ps = ps
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._scratch = torch.zeros(1, dtype=torch.float64, device=dev)
        self._step = 0
        off = 0
        with torch.no_grad():
            for p in ps:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)     # parameters become views of the flat buffer
                off += k
        self._attach_grads()

    def add_param_group(self, param_group):
        if getattr(self, "param_groups", None):
            raise ValueError("FusedAdamW supports a single param group")
        super().add_param_group(param_group)

    def _attach_grads(self):
        off = 0
        for p in self._ps:
            k = p.numel()
            p.grad = self.flat_g[off:off + k].view_as(p)
            off += k

    def _publish_state(self):
        """Expose the flat moments through `self.state` in torch.optim.AdamW's layout (views, no copies)."""
        off = 0
        for p in self._ps:
            k = p.numel()
            # one `step` tensor PER parameter (torch.optim.AdamW increments each entry in place: a shared tensor would be bumped
            # once per parameter after a reload)
            self.state[p] = {"step": torch.tensor(float(self._step)), "exp_avg": self.exp_avg[off:off + k].view_as(p),
                             "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p)}
            off += k

    def state_dict(self):
        if self._step > 0:
            self._publish_state()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)   # validates the groups / sizes and fills self.state with (copied) tensors
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdamW supports a single param group")
        steps = set()
        off = 0
        with torch.no_grad():
            for p in self._ps:
                k = p.numel()
                st = self.state.get(p)
                if st:
                    self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(int(float(st["step"])))
                else:
                    self.exp_avg[off:off + k].zero_()
                    self.exp_avg_sq[off:off + k].zero_()
                    steps.add(0)
                off += k
        if len(steps) != 1:
            raise ValueError(f"FusedAdamW needs one common step count, checkpoint has {sorted(steps)}")
        self._step = steps.pop()
        if self._step > 0:
            self._publish_state()
        else:
            self.state.clear()

    def zero_grad(self, set_to_none=True):
        # the reference calls zero_grad(set_to_none=True) (optimized_train.py:201); keep the flat aliasing instead
        self.flat_g.zero_()
        self._attach_grads()

    def _gather(self):
        off = 0
        for p in self._ps:
            k = p.numel()
            want = self.flat_g[off:off + k]
            if p.grad is None:
                want.zero_()
            elif p.grad.data_ptr() != want.data_ptr():
                want.copy_(p.grad.reshape(-1))
            off += k

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._gather()
        g = self.param_groups[0]
        self._step += 1
        _lib.check(_lib.load().dg_adamw_step(
            self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            self.flat_p.numel(), self._scratch.data_ptr(), float(g["max_grad_norm"]), float(g["lr"]), float(g["betas"][0]),
            float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self._step, 1.0,
            torch.cuda.current_stream().cuda_stream))
        _lib.bump_generation()   # parameters changed behind autograd's back: invalidate packed-weight caches
        return loss

    def grad_norm(self):
        """Total L2 norm of the gradient seen by the last step (the value clip_grad_norm_ returns)."""
        return float(self._scratch.sqrt().item())
