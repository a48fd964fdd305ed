"""Training path of the drop-in module: autograd bridge to dg_lw_forward / dg_lw_backward and the fused optimizer tail.

Reference call sequence (optimized_train.py:197-236, fp32 branch :220-233):
    optimizer.zero_grad(set_to_none=True); outputs = model(inputs); loss = criterion(outputs, targets)
    loss.backward(); clip_grad_norm_(model.parameters(), 1.0); optimizer.step()
`model(inputs)` lands in `lightweight_forward_train` when autograd is recording; gradients of every parameter come back
through `torch.autograd.Function.backward`, i.e. they land in `param.grad` the normal way (wandb.watch hooks keep working).
`FusedAdamW` is a `torch.optim.Optimizer` with AdamW's constructor that keeps parameters, gradients and both moments in flat
buffers and runs clip + update as two kernels (dg_adamw_step); with torch.distributed initialised it first all-reduces the
flat gradient (one bucket).
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
        grads, off = [], 0
        for p in params:
            n = p.numel()
            grads.append(flat[off:off + n].view_as(p) if p.requires_grad else None)
            off += n
        return (None, None, *grads)


def lightweight_forward_train(module, x):
    if x.requires_grad:
        raise NotImplementedError("gradient w.r.t. the input image is not implemented (the reference never needs it)")
    return _LightweightUNetFn.apply(module, x, *module.parameters())


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW (optimized_train.py:440-446) with flat storage and a fused clip + update (dg_adamw_step).

    max_grad_norm > 0 folds `torch.nn.utils.clip_grad_norm_(params, max_grad_norm)` (optimized_train.py:215,230) into the
    step; leave it 0 if the caller clips itself.  All parameters must be fp32 on one CUDA device."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0):
        params = [p for p in params]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm))
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

    def _attach_grads(self):
        off = 0
        for p in self._ps:
            k = p.numel()
            p.grad = self.flat_g[off:off + k].view_as(p)
            off += k

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
        scale = 1.0
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)   # the one collective of data-parallel training
            scale = 1.0 / dist.get_world_size()
        self._step += 1
        _lib.check(_lib.load().dg_adamw_step(
            self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            self.flat_p.numel(), self._scratch.data_ptr(), float(g["max_grad_norm"]), float(g["lr"]), float(g["betas"][0]),
            float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self._step, scale,
            torch.cuda.current_stream().cuda_stream))
        _lib.bump_generation()   # parameters changed behind autograd's back: invalidate packed-weight caches
        return loss

    def grad_norm(self):
        """Total L2 norm of the (scaled) gradient seen by the last step (the value clip_grad_norm_ returns)."""
        return float(self._scratch.sqrt().item())
