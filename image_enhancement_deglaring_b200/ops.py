"""Thin per-op wrappers over the C-ABI (used by tests and by the training autograd path).

Tensors are torch CUDA tensors; this module only marshals pointers -- the arithmetic is in csrc/.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (DG_BF16, DG_F16, DG_F32, DG_X_CONVT2, DG_X_IMAGE, DG_X_POOL2, DG_X_SAME, DG_X_UP2,  # noqa: F401
                   DgConv3x3Args, DgHeadArgs, DgSrc)

TORCH_DTYPE = {DG_F32: torch.float32, DG_F16: torch.float16, DG_BF16: torch.bfloat16}
DTYPE_OF_TORCH = {v: k for k, v in TORCH_DTYPE.items()}


def _ptr(t):
    return None if t is None else t.data_ptr()


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("image_enhancement_deglaring_b200 runs on CUDA tensors only (no CPU fallback)")


def pack_conv3x3(w):
    """nn.Conv2d weight [Co,Ci,3,3] -> [3,3,Ci,Co] fp32 contiguous."""
    return w.detach().float().permute(2, 3, 1, 0).contiguous()


def pack_convt2x2(w):
    """nn.ConvTranspose2d weight [Ci,Co,2,2] -> [2,2,Ci,Co] fp32 contiguous."""
    return w.detach().float().permute(2, 3, 0, 1).contiguous()


def pack_conv3x3_tc(w_packed, dtype, stream=None):
    """fp32 [3,3,Ci,Co] -> the HMMA kernels' B-tile packing in `dtype` (dg_pack_conv3x3_tc); None if unsupported."""
    lib = _lib.load()
    ci, co = int(w_packed.shape[2]), int(w_packed.shape[3])
    n = C.c_size_t(0)
    if dtype == DG_F32 or lib.dg_tc_conv3x3_bytes(ci, co, C.byref(n)) != 0:
        return None
    out = torch.empty(n.value // 2, dtype=TORCH_DTYPE[dtype], device=w_packed.device)
    _lib.check(lib.dg_pack_conv3x3_tc(w_packed.data_ptr(), out.data_ptr(), ci, co, dtype, _stream(stream)))
    return out


def pack_convt2x2_tc(w_packed, dtype, stream=None):
    """fp32 [2,2,Ci,Co] -> tensor-core packing (dg_pack_convt2x2_tc); None if unsupported."""
    lib = _lib.load()
    ci, co = int(w_packed.shape[2]), int(w_packed.shape[3])
    n = C.c_size_t(0)
    if dtype == DG_F32 or lib.dg_tc_convt2x2_bytes(ci, co, C.byref(n)) != 0:
        return None
    out = torch.empty(n.value // 2, dtype=TORCH_DTYPE[dtype], device=w_packed.device)
    _lib.check(lib.dg_pack_convt2x2_tc(w_packed.data_ptr(), out.data_ptr(), ci, co, dtype, _stream(stream)))
    return out


def pack_dec_composite(ct_w_packed, ct_b, conv_w_packed, dtype, stream=None):
    """ConvTranspose2d(2,2)+bias folded into the taps of the 3x3 conv that consumes cat((up, skip)) (dg_pack_dec_composite);
    ct_w_packed fp32 [2,2,Cl,Cu], ct_b [Cu], conv_w_packed fp32 [3,3,2Cu,Cu].  None where the composite kernel has no coverage."""
    lib = _lib.load()
    cl, cu = int(ct_w_packed.shape[2]), int(ct_w_packed.shape[3])
    n = C.c_size_t(0)
    if dtype == DG_F32 or lib.dg_dec_composite_bytes(cl, cu, C.byref(n)) != 0:
        return None
    out = torch.empty(n.value, dtype=torch.uint8, device=ct_w_packed.device)
    _lib.check(lib.dg_pack_dec_composite(ct_w_packed.data_ptr(), ct_b.data_ptr(), conv_w_packed.data_ptr(), out.data_ptr(), cl, cu,
                                         dtype, _stream(stream)))
    return out


def make_src(raw, channels, xform=DG_X_SAME, stats=None, gamma=None, beta=None, groups=1, silu=True,
             scale=None, ct_w=None, ct_b=None, ct_cout=0, ct_w_tc=None):
    _require_cuda(raw, stats, gamma, beta, scale, ct_w, ct_b, ct_w_tc)
    s = DgSrc()
    s.ct_w_tc = _ptr(ct_w_tc)
    s._keep_tc = ct_w_tc
    # the struct only carries raw pointers: pin the tensors to it so temporaries passed inline by the caller
    # stay alive until the launch (after which the caching allocator's stream ordering protects them)
    s._keep = (raw, stats, gamma, beta, scale, ct_w, ct_b)
    s.raw = _ptr(raw)
    s.stats = _ptr(stats)
    s.gamma = _ptr(gamma)
    s.beta = _ptr(beta)
    s.scale = _ptr(scale)
    s.ct_w = _ptr(ct_w)
    s.ct_b = _ptr(ct_b)
    s.channels = channels
    s.groups = groups
    s.xform = xform
    s.silu = 1 if silu else 0
    s.ct_cout = ct_cout
    return s


def _stream(stream):
    return (torch.cuda.current_stream() if stream is None else stream).cuda_stream


def conv3x3_fused(srcs, weight, cout, N, H, W, dtype, out=None, out_stats=None, act_sum=None, path=0, stream=None,
                  eps=1e-5, weight_tc=None, weight_comp=None):
    """Fused 3x3 conv over the concat of `srcs` (list of DgSrc).  Returns (raw NHWC out, stats [N,cout,2] f64)."""
    lib = _lib.load()
    dev = weight.device
    if out is None:
        out = torch.empty((N, H, W, cout), dtype=TORCH_DTYPE[dtype], device=dev)
    if out_stats is None:
        out_stats = torch.zeros((N, cout, 2), dtype=torch.float64, device=dev)
    a = DgConv3x3Args()
    for i, s in enumerate(srcs):
        a.src[i] = s  # copies the struct; `srcs` (and the tensors pinned to it) outlive the call below
    a.nsrc = len(srcs)
    a.dtype = dtype
    a.N, a.H, a.W, a.cout = N, H, W, cout
    a.weight = _ptr(weight)
    a.weight_tc = _ptr(weight_tc)
    a.weight_comp = _ptr(weight_comp)
    a.out = _ptr(out)
    a.out_stats = _ptr(out_stats)
    a.act_sum = _ptr(act_sum)
    a.eps = eps
    a.path = path
    _lib.check(lib.dg_conv3x3_fused(C.byref(a), _stream(stream)))
    return out, out_stats


def conv3x3_wgrad(srcs, dR, cin_total, cout, N, H, W, dtype, path=0, stream=None, eps=1e-5):
    """Weight gradient of conv3x3_fused's conv: dW [cout, cin_total, 3, 3] fp32 (the nn.Conv2d.weight layout) from the same
    source descriptions and dR = fp32 NHWC [N,H,W,cout] gradient at the raw conv output."""
    lib = _lib.load()
    dW = torch.zeros((cout, cin_total, 3, 3), dtype=torch.float32, device=dR.device)
    a = DgConv3x3Args()
    for i, s in enumerate(srcs):
        a.src[i] = s
    a.nsrc = len(srcs)
    a.dtype = dtype
    a.N, a.H, a.W, a.cout = N, H, W, cout
    a.eps = eps
    a.path = path
    _lib.check(lib.dg_conv3x3_wgrad(C.byref(a), _ptr(dR), _ptr(dW), 1, 9, 9 * cin_total, _stream(stream)))
    return dW


def conv3x3_dgrad(dR, weight_tc_bf16, cin, cout, stream=None):
    """Data gradient of the 3x3 conv on the tensor cores (dg_conv3x3_dgrad): dR fp32 NHWC [N,H,W,cout] -> dX fp32 NHWC [N,H,W,cin];
    weight_tc_bf16 = pack_conv3x3_tc(pack_conv3x3(weight), DG_BF16) of the FORWARD weights."""
    _require_cuda(dR, weight_tc_bf16)
    N, H, W, _ = dR.shape
    dX = torch.empty((N, H, W, cin), dtype=torch.float32, device=dR.device)
    _lib.check(_lib.load().dg_conv3x3_dgrad(_ptr(dR), _ptr(weight_tc_bf16), _ptr(dX), N, H, W, cin, cout, _stream(stream)))
    return dX


def convt2x2_dgrad(dCat, ct_w_tc_bf16, cl, cu, stream=None):
    """Data gradient of ConvTranspose2d(2,2) on the tensor cores (dg_convt2x2_dgrad): the first cu channels of dCat fp32 NHWC
    [N,H,W,stride] -> dLow fp32 NHWC [N,H/2,W/2,cl]; ct_w_tc_bf16 = pack_convt2x2_tc(pack_convt2x2(weight), DG_BF16)."""
    _require_cuda(dCat, ct_w_tc_bf16)
    N, H, W, stride = dCat.shape
    dLow = torch.empty((N, H // 2, W // 2, cl), dtype=torch.float32, device=dCat.device)
    _lib.check(_lib.load().dg_convt2x2_dgrad(_ptr(dCat), stride, _ptr(ct_w_tc_bf16), _ptr(dLow), N, H, W, cl, cu, _stream(stream)))
    return dLow


def head1x1(src, weight, bias, N, H, W, dtype, out=None, target=None, l1_sum=None, stream=None, eps=1e-5):
    lib = _lib.load()
    cout = weight.shape[0]
    if out is None:
        out = torch.empty((N, cout, H, W), dtype=torch.float32, device=weight.device)
    a = DgHeadArgs()
    a.src = src
    a.dtype = dtype
    a.N, a.H, a.W, a.cout = N, H, W, cout
    a.weight = _ptr(weight)
    a.bias = _ptr(bias)
    a.out = _ptr(out)
    a.target = _ptr(target)
    a.l1_sum = _ptr(l1_sum)
    a.eps = eps
    _lib.check(lib.dg_head1x1(C.byref(a), _stream(stream)))
    return out


def convt2x2_fused(src, N, H, W, dtype, out=None, path=0, stream=None, eps=1e-5):
    """Stand-alone tensor-core ConvTranspose2d(2,2)+bias of the activated low-res source -> NHWC [N,H,W,ct_cout]."""
    lib = _lib.load()
    if out is None:
        out = torch.empty((N, H, W, src.ct_cout), dtype=TORCH_DTYPE[dtype], device=src._keep[0].device)
    _lib.check(lib.dg_convt2x2_fused(C.byref(src), dtype, N, H, W, _ptr(out), eps, path, _stream(stream)))
    return out
