"""The network applied to ONE large image, exactly, with the image's rows sharded over GPUs
(BASELINE.json configs[2], SURVEY 8e "definition B").

`tiling.infer_tiled` (definition A) runs the reference on independent 512x512 tiles: a different function, because GroupNorm
(src/model.py:94,97) normalises over the whole image and the 3x3 convs / pools / transposed convs (src/model.py:93,96,35-53)
see across what would be tile borders.  Here the function is the reference's `forward` on the full image:

  * rank r owns a band of `Hb` consecutive rows (a multiple of 32) at every level of the UNet (Hb >> level rows there);
  * every raw conv output is held with TWO halo rows of the neighbouring bands above and below (none at the image border,
    where the kernels' own zero padding of the activated tensor applies): one row feeds a 3x3 conv, two rows feed the
    2x2 pool in front of the next level's conv, two rows of `up` come from one halo row of the low-resolution tensor;
  * after each of the 18 convs (a) the band's GroupNorm partial sums -- the producing kernel's epilogue statistics minus
    the contribution of the halo rows it also computed -- are summed over ranks and turned into the per-channel affine
    (a, b) that the consumer kernels take ready-made (`dg_src.coef`), and (b) the two outermost owned rows go to both
    neighbours' halo rows.  Both exchanges ride on ONE all-gather per conv (packet = [C x 2 doubles | 2 top rows | 2 bottom
    rows], 0.26 MB per rank at any level of a 4096-wide image): the band forward is latency-bound (35 separate small
    collectives cost more than the convolutions), and adding the gathered partial sums in rank order gives every rank the
    same bits.

So the data path has exactly the two exchange steps per layer SURVEY 8e names; everything else is the per-op C-ABI
(`dg_conv3x3_fused`, `dg_convt2x2_fused`, `dg_head1x1`) on band-sized tensors.  The arithmetic backend is pluggable only so
that the host logic (halo bookkeeping, statistics, exchange) can be tested on CPU under gloo with a torch-functional
stand-in supplied BY THE TESTS; the product backend is `KernelBackend` and has no fallback.
"""
import ctypes as C
import threading

import torch

from . import _lib
from ._lib import DG_F32, DG_X_CONVT2, DG_X_IMAGE, DG_X_POOL2, DG_X_SAME, DgConv3x3Args, DgHeadArgs, DgSrc
from .ops import TORCH_DTYPE

HALO = 2
_BLOCKS = ("enc1", "enc2", "enc3", "enc4", "bottleneck", "dec4", "dec3", "dec2", "dec1")


def band_rows(H, rank, world):
    """Rows [r0, r1) of an H-row image owned by `rank`: equal bands, each a multiple of 32 rows (two rows at the bottleneck)."""
    if H % world or (H // world) % 32:
        raise RuntimeError(f"{H} rows do not split into {world} bands of a multiple of 32 rows")
    hb = H // world
    return rank * hb, (rank + 1) * hb


def band_with_halo(image, rank, world):
    """The rows of `image` [H, W] rank needs: its band plus HALO rows of each existing neighbour."""
    r0, r1 = band_rows(image.shape[-2], rank, world)
    return image[..., max(r0 - HALO, 0):min(r1 + HALO, image.shape[-2]), :]


# ---- communicators --------------------------------------------------------------------------------------------------
class DistComm:
    """One process per band over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def gather(self, packed):
        """packed: contiguous 1-D uint8 tensor, the same length on every rank -> [world, n] (every rank's packet, rank order)."""
        out = torch.empty((self.world, packed.numel()), dtype=packed.dtype, device=packed.device)
        if self.world == 1:
            out[0].copy_(packed)
        elif self.dist.get_backend(self.group) == "nccl":
            self.dist.all_gather_into_tensor(out, packed, group=self.group)
        else:
            self.dist.all_gather(list(out.unbind(0)), packed, group=self.group)
        return out

    def gather_rows(self, band):
        parts = [torch.empty_like(band) for _ in range(self.world)]
        if self.world > 1:
            self.dist.all_gather(parts, band.contiguous(), group=self.group)
        else:
            parts[0] = band
        return torch.cat(parts, 0)


class LocalBands:
    """`world` bands as threads of ONE process on one device (same stream, so kernels run in issue order and the barriers
    below order the issue): the single-GPU harness of the band algorithm, and a way to bound the per-call workspace."""

    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world

    def comm(self, rank):
        return _LocalComm(self, rank)


class _LocalComm:
    def __init__(self, shared, rank):
        self.s, self.rank, self.world = shared, rank, shared.world

    def gather(self, packed):
        s = self.s
        s.slots[self.rank] = packed
        s.barrier.wait()
        out = torch.stack(list(s.slots), 0)
        s.barrier.wait()
        return out

    def gather_rows(self, band):
        s = self.s
        s.slots[self.rank] = band
        s.barrier.wait()
        out = torch.cat(list(s.slots), 0)
        s.barrier.wait()
        return out


# ---- arithmetic backend: the per-op C-ABI -----------------------------------------------------------------------------
class KernelBackend:
    """Band-sized calls into libdeglare.so.  A source is (kind, tensor, coef): kind in {"image", "same", "pool", "convt"},
    tensor = fp32 [rows, W] image band or raw NHWC [rows, W, C] in the storage type, coef = float32 [C, 2] finished GroupNorm
    affine of that tensor (None for the image)."""

    def __init__(self, net):
        dev = net.output_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("whole-image inference runs on CUDA only (no CPU fallback)")
        self.net, self.device = net, dev
        self.pc = net.c_params()
        self.dtype = TORCH_DTYPE[self.pc.dtype]
        self._dummy = torch.zeros(2048 * 2, dtype=torch.float64, device=dev)   # non-null `stats` where a kernel's routing asks for it

    def alloc(self, rows, W, channels):
        return torch.empty((rows, W, channels), dtype=self.dtype, device=self.device)

    def _src(self, kind, t, coef, producer, u=None, keep=None):
        s = DgSrc()
        s.raw = t.data_ptr()
        if kind == "image":
            s.channels, s.groups, s.xform = self.net.in_channels, 1, DG_X_IMAGE
            return s
        b, j = divmod(producer, 2)
        s.stats = self._dummy.data_ptr()
        s.gamma, s.beta = self.pc.gn_w[b][j], self.pc.gn_b[b][j]
        s.coef = coef.data_ptr()
        s.channels, s.groups, s.silu = t.shape[-1], self.net._block_groups[b], 1
        s.xform = {"same": DG_X_SAME, "pool": DG_X_POOL2, "convt": DG_X_CONVT2}[kind]
        if kind == "convt":
            s.ct_w, s.ct_b, s.ct_w_tc = self.pc.up_w[u], self.pc.up_b[u], self.pc.up_w_tc[u]
            s.ct_cout = t.shape[-1] // 2
        return s

    def conv(self, i, srcs, out):
        """Conv `i` (0..17, parameters() order) of the activated sources over the rows of `out` [H, W, C]; returns the
        kernel's epilogue statistics [C, 2] (sum, sum of squares over ALL rows of `out`) as float64."""
        lib = _lib.load()
        H, W, cout = out.shape
        b, j = divmod(i, 2)
        stream = torch.cuda.current_stream().cuda_stream
        a = DgConv3x3Args()
        keep = []
        for k, (kind, t, coef, producer) in enumerate(srcs):
            u = b - 5 if kind == "convt" else None
            s = self._src(kind, t, coef, producer, u)
            if kind == "convt" and self.pc.dtype != DG_F32 and (self.pc.path & 3) != 1 and s.ct_cout >= 32:
                # same split as the native orchestrator (csrc/api.cu:lw_up_materialised): a stand-alone tensor-core ConvTranspose
                # feeds the concat conv as an identity source; the concat itself is never stored
                up = torch.empty((H, W, s.ct_cout), dtype=self.dtype, device=self.device)
                _lib.check(lib.dg_convt2x2_fused(C.byref(s), self.pc.dtype, 1, H, W, up.data_ptr(), 1e-5, self.pc.path, stream))
                keep.append(up)
                s = DgSrc()
                s.raw, s.channels, s.groups, s.xform = up.data_ptr(), up.shape[-1], 1, DG_X_SAME
            a.src[k] = s
        a.nsrc, a.dtype = len(srcs), self.pc.dtype
        a.N, a.H, a.W, a.cout = 1, H, W, cout
        a.weight, a.weight_tc = self.pc.conv_w[b][j], self.pc.conv_w_tc[b][j]
        stats = torch.zeros((cout, 2), dtype=torch.float64, device=self.device)
        a.out, a.out_stats = out.data_ptr(), stats.data_ptr()
        a.eps, a.path = 1e-5, self.pc.path
        _lib.check(lib.dg_conv3x3_fused(C.byref(a), stream))
        return stats

    def band_stats(self, stats, t, c0, own0, own1, c1, out):
        """Partial sums of the rows the band owns, written to `out` (float64 [C, 2] view of the packet): `stats` (the conv's epilogue
        sums over the computed rows [c0, c1) of t) minus the computed halo rows [c0, own0) and [own1, c1) -- one launch."""
        if t.shape[-1] > 1024:
            return False     # the driver's tensor-op path
        _lib.check(_lib.load().dg_band_stats(stats.data_ptr(), t.data_ptr(), self.pc.dtype, t.shape[1], t.shape[2], c0, own0, own1, c1,
                                             out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return True

    def gn_affine(self, parts, stride_bytes, c, i, plane):
        """Every rank's partial sums of conv `i`'s output (`parts`: the gathered packets, one every `stride_bytes`, the [C, 2]
        doubles at their start) -> summed in rank order -> its GroupNorm's finished affine [C, 2] -- one launch (dg_gn_affine)."""
        b, j = divmod(i, 2)
        coef = torch.empty((c, 2), dtype=torch.float32, device=self.device)
        _lib.check(_lib.load().dg_gn_affine(parts.data_ptr(), parts.shape[0], stride_bytes // 8, self.pc.gn_w[b][j], self.pc.gn_b[b][j], c,
                                            self.net._block_groups[b], float(plane), 1e-5, coef.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
        return coef

    def head(self, t, coef, out):
        """GroupNorm + SiLU + 1x1 conv + bias of raw [H, W, C] -> fp32 [out_channels, H, W]."""
        h = DgHeadArgs()
        h.src = self._src("same", t, coef, 17)
        h.dtype = self.pc.dtype
        h.N, h.H, h.W, h.cout = 1, t.shape[0], t.shape[1], self.net.out_channels
        h.weight, h.bias, h.out, h.eps = self.pc.head_w, self.pc.head_b, out.data_ptr(), 1e-5
        _lib.check(_lib.load().dg_head1x1(C.byref(h), torch.cuda.current_stream().cuda_stream))
        return out


# ---- the band algorithm -------------------------------------------------------------------------------------------------
def _row_stats(t):
    """[rows, W, C] -> float64 [C, 2] (sum, sum of squares) of the stored values."""
    if t.shape[0] == 0:
        return torch.zeros((t.shape[-1], 2), dtype=torch.float64, device=t.device)
    d = t.to(torch.float64)
    return torch.stack((d.sum((0, 1)), (d * d).sum((0, 1))), -1)


def _gn_coef(stats, gamma, beta, groups, count, eps=1e-5):
    """Whole-image statistics [C, 2] -> GroupNorm affine [C, 2] float32: y = x * a + b, a = rstd * gamma, b = beta - mean * a
    (nn.GroupNorm, biased variance, src/model.py:94,97)."""
    C_ = stats.shape[0]
    g = stats.view(groups, C_ // groups, 2).sum(1) / float(count * (C_ // groups))
    mean, var = g[:, 0], (g[:, 1] - g[:, 0] * g[:, 0]).clamp_min(0.0)
    rstd = torch.rsqrt(var + eps)
    mean = mean.repeat_interleave(C_ // groups)
    a = rstd.repeat_interleave(C_ // groups) * gamma.to(torch.float64)
    return torch.stack((a, beta.to(torch.float64) - mean * a), -1).to(torch.float32).contiguous()


def forward_band(net, x_band, comm, H_total, backend=None):
    """One rank's part of `net` applied to a whole [H_total, W] image.

    x_band: float32 [rows, W] -- `band_with_halo(image, comm.rank, comm.world)` on the backend's device.
    Returns float32 [out_channels, Hb, W]: the rank's band of the output (no halo rows).  Every rank must call it."""
    be = backend or KernelBackend(net)
    rank, world = comm.rank, comm.world
    r0, r1 = band_rows(H_total, rank, world)
    Hb, W = r1 - r0, x_band.shape[-1]
    ht, hb = (HALO if rank > 0 else 0), (HALO if rank < world - 1 else 0)
    if W % 16 or tuple(x_band.shape) != (ht + Hb + hb, W):
        raise RuntimeError(f"band of rank {rank}: expected [{ht + Hb + hb}, W = 16k], got {tuple(x_band.shape)}")
    f = [net.features_start << l for l in range(5)]
    x_band = x_band.contiguous()
    T, coef = [None] * 18, [None] * 18

    def finish(i, lvl, stats, c0, c1, exchange=True):
        """Rows [c0, c1) of T[i] were just computed.  ONE all-gather carries both exchanges of SURVEY 8e: every rank's packet is
        [its partial sums (C x 2 doubles) | its two top-most owned rows | its two bottom-most owned rows]; afterwards the sums are
        added in rank order (the same bits on every rank) -> coef[i], and the neighbours' rows land in the halo rows."""
        t, hb_l = T[i], Hb >> lvl
        c = t.shape[-1]
        exchange = exchange and world > 1
        sb = 16 * c                                              # statistics bytes
        rb = HALO * t.shape[1] * c * t.element_size() if exchange else 0   # bytes of two rows
        packet = torch.empty(sb + 2 * rb, dtype=torch.uint8, device=t.device)
        own = packet[:sb].view(torch.float64).view(c, 2)
        if not (hasattr(be, "band_stats") and be.band_stats(stats, t, c0, ht, ht + hb_l, c1, own)):
            own.copy_(stats - _row_stats(t[c0:ht]) - _row_stats(t[ht + hb_l:c1]))   # backends without the fused step
        if exchange:
            packet[sb:sb + rb].view(t.dtype).copy_(t[ht:ht + HALO].reshape(-1))
            packet[sb + rb:].view(t.dtype).copy_(t[ht + hb_l - HALO:ht + hb_l].reshape(-1))
        allp = comm.gather(packet)
        b, j = divmod(i, 2)
        plane = (H_total >> lvl) * (W >> lvl)
        if hasattr(be, "gn_affine"):
            coef[i] = be.gn_affine(allp, allp.shape[1], c, i, plane)
        else:
            total = allp[:, :sb].contiguous().view(torch.float64).view(world, c, 2).sum(0)
            gn = getattr(net, _BLOCKS[b])[1 if j == 0 else 4]
            coef[i] = _gn_coef(total, gn.weight.detach(), gn.bias.detach(), net._block_groups[b], plane)
        if exchange:
            if rank > 0:                 # the upper neighbour's bottom rows
                t[0:ht].reshape(-1).copy_(allp[rank - 1, sb + rb:].view(t.dtype))
            if rank < world - 1:         # the lower neighbour's top rows
                t[ht + hb_l:].reshape(-1).copy_(allp[rank + 1, sb:sb + rb].view(t.dtype))

    for b in range(9):
        lvl = b if b < 5 else 8 - b
        rows, w_l = ht + (Hb >> lvl) + hb, W >> lvl
        i = 2 * b
        T[i] = be.alloc(rows, w_l, f[lvl])
        if b == 0:
            st = be.conv(i, [("image", x_band, None, None)], T[i])
            c0, c1 = 0, rows
        elif b < 5:     # AvgPool2d(2,2) of the level above: (ht + 2 Hb' + hb) / 2 rows, written below the outer halo row
            c0, c1 = ht // 2, rows - hb // 2
            st = be.conv(i, [("pool", T[i - 1], coef[i - 1], i - 1)], T[i][c0:c1])
        else:           # ConvTranspose2d(2,2) of the level below without its outer halo row, cat (up, skip)
            low, skip = T[i - 1], T[2 * lvl + 1]
            low = low[ht // 2:low.shape[0] - hb // 2]
            st = be.conv(i, [("convt", low, coef[i - 1], i - 1), ("same", skip, coef[2 * lvl + 1], 2 * lvl + 1)], T[i])
            c0, c1 = 0, rows
        finish(i, lvl, st, c0, c1)
        T[i + 1] = be.alloc(rows, w_l, f[lvl])
        st = be.conv(i + 1, [("same", T[i], coef[i], i)], T[i + 1])
        finish(i + 1, lvl, st, 0, rows, exchange=(i + 1 != 17))
        if b >= 5:
            T[i - 1] = T[2 * lvl + 1] = None   # the low-resolution tensor and the skip are dead: let the allocator reuse them
    y = torch.empty((net.out_channels, ht + Hb + hb, W), dtype=torch.float32, device=x_band.device)
    be.head(T[17], coef[17], y)
    return y[:, ht:ht + Hb]


class BandGraph:
    """forward_band captured ONCE in a CUDA graph (kernels, the small statistics / affine ops and the NCCL all-reduces and
    send/recv pairs alike) and replayed per image: a band's forward is ~60 kernels of a few microseconds each plus ~35 tiny
    collectives, i.e. host-launch bound when issued from Python.  Fixed shape; weights are read at capture time."""

    def __init__(self, net, comm, H_total, W, backend=None, warmup=2):
        be = backend or KernelBackend(net)
        r0, r1 = band_rows(H_total, comm.rank, comm.world)
        rows = (r1 - r0) + (HALO if comm.rank > 0 else 0) + (HALO if comm.rank < comm.world - 1 else 0)
        self.x = torch.zeros((rows, W), dtype=torch.float32, device=be.device)
        side = torch.cuda.Stream(device=be.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():      # lazy initialisations (kernel attributes, NCCL channels) outside capture
            for _ in range(warmup):
                forward_band(net, self.x, comm, H_total, be)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.y = forward_band(net, self.x, comm, H_total, be)

    def __call__(self, x_band):
        self.x.copy_(x_band)
        self.graph.replay()
        return self.y


def infer_whole_sharded(net, image, group=None, gather=True, backend=None):
    """`net` on the whole `image` [H, W] (float32, H a multiple of 32 x world, W of 16), rows sharded over the ranks of `group`
    (torch.distributed, one process per GPU).  Every rank passes the same image (or at least its own `band_with_halo` rows of
    it in place).  gather=True: every rank returns [out_channels, H, W]; gather=False: its band [out_channels, H / world, W]."""
    comm = DistComm(group)
    dev = (backend.device if backend is not None else net.output_conv.weight.device)
    band = band_with_halo(image, comm.rank, comm.world).to(dev, torch.float32)
    y = forward_band(net, band, comm, image.shape[-2], backend)
    if not gather:
        return y
    return comm.gather_rows(y.permute(1, 0, 2).contiguous()).permute(1, 0, 2).contiguous()


def infer_whole_local(net, image, bands, backend=None):
    """The same algorithm with `bands` bands run as threads of this process on the module's device (see LocalBands)."""
    shared = LocalBands(bands)
    dev = (backend.device if backend is not None else net.output_conv.weight.device)
    be = backend or KernelBackend(net)
    outs, errs = [None] * bands, []

    def work(r):
        try:
            if dev.type == "cuda":
                torch.cuda.set_device(dev)
            with torch.no_grad():
                band = band_with_halo(image, r, bands).to(dev, torch.float32)
                outs[r] = forward_band(net, band, shared.comm(r), image.shape[-2], be)
        except BaseException as e:   # noqa: BLE001 -- a failed band must not leave the others waiting on the barrier
            errs.append(e)
            shared.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(bands)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        raise next((e for e in errs if not isinstance(e, threading.BrokenBarrierError)), errs[0])
    return torch.cat(outs, 1)
