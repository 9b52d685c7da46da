"""Host-side plumbing between PyTorch autograd and the C ABI (libdp_b200.so).

Everything numerical happens in the CUDA kernels; this file owns
  * the internal activation format: a plain torch tensor (B,T,H,W,Cp), bf16 (product path) or
    fp32 (validation mode), channel-padded with zeros to a multiple of 16, tagged with `_dp_c`
    (logical channel count);
  * one "layer" = Conv3d -> BatchNorm3d -> LeakyReLU [-> + residual -> LeakyReLU]
    (reference: Conv3dBlock, /root/reference/src/models/R2Plus1D.py:25-58; residual tail :181-187);
  * torch.autograd.Function wrappers at layer and residual-block granularity.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L

_STATE = {"dtype": torch.bfloat16, "impl": L.IMPL_AUTO, "fuse_bn_reduce": True, "fuse_eval": True,
          # experiment, off: BatchNorm finalisation by the last CTA of the kernel that produces the partial sums
          # (dp_bn_fin) instead of stand-alone dp_bn_finalize / dp_bn_bwd_finalize launches.  64 launches less per step,
          # but measured SLOWER (18.98 vs 18.35 ms/step): one CTA walks the 148-592 partial rows of all channel groups
          # while 147 SMs idle (+15 us per conv, +26 us per reduction), the stand-alone kernels spread the groups over
          # CTAs and cost ~5 us including the launch gap inside the CUDA graph.  DP_FUSE_FIN=1 enables it.
          "fuse_fin": os.environ.get("DP_FUSE_FIN", "0") == "1",
          # experiment, off: weight gradients on a second stream.  Measured 18.97 vs 19.02 ms/step: the wgrad CTAs cannot
          # co-reside with the 384-thread gather CTAs and the BatchNorm passes already fill the machine.
          "wgrad_stream": os.environ.get("DP_WGRAD_STREAM", "0") == "1"}
_GRAD = [True]  # autograd mode of the CALLER of the layer functions (inside Function.forward it is always off)
_SIDE = {}
_PENDING = []   # (done event, tensors kept alive until the main stream has waited for it)


def _side_stream(device) -> "torch.cuda.Stream":
    s = _SIDE.get(device.index)
    if s is None:
        s = torch.cuda.Stream(device=device)
        _SIDE[device.index] = s
    return s


def join_side_stream() -> None:
    """Make the current stream wait for every weight gradient still running on the side stream."""
    if not _PENDING:
        return
    cur = torch.cuda.current_stream()
    for done, _keep in _PENDING:
        cur.wait_event(done)
    _PENDING.clear()


def set_compute_mode(mode: str) -> None:
    """'bf16' = product path (bf16 storage, fp32 accumulate); 'fp32' = validation mode."""
    if mode not in ("bf16", "fp32"):
        raise ValueError("mode must be 'bf16' or 'fp32'")
    _STATE["dtype"] = torch.bfloat16 if mode == "bf16" else torch.float32


def get_compute_mode() -> str:
    return "bf16" if _STATE["dtype"] == torch.bfloat16 else "fp32"


def set_conv_impl(impl: str) -> None:
    """'auto' | 'simt' | 'tc' -- which kernel family the conv entry points use."""
    _STATE["impl"] = {"auto": L.IMPL_AUTO, "simt": L.IMPL_SIMT, "tc": L.IMPL_TC}[impl]


@contextlib.contextmanager
def compute_mode(mode: str, impl: Optional[str] = None):
    old = dict(_STATE)
    set_compute_mode(mode)
    if impl is not None:
        set_conv_impl(impl)
    try:
        yield
    finally:
        _STATE.update(old)


class KernelProfiler:
    """CUDA-event timing of each C-ABI kernel call on the launching stream (bench.py's roofline leg).
    Records (family, algorithmic flops, algorithmic bytes, start event, end event)."""

    def __init__(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for fam, flops, nbytes, e0, e1 in self.records:
            d = out.setdefault(fam, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
            d["launches"] += 1
        return out


PROFILER: Optional[KernelProfiler] = None


def _pb():
    if PROFILER is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _pe(e0, family, flops=0.0, nbytes=0.0):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    PROFILER.records.append((family, flops, nbytes, e0, e1))


def ceil16(c: int) -> int:
    return (c + 15) // 16 * 16


def _code(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return L.DP_BF16
    if t.dtype == torch.float32:
        return L.DP_F32
    raise L.DpError(f"unsupported activation dtype {t.dtype}")


def tag(t: torch.Tensor, c: int) -> torch.Tensor:
    t._dp_c = c
    return t


def is_internal(t) -> bool:
    return isinstance(t, torch.Tensor) and getattr(t, "_dp_c", None) is not None


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


# ----------------------------------------------------------------------------------------------
# geometry cache
# ----------------------------------------------------------------------------------------------
class ConvGeom:
    __slots__ = ("desc", "rows_out", "rows_in", "out_shape", "in_shape", "taps", "ws_bytes", "key", "flops", "esize",
                 "fam", "stem", "cls")

    def __init__(self, C_in, K, kernel, stride, padding, B, T, H, W, dtype_code):
        kt, kh, kw = kernel
        st, sh, sw = stride
        pt, ph, pw = padding
        To = (T + 2 * pt - kt) // st + 1
        Ho = (H + 2 * ph - kh) // sh + 1
        Wo = (W + 2 * pw - kw) // sw + 1
        if min(To, Ho, Wo) <= 0:
            raise L.DpError(f"conv geometry gives empty output: in {(T, H, W)} k {kernel} s {stride} p {padding}")
        self.desc = L.ConvDesc(B, T, H, W, C_in, ceil16(C_in), To, Ho, Wo, K, ceil16(K), kt, kh, kw, st, sh, sw,
                               pt, ph, pw, dtype_code)
        self.rows_out = B * To * Ho * Wo
        self.rows_in = B * T * H * W
        self.out_shape = (B, To, Ho, Wo, ceil16(K))
        self.in_shape = (B, T, H, W, ceil16(C_in))
        self.taps = kt * kh * kw
        self.ws_bytes = None
        self.flops = 2.0 * self.rows_out * K * C_in * self.taps      # algorithmic (unpadded) FLOPs of one pass
        self.esize = 2 if dtype_code == L.DP_BF16 else 4
        self.fam = None
        self.stem = False   # True: x is the packed-rows clip of the stem fast path (csrc/stem.cu)
        self.cls = None     # (impl, elements of the class-packed dgrad weights; 0 = per-class launches), see dgrad_classes()

    def dgrad_classes(self, impl) -> int:
        """Elements of the class-packed weights when the strided data gradient of this geometry runs as ONE launch
        (dp_conv_dgrad_classes), else 0."""
        if self.cls is None or self.cls[0] != impl:
            n = 0
            d = self.desc
            if not self.stem and d.dtype == L.DP_BF16 and (d.st, d.sh, d.sw) != (1, 1, 1):
                n = int(L.load().dp_dgrad_classes_weight_elems(C.byref(d), impl))
            self.cls = (impl, n)
        return self.cls[1]

    def families(self, impl):
        """Which kernel family serves fwd / dgrad / wgrad for this geometry (profiling labels)."""
        if self.stem:
            return (impl, "tc_gather_gemm", "tc_gather_gemm", "tc_wgrad")
        if self.fam is None or self.fam[0] != impl:
            lib = L.load()
            tc = [bool(impl != L.IMPL_SIMT and lib.dp_conv_supported(C.byref(self.desc), op, L.IMPL_TC)) for op in range(3)]
            self.fam = (impl, "tc_gather_gemm" if tc[0] else "simt_gather_gemm",
                        "tc_gather_gemm" if tc[1] else "simt_gather_gemm", "tc_wgrad" if tc[2] else "simt_wgrad")
        return self.fam


_GEOMS = {}
_FUSED = {}


def _fused_bnstats(geom: "ConvGeom") -> bool:
    """True when the data gradient of this geometry produces the producer's BatchNorm-backward sums in its epilogue
    (otherwise the stand-alone reduction runs in the producer's own backward, as without the fusion)."""
    key = (id(geom), _STATE["impl"])
    r = _FUSED.get(key)
    if r is None:
        r = bool(L.load().dp_conv_supported(C.byref(geom.desc), 3, _STATE["impl"]))
        _FUSED[key] = r
    return r


def conv_geom(C_in, K, kernel, stride, padding, x: torch.Tensor) -> ConvGeom:
    B, T, H, W, Cp = x.shape
    key = (C_in, K, kernel, stride, padding, B, T, H, W, x.dtype)
    g = _GEOMS.get(key)
    if g is None:
        if Cp != ceil16(C_in):
            raise L.DpError(f"activation has {Cp} padded channels, conv expects {ceil16(C_in)} ({C_in} logical)")
        g = ConvGeom(C_in, K, kernel, stride, padding, B, T, H, W, _code(x))
        _GEOMS[key] = g
    return g


_WS = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.index, "ws")
    t = _WS.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = t
    return t


def _zeros_ws(name: str, nbytes: int, device) -> torch.Tensor:
    """Persistent zero-initialised workspace (self-resetting kernels rely on it)."""
    key = (device.index, name)
    t = _WS.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _WS[key] = t
    return t


_TICKETS = {}
_N_TICKETS = 4096


def _ticket(device) -> int:
    """Address of a zero 32-bit word for a kernel with a "last CTA done" tail (dp_bn_fin.ticket).  The kernel re-arms
    the word before it ends; the pool is handed out round-robin, far deeper than the launches one stream keeps in
    flight, and a CUDA graph keeps the words it captured."""
    key = device.index
    pool = _TICKETS.get(key)
    if pool is None:
        pool = [torch.zeros(_N_TICKETS, dtype=torch.int32, device=device), 0]
        _TICKETS[key] = pool
    i = pool[1]
    pool[1] = (i + 1) % _N_TICKETS
    return pool[0].data_ptr() + 4 * i


def _bn_fin_fwd(d, rows, gamma, beta, cfg, running_mean, running_var, stats, device) -> "L.BnFin":
    return L.BnFin(kind=1, C=d.K, Cp=d.Kp, coef_zero=0, count=float(rows), gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                   eps=cfg.eps, momentum=cfg.momentum, running_mean=_p(running_mean), running_var=_p(running_var),
                   mean=stats[0].data_ptr(), rstd=stats[1].data_ptr(), scale=stats[2].data_ptr(),
                   shift=stats[3].data_ptr(), dgamma=None, dbeta=None, coef=None, ticket=_ticket(device))


def _bn_fin_bwd(K, Kp, rows, stats, training, device):
    """(fin, dgamma, dbeta, coef) for the backward sums of a layer with K (Kp padded) output channels."""
    dgamma = torch.empty(K, dtype=torch.float32, device=device)
    dbeta = torch.empty(K, dtype=torch.float32, device=device)
    coef = torch.empty((2, Kp), dtype=torch.float32, device=device)
    fin = L.BnFin(kind=2, C=K, Cp=Kp, coef_zero=0 if training else 1, count=float(rows), gamma=None, beta=None, eps=0.0,
                  momentum=0.0, running_mean=None, running_var=None, mean=stats[0].data_ptr(), rstd=stats[1].data_ptr(),
                  scale=None, shift=None, dgamma=dgamma.data_ptr(), dbeta=dbeta.data_ptr(), coef=coef.data_ptr(),
                  ticket=_ticket(device))
    return fin, dgamma, dbeta, coef


# ----------------------------------------------------------------------------------------------
# layout
# ----------------------------------------------------------------------------------------------
def _to_internal_raw(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    L.require_device()
    if x.dim() != 5 or x.dtype != torch.float32 or not x.is_cuda:
        raise L.DpError(f"expected a CUDA float32 NCDHW tensor, got {tuple(x.shape)} {x.dtype} on {x.device}")
    x = x.contiguous()
    B, Cc, T, H, W = x.shape
    Cp = ceil16(Cc)
    out = torch.empty((B, T, H, W, Cp), dtype=dtype, device=x.device)
    code = L.DP_BF16 if dtype == torch.bfloat16 else L.DP_F32
    L.check(L.load().dp_ncdhw_f32_to_ndhwc(x.data_ptr(), out.data_ptr(), B, Cc, Cp, T, H, W, code, L.stream_ptr()),
            "dp_ncdhw_f32_to_ndhwc")
    return out


def _to_ncdhw_raw(x: torch.Tensor, Cc: int) -> torch.Tensor:
    x = x.contiguous()
    B, T, H, W, Cp = x.shape
    out = torch.empty((B, Cc, T, H, W), dtype=torch.float32, device=x.device)
    L.check(L.load().dp_ndhwc_to_ncdhw_f32(x.data_ptr(), out.data_ptr(), B, Cc, Cp, T, H, W, _code(x), L.stream_ptr()),
            "dp_ndhwc_to_ncdhw_f32")
    return out


class _ToInternal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.c = x.shape[1]
        return _to_internal_raw(x, dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        return _to_ncdhw_raw(g, ctx.c), None


class _ToNCDHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, c):
        ctx.dtype = x.dtype
        return _to_ncdhw_raw(x, c)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        return _to_internal_raw(g, ctx.dtype), None


def to_internal(x: torch.Tensor) -> torch.Tensor:
    """NCDHW fp32 (B,C,T,H,W) -> tagged internal NDHWC tensor."""
    c = x.shape[1]
    return tag(_ToInternal.apply(x, _STATE["dtype"]), c)


def to_ncdhw(x: torch.Tensor) -> torch.Tensor:
    return _ToNCDHW.apply(x, x._dp_c)


def frames_u8_to_internal(frames: torch.Tensor, mean_bgr: Sequence[float]) -> torch.Tensor:
    """(B,T,H,W,3) uint8 BGR frames -> internal tensor, minus per-channel mean
    (reference: DatasetForVideo.normalize/to_tensor, /root/reference/src/dataset.py:104-110,201-205)."""
    L.require_device()
    if frames.dtype != torch.uint8 or frames.dim() != 5 or frames.shape[-1] != 3 or not frames.is_cuda:
        raise L.DpError("expected CUDA uint8 frames of shape (B,T,H,W,3)")
    frames = frames.contiguous()
    B, T, H, W, _ = frames.shape
    dtype = _STATE["dtype"]
    out = torch.empty((B, T, H, W, 16), dtype=dtype, device=frames.device)
    m = (C.c_float * 3)(*[float(v) for v in mean_bgr])
    code = L.DP_BF16 if dtype == torch.bfloat16 else L.DP_F32
    L.check(L.load().dp_u8_frames_to_ndhwc(frames.data_ptr(), out.data_ptr(), m, B, T, H, W, 16, code, L.stream_ptr()),
            "dp_u8_frames_to_ndhwc")
    return tag(out, 3)


# ----------------------------------------------------------------------------------------------
# one layer: conv -> BN -> LeakyReLU [-> +residual -> LeakyReLU]
# ----------------------------------------------------------------------------------------------
class LayerCfg:
    """Static description of one Conv3dBlock (reference R2Plus1D.py:25-58)."""
    __slots__ = ("C", "K", "kernel", "stride", "padding", "slope", "eps", "momentum")

    def __init__(self, C_in, K, kernel, stride, padding, slope, eps=1e-5, momentum=0.1):
        self.C, self.K = int(C_in), int(K)
        self.kernel, self.stride, self.padding = tuple(kernel), tuple(stride), tuple(padding)
        self.slope, self.eps, self.momentum = float(slope), float(eps), float(momentum)


class PackedWeights:
    """Private packed copies of one fp32 master weight, refreshed when the master changes."""
    __slots__ = ("version", "ptr", "dtype", "wf", "wd")

    def __init__(self):
        self.version = -1
        self.ptr = 0
        self.dtype = None
        self.wf = None
        self.wd = None


_WEIGHT_EPOCH = [0]


def bump_weight_epoch() -> None:
    """Called by code that updates master weights through raw pointers (the fused optimiser kernel), which
    autograd's version counter cannot see: every packed copy is rebuilt at its next use."""
    _WEIGHT_EPOCH[0] += 1


def pack_weights(weight: torch.Tensor, geom: ConvGeom, dtype: torch.dtype, cache: Optional[PackedWeights]):
    version = (weight._version, _WEIGHT_EPOCH[0])
    if cache is not None and cache.version == version and cache.ptr == weight.data_ptr() and cache.dtype == dtype:
        return cache.wf, cache.wd
    d = geom.desc
    if weight.dtype != torch.float32 or not weight.is_contiguous():
        raise L.DpError("conv master weights must be contiguous float32")
    if geom.stem:
        wf = torch.empty(int(L.load().dp_stem_weight_elems(C.byref(d))), dtype=torch.bfloat16, device=weight.device)
        wd = wf   # no data gradient on the stem path
        L.check(L.load().dp_stem_pack_weights(C.byref(d), weight.data_ptr(), wf.data_ptr(), L.stream_ptr()),
                "dp_stem_pack_weights")
    else:
        wf = torch.empty((d.Kp, geom.taps, d.Cp), dtype=dtype, device=weight.device)
        ncls = geom.dgrad_classes(_STATE["impl"]) if dtype == torch.bfloat16 else 0
        if ncls:
            # strided layer: the data gradient takes class-packed weights (all stride-parity classes in one launch)
            wd = torch.empty(ncls, dtype=dtype, device=weight.device)
            L.check(L.load().dp_pack_weights(C.byref(d), weight.data_ptr(), wf.data_ptr(), None, L.stream_ptr()), "dp_pack_weights")
            L.check(L.load().dp_pack_weights_dgrad_classes(C.byref(d), weight.data_ptr(), wd.data_ptr(), L.stream_ptr()),
                    "dp_pack_weights_dgrad_classes")
        else:
            wd = torch.empty((d.Cp, geom.taps, d.Kp), dtype=dtype, device=weight.device)
            L.check(L.load().dp_pack_weights(C.byref(d), weight.data_ptr(), wf.data_ptr(), wd.data_ptr(), L.stream_ptr()),
                    "dp_pack_weights")
    if cache is not None:
        cache.version, cache.ptr, cache.dtype, cache.wf, cache.wd = version, weight.data_ptr(), dtype, wf, wd
    return wf, wd


def _conv_fwd_call(lib, geom, x, wf, y, part, nparts, impl, st):
    d = geom.desc
    if geom.stem:
        return lib.dp_stem_conv_fwd(C.byref(d), x.data_ptr(), wf.data_ptr(), y.data_ptr(), part, nparts, st)
    return lib.dp_conv_fwd(C.byref(d), x.data_ptr(), wf.data_ptr(), y.data_ptr(), part, nparts, impl, st)


def layer_forward(x, weight, gamma, beta, running_mean, running_var, cfg: LayerCfg, training: bool,
                  residual=None, slope_res: float = 1.0, cache: Optional[PackedWeights] = None, geom=None,
                  fuse_eval: bool = False, x_view=None, x_ptr=None):
    """Returns (z, saved) with saved = (x, y, out_or_None, stats[4,Kp], w_dgrad, geom).
    `geom` is given only by the stem fast path, where x is the packed-rows clip.
    `fuse_eval` (eval mode, no gradient needed): one kernel per layer (dp_conv_fwd_bnact), saved y is None.
    `x_view` / `x_ptr` (fuse_eval only): element strides (w,h,t,b) and base address of an input VIEW inside `x`
    (sliding windows over a per-frame cache); `geom` must then describe the view."""
    lib = L.load()
    st = L.stream_ptr()
    if geom is None:
        geom = conv_geom(cfg.C, cfg.K, cfg.kernel, cfg.stride, cfg.padding, x)
    d = geom.desc
    act_dtype = torch.bfloat16 if geom.stem else x.dtype
    wf, wd = pack_weights(weight, geom, act_dtype, cache)
    dev = x.device
    y = torch.empty(geom.out_shape, dtype=act_dtype, device=dev)
    stats = torch.empty((4, d.Kp), dtype=torch.float32, device=dev)  # mean, rstd, scale, shift
    impl = _STATE["impl"]
    if training:
        part = torch.empty((L.DP_MAX_PARTS, 2, d.Kp), dtype=torch.float32, device=dev)
        nparts = C.c_int(0)
        t0 = _pb()
        if _STATE["fuse_fin"]:
            # BatchNorm finalisation (mean / rstd / scale / shift / running statistics) by the conv kernel's last CTA
            fin = _bn_fin_fwd(d, geom.rows_out, gamma, beta, cfg, running_mean, running_var, stats, dev)
            if geom.stem:
                L.check(lib.dp_stem_conv_fwd_fin(C.byref(d), x.data_ptr(), wf.data_ptr(), y.data_ptr(), part.data_ptr(),
                                                 C.byref(fin), st), "dp_stem_conv_fwd_fin")
            else:
                L.check(lib.dp_conv_fwd_fin(C.byref(d), x.data_ptr(), wf.data_ptr(), y.data_ptr(), part.data_ptr(),
                                            C.byref(fin), impl, st), "dp_conv_fwd_fin")
        else:
            L.check(_conv_fwd_call(lib, geom, x, wf, y, part.data_ptr(), C.byref(nparts), impl, st), "dp_conv_fwd")
        if t0 is not None:
            _pe(t0, geom.families(impl)[1], geom.flops, geom.esize * (geom.rows_in * d.C + geom.rows_out * d.K))
        if not _STATE["fuse_fin"]:
            L.check(lib.dp_bn_finalize(part.data_ptr(), nparts.value, d.K, d.Kp, float(geom.rows_out), gamma.data_ptr(),
                                       beta.data_ptr(), cfg.eps, cfg.momentum, _p(running_mean), _p(running_var),
                                       stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(),
                                       st), "dp_bn_finalize")
    else:
        if running_mean is None or running_var is None:
            raise L.DpError("eval-mode BatchNorm needs running statistics")
        stats.zero_()
        stats[0, :d.K] = running_mean
        stats[1, :d.K] = torch.rsqrt(running_var + cfg.eps)
        L.check(lib.dp_bn_eval_coeffs(running_mean.data_ptr(), running_var.data_ptr(), gamma.data_ptr(),
                                      beta.data_ptr(), cfg.eps, d.K, d.Kp, stats[2].data_ptr(), stats[3].data_ptr(),
                                      st), "dp_bn_eval_coeffs")
        if fuse_eval and _STATE["fuse_eval"]:
            # inference: BatchNorm (running statistics) + LeakyReLU [+ residual + LeakyReLU] in the conv epilogue; the raw
            # conv output never reaches HBM and nothing is kept for a backward pass
            if residual is not None and (residual.shape != y.shape or residual.dtype != y.dtype):
                raise L.DpError(f"residual {tuple(residual.shape)} {residual.dtype} does not match {tuple(y.shape)} {y.dtype}")
            t0 = _pb()
            if geom.stem:
                L.check(lib.dp_stem_conv_fwd_bnact(C.byref(d), x.data_ptr(), wf.data_ptr(), stats[2].data_ptr(), cfg.slope,
                                                   y.data_ptr(), st), "dp_stem_conv_fwd_bnact")
            else:
                xs = (C.c_longlong * 4)(*x_view) if x_view is not None else None
                L.check(lib.dp_conv_fwd_bnact(C.byref(d), xs, x.data_ptr() if x_ptr is None else x_ptr, wf.data_ptr(),
                                              stats[2].data_ptr(), cfg.slope, _p(residual), float(slope_res), y.data_ptr(),
                                              impl, st), "dp_conv_fwd_bnact")
            if t0 is not None:
                _pe(t0, geom.families(impl)[1], geom.flops, geom.esize * (geom.rows_in * d.C + geom.rows_out * d.K))
            return y, (x, None, None, stats, wd, geom)
        L.check(_conv_fwd_call(lib, geom, x, wf, y, None, None, impl, st), "dp_conv_fwd")
    z = torch.empty_like(y)
    if residual is not None and (residual.shape != y.shape or residual.dtype != y.dtype):
        raise L.DpError(f"residual {tuple(residual.shape)} {residual.dtype} does not match {tuple(y.shape)} {y.dtype}")
    t0 = _pb()
    L.check(lib.dp_bn_act_apply(y.data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(), cfg.slope, _p(residual),
                                float(slope_res), z.data_ptr(), geom.rows_out, d.Kp, d.dtype, st), "dp_bn_act_apply")
    if t0 is not None:
        _pe(t0, "bn_act_apply", 0.0, geom.esize * geom.rows_out * d.K * (3 if residual is not None else 2))
    return z, (x, y, z if residual is not None else None, stats, wd, geom)


def layer_backward(saved, dz, weight_shape, cfg: LayerCfg, training: bool, slope_res: float, need_dx: bool,
                   addend=None, want_dres: bool = False, pre_part=None, next_bn=None):
    """Returns (dx, dw, dgamma, dbeta, dres).

    `next_bn` = (y_prev, stats_prev, slope_prev, box) of the layer that produced this layer's input x: the data
    gradient then also writes the BatchNorm-backward sums of that layer (dp_conv_dgrad_bnstats) into box[0] =
    (part, nparts), which the caller hands to that layer's layer_backward as `pre_part` so its reduction pass is
    skipped."""
    lib = L.load()
    st = L.stream_ptr()
    x, y, out, stats, wd, geom = saved
    d = geom.desc
    dev = y.device
    dz = dz.contiguous()
    if dz.dtype != y.dtype or dz.shape != y.shape:
        raise L.DpError(f"grad {tuple(dz.shape)} {dz.dtype} does not match activation {tuple(y.shape)} {y.dtype}")
    mean, rstd, scale, shift = (stats[i].data_ptr() for i in range(4))
    elems = geom.esize * geom.rows_out * d.K
    fuse_fin = _STATE["fuse_fin"]
    if pre_part is not None and out is None and len(pre_part) == 3:
        dgamma, dbeta, coef = pre_part   # sums AND finalisation came out of the consumer's data-gradient launch
    else:
        if pre_part is not None and out is None:
            part, nparts = pre_part          # written by the consumer's data-gradient epilogue
            fused_here = False
        else:
            part = torch.empty((L.DP_MAX_PARTS, 2, d.Kp), dtype=torch.float32, device=dev)
            nparts = C.c_int(0)
            t0 = _pb()
            fused_here = fuse_fin
            if fused_here:
                fin, dgamma, dbeta, coef = _bn_fin_bwd(d.K, d.Kp, geom.rows_out, stats, training, dev)
                L.check(lib.dp_bn_act_bwd_reduce_fin(dz.data_ptr(), y.data_ptr(), _p(out), scale, shift, cfg.slope,
                                                     float(slope_res), part.data_ptr(), geom.rows_out, d.Kp, d.dtype,
                                                     C.byref(fin), st), "dp_bn_act_bwd_reduce_fin")
            else:
                L.check(lib.dp_bn_act_bwd_reduce(dz.data_ptr(), y.data_ptr(), _p(out), scale, shift, mean, rstd, cfg.slope,
                                                 float(slope_res), part.data_ptr(), C.byref(nparts), geom.rows_out, d.Kp,
                                                 d.dtype, st), "dp_bn_act_bwd_reduce")
            if t0 is not None:
                _pe(t0, "bn_act_bwd_reduce", 0.0, elems * (3 if out is not None else 2))
        if not fused_here:
            dgamma = torch.empty(d.K, dtype=torch.float32, device=dev)
            dbeta = torch.empty(d.K, dtype=torch.float32, device=dev)
            coef = torch.empty((2, d.Kp), dtype=torch.float32, device=dev)
            L.check(lib.dp_bn_bwd_finalize(part.data_ptr(), nparts.value, d.K, d.Kp, float(geom.rows_out), mean, rstd,
                                           dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), st), "dp_bn_bwd_finalize")
            if not training:
                coef.zero_()  # eval-mode BN: statistics are constants, no mean/variance terms
    dy = torch.empty_like(y)
    dres = torch.empty_like(y) if (want_dres and out is not None) else None
    t0 = _pb()
    L.check(lib.dp_bn_act_bwd_apply(dz.data_ptr(), y.data_ptr(), _p(out), scale, shift, mean, rstd, coef.data_ptr(),
                                    cfg.slope, float(slope_res), dy.data_ptr(), _p(dres), geom.rows_out, d.Kp,
                                    d.dtype, st), "dp_bn_act_bwd_apply")
    if t0 is not None:
        _pe(t0, "bn_act_bwd_apply", 0.0, elems * (3 + (1 if out is not None else 0) + (1 if dres is not None else 0)))
    impl = _STATE["impl"]
    dw = torch.empty(weight_shape, dtype=torch.float32, device=dev)
    if geom.ws_bytes is None:
        geom.ws_bytes = int(lib.dp_stem_wgrad_workspace(C.byref(d)) if geom.stem
                            else lib.dp_conv_wgrad_workspace(C.byref(d), impl))
    ws = _workspace(geom.ws_bytes, dev)
    io_bytes = geom.esize * (geom.rows_in * d.C + geom.rows_out * d.K)
    side = _side_stream(dev) if (_STATE["wgrad_stream"] and PROFILER is None) else None
    if side is not None:
        side.wait_stream(torch.cuda.current_stream())      # dy (and x) are complete on the main stream
        wctx = torch.cuda.stream(side)
    else:
        wctx = contextlib.nullcontext()
    with wctx:
        stw = L.stream_ptr()
        t0 = _pb()
        if geom.stem:
            if need_dx:
                raise L.DpError("the stem fast path has no data gradient")
            L.check(lib.dp_stem_conv_wgrad(C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel(),
                                           stw), "dp_stem_conv_wgrad")
        else:
            L.check(lib.dp_conv_wgrad(C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel(),
                                      impl, stw), "dp_conv_wgrad")
        if t0 is not None:
            _pe(t0, geom.families(impl)[3], geom.flops, io_bytes)
        if side is not None:
            done = torch.cuda.Event()
            done.record(side)
            _PENDING.append((done, (x, dy, dw, ws)))
    dx = None
    if need_dx:
        dx = torch.empty_like(x)
        if addend is not None:
            addend = addend.contiguous()
            if addend.shape != x.shape or addend.dtype != x.dtype:
                raise L.DpError("dgrad addend does not match the input activation")
        t0 = _pb()
        if next_bn is not None:
            y_prev, stats_prev, slope_prev, box = next_bn
            npart = torch.empty((L.DP_MAX_PARTS, 2, d.Cp), dtype=torch.float32, device=dev)
            if fuse_fin:
                # the producer's channels are this conv's input channels, its pixels this conv's input pixels
                fin_p, dg_p, db_p, coef_p = _bn_fin_bwd(d.C, d.Cp, geom.rows_in, stats_prev, training, dev)
                L.check(lib.dp_conv_dgrad_bnstats_fin(C.byref(d), dy.data_ptr(), wd.data_ptr(), _p(addend), dx.data_ptr(),
                                                      y_prev.data_ptr(), stats_prev[2].data_ptr(), float(slope_prev),
                                                      npart.data_ptr(), C.byref(fin_p), impl, st),
                        "dp_conv_dgrad_bnstats_fin")
                box.append((dg_p, db_p, coef_p))
            else:
                nn_ = C.c_int(0)
                L.check(lib.dp_conv_dgrad_bnstats(C.byref(d), dy.data_ptr(), wd.data_ptr(), _p(addend), dx.data_ptr(),
                                                  y_prev.data_ptr(), stats_prev[2].data_ptr(), float(slope_prev),
                                                  npart.data_ptr(), C.byref(nn_), impl, st), "dp_conv_dgrad_bnstats")
                box.append((npart, nn_))
        elif wd.dim() == 1:     # class-packed weights (pack_weights): every stride-parity class in one launch
            L.check(lib.dp_conv_dgrad_classes(C.byref(d), dy.data_ptr(), wd.data_ptr(), _p(addend), dx.data_ptr(), st),
                    "dp_conv_dgrad_classes")
        else:
            L.check(lib.dp_conv_dgrad(C.byref(d), dy.data_ptr(), wd.data_ptr(), _p(addend), dx.data_ptr(), impl, st),
                    "dp_conv_dgrad")
        if t0 is not None:
            _pe(t0, geom.families(impl)[2], geom.flops, io_bytes)
    elif addend is not None:
        dx = addend
    return dx, dw, dgamma, dbeta, dres


# ----------------------------------------------------------------------------------------------
# autograd wrappers
# ----------------------------------------------------------------------------------------------
class ConvBnActFn(torch.autograd.Function):
    """One Conv3dBlock.  inputs: x, weight, gamma, beta, (mod = the nn.Module holding buffers/cfg)."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, mod):
        training = mod.training
        rm, rv = (mod.bn.running_mean, mod.bn.running_var) if mod.bn.track_running_stats else (None, None)
        if not training and rm is None:
            training = True
        fused = not training and (not _GRAD[0] or not any(ctx.needs_input_grad))
        z, saved = layer_forward(x, weight, gamma, beta, rm, rv, mod._cfg, training, cache=mod._packed, fuse_eval=fused)
        xs, y, _, stats, wd, geom = saved
        if y is not None:
            ctx.save_for_backward(xs, y, stats, wd)
        ctx.geom, ctx.cfg, ctx.training, ctx.wshape = geom, mod._cfg, training, weight.shape
        return z

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dz):
        xs, y, stats, wd = ctx.saved_tensors
        dx, dw, dg, db, _ = layer_backward((xs, y, None, stats, wd, ctx.geom), dz, ctx.wshape, ctx.cfg, ctx.training,
                                           1.0, ctx.needs_input_grad[0])
        join_side_stream()
        return dx, dw, dg, db, None


class ConvBnActResFn(torch.autograd.Function):
    """Conv3d -> BatchNorm3d -> LeakyReLU(slope) -> (+ residual) -> LeakyReLU(slope_res) as ONE layer with the
    residual add fused into the BN-apply pass (the tail of a ResNet bottleneck, reference resnet.py:193-200, uses
    slope 1 = no activation before the add and slope_res 0 = ReLU after it).
    inputs: x, weight, gamma, beta, residual, mod, slope_res."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, residual, mod, slope_res):
        training = mod.training
        rm, rv = (mod.bn.running_mean, mod.bn.running_var) if mod.bn.track_running_stats else (None, None)
        if not training and rm is None:
            training = True
        fused = not training and (not _GRAD[0] or not any(ctx.needs_input_grad))
        z, saved = layer_forward(x, weight, gamma, beta, rm, rv, mod._cfg, training, residual=residual.contiguous(),
                                 slope_res=slope_res, cache=mod._packed, fuse_eval=fused)
        xs, y, out, stats, wd, geom = saved
        if y is not None:
            ctx.save_for_backward(xs, y, out, stats, wd)
        ctx.geom, ctx.cfg, ctx.training, ctx.wshape, ctx.slope_res = geom, mod._cfg, training, weight.shape, float(slope_res)
        return z

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dz):
        xs, y, out, stats, wd = ctx.saved_tensors
        dx, dw, dg, db, dres = layer_backward((xs, y, out, stats, wd, ctx.geom), dz, ctx.wshape, ctx.cfg, ctx.training,
                                              ctx.slope_res, ctx.needs_input_grad[0], want_dres=True)
        join_side_stream()
        return dx, dw, dg, db, dres, None, None


class ConvFn(torch.autograd.Function):
    """A bare Conv3d (bias-free) on internal tensors: the lateral connections of SlowFast
    (reference slowfast.py:56-63) have neither normalisation nor activation.  inputs: x, weight, mod."""

    @staticmethod
    def forward(ctx, x, weight, mod):
        lib = L.load()
        cfg = mod._cfg
        geom = conv_geom(cfg.C, cfg.K, cfg.kernel, cfg.stride, cfg.padding, x)
        wf, wd = pack_weights(weight, geom, x.dtype, mod._packed)
        y = torch.empty(geom.out_shape, dtype=x.dtype, device=x.device)
        L.check(lib.dp_conv_fwd(C.byref(geom.desc), x.data_ptr(), wf.data_ptr(), y.data_ptr(), None, None, _STATE["impl"],
                                L.stream_ptr()), "dp_conv_fwd")
        ctx.save_for_backward(x, wd)
        ctx.geom, ctx.wshape = geom, weight.shape
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dy):
        x, wd = ctx.saved_tensors
        lib = L.load()
        join_side_stream()   # this wgrad shares the split workspace with the ones on the side stream
        geom, d, st, impl = ctx.geom, ctx.geom.desc, L.stream_ptr(), _STATE["impl"]
        dy = dy.contiguous()
        dw = torch.empty(ctx.wshape, dtype=torch.float32, device=x.device)
        if geom.ws_bytes is None:
            geom.ws_bytes = int(lib.dp_conv_wgrad_workspace(C.byref(d), impl))
        ws = _workspace(geom.ws_bytes, x.device)
        L.check(lib.dp_conv_wgrad(C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), ws.numel(), impl,
                                  st), "dp_conv_wgrad")
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if wd.dim() == 1:
                L.check(lib.dp_conv_dgrad_classes(C.byref(d), dy.data_ptr(), wd.data_ptr(), None, dx.data_ptr(), st),
                        "dp_conv_dgrad_classes")
            else:
                L.check(lib.dp_conv_dgrad(C.byref(d), dy.data_ptr(), wd.data_ptr(), None, dx.data_ptr(), impl, st),
                        "dp_conv_dgrad")
        return dx, dw, None


_STEM_GEOMS = {}


def stem_geom(cfg: LayerCfg, B: int, T: int, H: int, W: int) -> Optional[ConvGeom]:
    """Geometry of the stem fast path for this Conv3dBlock and clip size, or None if it does not apply
    (fp32 validation mode, other kernels/strides, or the tcgen05 kernels disabled)."""
    if _STATE["dtype"] != torch.bfloat16 or _STATE["impl"] == L.IMPL_SIMT:
        return None
    key = (cfg.C, cfg.K, cfg.kernel, cfg.stride, cfg.padding, B, T, H, W)
    g = _STEM_GEOMS.get(key)
    if g is None:
        g = ConvGeom(cfg.C, cfg.K, cfg.kernel, cfg.stride, cfg.padding, B, T, H, W, L.DP_BF16)
        g.stem = True
        if not L.load().dp_stem_supported(C.byref(g.desc)):
            g = False
        _STEM_GEOMS[key] = g
    return g or None


def stem_pack_input(x: torch.Tensor, geom: ConvGeom, mean_bgr=None) -> torch.Tensor:
    """NCDHW fp32 clip (or (B,T,H,W,3) uint8 frames + BGR mean) -> packed rows XP for the stem fast path."""
    L.require_device()
    lib = L.load()
    d = geom.desc
    xp = torch.empty(int(lib.dp_stem_input_elems(C.byref(d))), dtype=torch.bfloat16, device=x.device)
    x = x.contiguous()
    if x.dtype == torch.uint8:
        m = (C.c_float * 3)(*[float(v) for v in mean_bgr])
        L.check(lib.dp_stem_pack_input_u8(C.byref(d), x.data_ptr(), m, xp.data_ptr(), L.stream_ptr()),
                "dp_stem_pack_input_u8")
    else:
        L.check(lib.dp_stem_pack_input_f32(C.byref(d), x.data_ptr(), xp.data_ptr(), L.stream_ptr()),
                "dp_stem_pack_input_f32")
    return xp


class StemConvBnActFn(torch.autograd.Function):
    """The stem Conv3dBlock on the packed-rows fast path.  inputs: xp (packed clip, no grad), weight, gamma,
    beta, mod, geom."""

    @staticmethod
    def forward(ctx, xp, weight, gamma, beta, mod, geom):
        training = mod.training
        rm, rv = (mod.bn.running_mean, mod.bn.running_var) if mod.bn.track_running_stats else (None, None)
        if not training and rm is None:
            training = True
        fused = not training and (not _GRAD[0] or not any(ctx.needs_input_grad))
        z, saved = layer_forward(xp, weight, gamma, beta, rm, rv, mod._cfg, training, cache=mod._packed_stem, geom=geom,
                                 fuse_eval=fused)
        _, y, _, stats, _, _ = saved
        if y is not None:
            ctx.save_for_backward(xp, y, stats)
        ctx.geom, ctx.cfg, ctx.training, ctx.wshape = geom, mod._cfg, training, weight.shape
        return z

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dz):
        xp, y, stats = ctx.saved_tensors
        _, dw, dg, db, _ = layer_backward((xp, y, None, stats, None, ctx.geom), dz, ctx.wshape, ctx.cfg, ctx.training,
                                          1.0, False)
        join_side_stream()
        return None, dw, dg, db, None, None


class AddActFn(torch.autograd.Function):
    """out = lrelu(a + b, slope): the residual tail of SpatioTemporalResBlock (R2Plus1D.py:187)
    when the block runs module by module (hooks registered on inner modules)."""

    @staticmethod
    def forward(ctx, a, b, slope):
        lib = L.load()
        Cp = a.shape[-1]
        ident = torch.zeros((2, Cp), dtype=torch.float32, device=a.device)
        ident[0].fill_(1.0)
        out = torch.empty_like(a)
        rows = a.numel() // Cp
        L.check(lib.dp_bn_act_apply(a.data_ptr(), ident[0].data_ptr(), ident[1].data_ptr(), 1.0, b.data_ptr(),
                                    float(slope), out.data_ptr(), rows, Cp, _code(a), L.stream_ptr()), "dp_bn_act_apply")
        ctx.save_for_backward(out)
        ctx.slope = float(slope)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        lib = L.load()
        g = g.contiguous()
        Cp = out.shape[-1]
        rows = out.numel() // Cp
        ident = torch.zeros((6, Cp), dtype=torch.float32, device=out.device)
        ident[0].fill_(1.0)  # scale=1; shift, mean, rstd, coef0, coef1 = 0
        d = torch.empty_like(out)
        dres = torch.empty_like(out)
        # y := out (any tensor; with rstd=0, coef=0 and slope=1 only the residual derivative acts)
        L.check(lib.dp_bn_act_bwd_apply(g.data_ptr(), out.data_ptr(), out.data_ptr(), ident[0].data_ptr(),
                                        ident[1].data_ptr(), ident[2].data_ptr(), ident[3].data_ptr(),
                                        ident[4].data_ptr(), 1.0, ctx.slope, d.data_ptr(), dres.data_ptr(), rows, Cp,
                                        _code(out), L.stream_ptr()), "dp_bn_act_bwd_apply")
        return d, dres, None


class ResBlockFn(torch.autograd.Function):
    """A whole SpatioTemporalResBlock (R2Plus1D.py:164-187) with a hand-ordered backward:
    the shortcut gradient enters conv1's dgrad as its epilogue addend, so no stand-alone
    gradient-sum pass exists.  inputs: x, block module, then (weight, gamma, beta) per layer in the
    order conv1.spatio, conv1.temporal, conv2.spatio, conv2.temporal[, ds.spatio, ds.temporal]."""

    @staticmethod
    def forward(ctx, x, block, *params):
        layers = block._dp_layers()
        training = block.training
        saved_all = []
        metas = []
        fused = not training and (not _GRAD[0] or not any(ctx.needs_input_grad))

        def run(i, inp, residual=None, slope_res=1.0):
            m = layers[i]
            w, g, b = params[3 * i:3 * i + 3]
            z, saved = layer_forward(inp, w, g, b, m.bn.running_mean, m.bn.running_var, m._cfg, training,
                                     residual=residual, slope_res=slope_res, cache=m._packed, fuse_eval=fused)
            xs, y, out, stats, wd, geom = saved
            if fused:
                return z
            saved_all.extend([xs, y, stats, wd])
            metas.append((geom, m._cfg, w.shape))
            return z

        h = run(0, x)
        h = run(1, h)
        h = run(2, h)
        if block.downsample:
            sc = run(4, x)   # metas index 3 = ds.spatio, 4 = ds.temporal (kept in call order)
            sc = run(5, sc)
        else:
            sc = x
        out = run(3, h, residual=sc, slope_res=block.relu.negative_slope)
        if fused:
            return out
        ctx.save_for_backward(out, *saved_all)
        ctx.metas, ctx.training, ctx.downsample = metas, training, block.downsample
        ctx.slope_res = float(block.relu.negative_slope)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        t = ctx.saved_tensors
        out = t[0]
        sv = [t[1 + 4 * i:1 + 4 * i + 4] for i in range(len(ctx.metas))]
        # call order in forward: 0 c1.s, 1 c1.t, 2 c2.s, [3 ds.s, 4 ds.t], last = c2.t
        last = len(ctx.metas) - 1

        fuse = _STATE.get("fuse_bn_reduce", True) and out.dtype == torch.bfloat16

        def bwd(slot, dz, need_dx=True, addend=None, with_out=False, want_dres=False, pre=None, nxt=None):
            xs, y, stats, wd = sv[slot]
            geom, cfg, wshape = ctx.metas[slot]
            box = []
            next_bn = None
            if nxt is not None and fuse and need_dx and _fused_bnstats(geom):
                _, y_n, stats_n, _ = sv[nxt]
                next_bn = (y_n, stats_n, ctx.metas[nxt][1].slope, box)
            res = layer_backward((xs, y, out if with_out else None, stats, wd, geom), dz, wshape, cfg, ctx.training,
                                 ctx.slope_res, need_dx, addend=addend, want_dres=want_dres, pre_part=pre, next_bn=next_bn)
            return res + (box[0] if box else None,)

        # the data gradient of layer i+1 also produces the BatchNorm-backward sums of layer i (inside the block the
        # producer of every conv input is known): c2.t -> c2.s -> c1.t -> c1.s
        grads = {}
        d, dw, dg, db, dres, pp = bwd(last, dout, with_out=True, want_dres=True, nxt=2)
        grads[3] = (dw, dg, db)
        d, dw, dg, db, _, pp = bwd(2, d, pre=pp, nxt=1)
        grads[2] = (dw, dg, db)
        d, dw, dg, db, _, pp = bwd(1, d, pre=pp, nxt=0)
        grads[1] = (dw, dg, db)
        need_dx = ctx.needs_input_grad[0]
        if ctx.downsample:
            ds, dw, dg, db, _, pq = bwd(4, dres, nxt=3)
            grads[5] = (dw, dg, db)
            ds, dw, dg, db, _, _ = bwd(3, ds, need_dx=need_dx, pre=pq)
            grads[4] = (dw, dg, db)
            dx, dw, dg, db, _, _ = bwd(0, d, need_dx=need_dx, addend=ds, pre=pp)
        else:
            dx, dw, dg, db, _, _ = bwd(0, d, need_dx=need_dx, addend=dres, pre=pp)
        grads[0] = (dw, dg, db)
        flat = []
        for i in range(len(ctx.metas)):
            flat.extend(grads[i])
        join_side_stream()
        return (dx, None, *flat)


def _caller(fn):
    """Entry point of a layer Function for module code: records whether the CALLER runs under autograd (inside
    Function.forward grad mode is always off, and `needs_input_grad` is True for parameters even under no_grad), so
    eval-mode inference takes the fused conv + BatchNorm + activation kernels."""
    def call(*args):
        _GRAD[0] = torch.is_grad_enabled()
        try:
            return fn.apply(*args)
        finally:
            _GRAD[0] = True
    return call


conv_bn_act = _caller(ConvBnActFn)
conv_bn_act_res = _caller(ConvBnActResFn)
stem_conv_bn_act = _caller(StemConvBnActFn)
res_block = _caller(ResBlockFn)


class AvgPoolFn(torch.autograd.Function):
    """AdaptiveAvgPool3d(1) + view (R2Plus1D.py:215,224-225): internal (B,T,H,W,Cp) -> (B,C) fp32."""

    @staticmethod
    def forward(ctx, x, c):
        B = x.shape[0]
        Cp = x.shape[-1]
        pixels = x.numel() // (B * Cp)
        out = torch.empty((B, c), dtype=torch.float32, device=x.device)
        L.check(L.load().dp_avgpool_fwd(x.data_ptr(), out.data_ptr(), B, pixels, c, Cp, _code(x), L.stream_ptr()),
                "dp_avgpool_fwd")
        ctx.shape, ctx.dtype, ctx.c = x.shape, x.dtype, c
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        g = g.contiguous().float()
        B, Cp = ctx.shape[0], ctx.shape[-1]
        dx = torch.empty(ctx.shape, dtype=ctx.dtype, device=g.device)
        pixels = dx.numel() // (B * Cp)
        code = L.DP_BF16 if ctx.dtype == torch.bfloat16 else L.DP_F32
        L.check(L.load().dp_avgpool_bwd(g.data_ptr(), dx.data_ptr(), B, pixels, ctx.c, Cp, code, L.stream_ptr()),
                "dp_avgpool_bwd")
        return dx, None


# ----------------------------------------------------------------------------------------------
# SlowFast auxiliaries (csrc/slowfast_ops.cu)
# ----------------------------------------------------------------------------------------------
class SeSwishFn(torch.autograd.Function):
    """swish(x * gate[b, c]) on an internal tensor; gate = None gives the plain Swish
    (reference resnet.py:63-81 SwishEfficient; the squeeze-excite scaling of Bottleneck3D.forward :186-192 fused in)."""

    @staticmethod
    def forward(ctx, x, gate, c):
        B, Cp = x.shape[0], x.shape[-1]
        pixels = x.numel() // (B * Cp)
        x = x.contiguous()
        g = None if gate is None else gate.contiguous().float()
        out = torch.empty_like(x)
        L.check(L.load().dp_se_swish_fwd(x.data_ptr(), _p(g), out.data_ptr(), B, pixels, c, Cp, _code(x), L.stream_ptr()),
                "dp_se_swish_fwd")
        ctx.save_for_backward(x, g)
        ctx.c = c
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        x, g = ctx.saved_tensors
        B, Cp = x.shape[0], x.shape[-1]
        pixels = x.numel() // (B * Cp)
        dout = dout.contiguous()
        dx = torch.empty_like(x)
        dgate = None if g is None else torch.empty((B, ctx.c), dtype=torch.float32, device=x.device)
        L.check(L.load().dp_se_swish_bwd(x.data_ptr(), _p(g), dout.data_ptr(), dx.data_ptr(), _p(dgate), B, pixels, ctx.c, Cp,
                                         _code(x), L.stream_ptr()), "dp_se_swish_bwd")
        return dx, dgate, None


def se_swish(x: torch.Tensor, gate: Optional[torch.Tensor]) -> torch.Tensor:
    return tag(SeSwishFn.apply(x, gate, x._dp_c), x._dp_c)


class MaxPoolHWFn(torch.autograd.Function):
    """nn.MaxPool3d((1,3,3),(1,2,2),(0,1,1)) on an internal tensor (reference resnet.py:220-225)."""

    @staticmethod
    def forward(ctx, x):
        B, T, H, W, Cp = x.shape
        x = x.contiguous()
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        out = torch.empty((B, T, Ho, Wo, Cp), dtype=x.dtype, device=x.device)
        need_idx = any(ctx.needs_input_grad)
        idx = torch.empty(out.shape, dtype=torch.uint8, device=x.device) if need_idx else None
        L.check(L.load().dp_maxpool_hw_fwd(x.data_ptr(), out.data_ptr(), _p(idx), B * T, H, W, Cp, _code(x), L.stream_ptr()),
                "dp_maxpool_hw_fwd")
        if need_idx:
            ctx.save_for_backward(idx)
        ctx.shape, ctx.dtype = x.shape, x.dtype
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        B, T, H, W, Cp = ctx.shape
        dout = dout.contiguous()
        dx = torch.empty(ctx.shape, dtype=ctx.dtype, device=dout.device)
        code = L.DP_BF16 if ctx.dtype == torch.bfloat16 else L.DP_F32
        L.check(L.load().dp_maxpool_hw_bwd(dout.data_ptr(), idx.data_ptr(), dx.data_ptr(), B * T, H, W, Cp, code,
                                           L.stream_ptr()), "dp_maxpool_hw_bwd")
        return dx


def maxpool_hw(x: torch.Tensor) -> torch.Tensor:
    return tag(MaxPoolHWFn.apply(x), x._dp_c)


class ConcatChannelsFn(torch.autograd.Function):
    """torch.cat([a, b], dim=1) of the reference's NCDHW tensors (slowfast.py:26-36) on internal tensors."""

    @staticmethod
    def forward(ctx, a, b, ca, cb):
        if a.shape[:-1] != b.shape[:-1] or a.dtype != b.dtype:
            raise L.DpError(f"concat: {tuple(a.shape)} {a.dtype} vs {tuple(b.shape)} {b.dtype}")
        a, b = a.contiguous(), b.contiguous()
        cop = ceil16(ca + cb)
        rows = a.numel() // a.shape[-1]
        out = torch.empty((*a.shape[:-1], cop), dtype=a.dtype, device=a.device)
        L.check(L.load().dp_concat_channels(a.data_ptr(), b.data_ptr(), out.data_ptr(), rows, ca, a.shape[-1], cb, b.shape[-1],
                                            cop, _code(a), L.stream_ptr()), "dp_concat_channels")
        ctx.meta = (a.shape, b.shape, ca, cb, cop, rows)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        sa, sb, ca, cb, cop, rows = ctx.meta
        dout = dout.contiguous()
        da = torch.empty(sa, dtype=dout.dtype, device=dout.device)
        db = torch.empty(sb, dtype=dout.dtype, device=dout.device)
        L.check(L.load().dp_split_channels(dout.data_ptr(), da.data_ptr(), db.data_ptr(), rows, ca, sa[-1], cb, sb[-1], cop,
                                           _code(dout), L.stream_ptr()), "dp_split_channels")
        return da, db, None, None


def concat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return tag(ConcatChannelsFn.apply(a, b, a._dp_c, b._dp_c), a._dp_c + b._dp_c)


def global_avgpool(x: torch.Tensor) -> torch.Tensor:
    """AdaptiveAvgPool3d(1) + flatten: internal tensor -> (B, C) fp32."""
    return AvgPoolFn.apply(x, x._dp_c)


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------
class LossFn(torch.autograd.Function):
    """Fused forward+backward of CE / Focal / LDAM (reference src/loss.py:14-81)."""

    @staticmethod
    def forward(ctx, logits, target, weight, margins, kind, gamma, s):
        L.require_device()
        if not logits.is_cuda or logits.dim() != 2:
            raise L.DpError("loss: logits must be a CUDA tensor of shape (N, num_classes)")
        lg = logits.contiguous().float()
        tg = target.contiguous().view(-1)
        if tg.dtype != torch.int64:
            tg = tg.long()
        if tg.device != lg.device:
            tg = tg.to(lg.device)
        n, c = lg.shape
        if tg.numel() != n:
            raise L.DpError(f"loss: {n} rows of logits but {tg.numel()} targets")
        dev = lg.device
        res = torch.empty(2, dtype=torch.float32, device=dev)
        dlogits = torch.empty_like(lg)
        lib = L.load()
        ws = _zeros_ws("loss", int(lib.dp_loss_workspace(n)), dev)
        L.check(lib.dp_loss_fwd_bwd(kind, lg.data_ptr(), tg.data_ptr(), _p(weight), _p(margins), float(gamma),
                                    float(s), n, c, res.data_ptr(), dlogits.data_ptr(), ws.data_ptr(), L.stream_ptr()),
                "dp_loss_fwd_bwd")
        ctx.save_for_backward(dlogits, res)
        ctx.in_dtype = logits.dtype
        return res[0].clone()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        dlogits, res = ctx.saved_tensors
        g = g.contiguous().float()
        out = torch.empty_like(dlogits)
        L.check(L.load().dp_loss_bwd_scale(dlogits.data_ptr(), g.data_ptr(), res.data_ptr(), out.data_ptr(),
                                           out.numel(), L.stream_ptr()), "dp_loss_bwd_scale")
        return out.to(ctx.in_dtype), None, None, None, None, None, None
