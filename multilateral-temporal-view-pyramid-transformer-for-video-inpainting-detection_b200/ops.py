"""Tensor-level wrappers over the C ABI (include/mumpy_b200.h).

PyTorch is used here for device memory and the current stream only; every computation is a libmumpy_b200
kernel.  All tensors must live on a CUDA device and be contiguous -- anything else raises (no fallback).
"""
import ctypes
import os

import torch

from . import _lib

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3
RS_IDENTITY, RS_UP_ALIGNED, RS_UP_HALFPIX, RS_AVGPOOL2, RS_PIXEL_SHUFFLE2 = 0, 1, 2, 3, 4

# Default operand precision: IEEE half.  Same tcgen05 kernels and speed as bfloat16, three more mantissa bits on every GEMM /
# attention operand: the mode that meets the north star's >= 99.9 % thresholded-mask identity (bf16: 99.6 %).  Its narrower
# range is guarded: conversions saturate and report through the overflow flag below (f16_overflowed()).
DEFAULT_PRECISION = "fp16"
_precision = DEFAULT_PRECISION
_bound_device = None
_overflow_flags = {}        # device index -> int32 tensor registered with mumpy_set_f16_overflow_flag
launch_count = 0            # libmumpy_b200 kernels launched so far (bench.py reports the per-step delta as gpu_launches)


def set_precision(mode: str):
    """'bf16' / 'fp16': tcgen05 GEMMs on 16-bit operands (bfloat16 or IEEE half: same tensor-core rate, half has 3 more
    mantissa bits and a narrower range) with fp32 accumulation, residual stream and statistics; 'fp32': exact fp32 FMA kernels."""
    global _precision
    if mode not in ("bf16", "fp16", "fp32"):
        raise ValueError("precision must be 'bf16', 'fp16' or 'fp32'")
    _precision = mode


def precision() -> str:
    return _precision


def act_dtype():
    """torch dtype of GEMM operands in the current precision mode."""
    return {"bf16": torch.bfloat16, "fp16": torch.float16}.get(_precision, torch.float32)


def tensor_cores() -> bool:
    """True in the 16-bit operand modes (tcgen05 GEMMs / implicit-GEMM convolutions)."""
    return _precision != "fp32"


def code(dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    if dtype == torch.float16:
        return F16
    raise _lib.MumpyError("unsupported dtype %s" % dtype)


def _prep(*tensors):
    """Validates tensors, binds the library to their device, returns (lib, stream)."""
    global _bound_device, launch_count
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.MumpyError("libmumpy_b200 needs CUDA tensors (got %s); there is no CPU fallback" % t.device)
        if not t.is_contiguous():
            raise _lib.MumpyError("non-contiguous tensor passed to a libmumpy_b200 op")
        dev = t.device.index if dev is None else dev
        if t.device.index != dev:
            raise _lib.MumpyError("tensors on different devices")
    lib = _lib.load()
    if dev != _bound_device:
        if _bound_device is not None:
            # the library caches per-function attributes (opt-in shared-memory sizes, SM counts) once per process, and they are
            # per device: the supported deployment is one process per GPU (torchrun), as bench.py and evaluate.py run it
            raise _lib.MumpyError("libmumpy_b200 is bound to cuda:%d in this process; use one process per GPU (got a tensor on cuda:%d)"
                                  % (_bound_device, dev))
        _lib.check(lib.mumpy_init(dev), "mumpy_init")
        if dev not in _overflow_flags:
            _overflow_flags[dev] = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", dev))
            torch.cuda.synchronize(dev)
        _lib.check(lib.mumpy_set_f16_overflow_flag(_overflow_flags[dev].data_ptr()), "mumpy_set_f16_overflow_flag")
        _bound_device = dev
    launch_count += 1
    return lib, torch.cuda.current_stream(dev).cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def f16_overflow_flag(device=None):
    """The int32 device word the kernels OR 1 into when an fp32 value beyond the IEEE-half range was converted (and saturated)
    in 'fp16' mode; None before the first library call on that device.  Copy it back with a step's results to check cheaply."""
    dev = None if device is None else torch.device(device).index      # torch.device("cuda") has no index: the current device
    if dev is None:
        dev = torch.cuda.current_device()
    return _overflow_flags.get(dev)


def f16_overflowed(device=None, reset=True) -> bool:
    """True when any kernel since the last reset had to saturate an fp32 -> half conversion (synchronises the device)."""
    flag = f16_overflow_flag(device)
    if flag is None:
        return False
    hit = bool(flag.item())
    if hit and reset:
        flag.zero_()
    return hit


def check_f16_range(device=None):
    """Raises MumpyError when the 'fp16' mode left the half range since the last check -- loud, never silent."""
    if f16_overflowed(device):
        raise _lib.MumpyError("fp16 operand overflow: an activation beyond +-65504 was saturated; results are not trustworthy. "
                              "Use set_precision('bf16') (same speed, wider range, ~99.6 % mask identity) or 'fp32' for this model.")


# ------------------------------------------------------------------------------------------------ GEMM / norms
def linear(a, w, bias=None, residual=None, act=ACT_NONE, out_dtype=torch.float32, out=None):
    """out = act(a @ w.T + bias) (+ residual).  a (..., K), w (N, K) same dtype; residual/out (..., N)."""
    K = a.shape[-1]
    N = w.shape[0]
    M = a.numel() // K
    if w.shape[1] != K or a.dtype != w.dtype:
        raise _lib.MumpyError("linear: operand mismatch a%s %s w%s %s" % (tuple(a.shape), a.dtype, tuple(w.shape), w.dtype))
    if out is None:
        out = torch.empty(a.shape[:-1] + (N,), dtype=out_dtype, device=a.device)
    lib, st = _prep(a, w, bias, residual, out)
    _lib.check(lib.mumpy_linear(_p(a), K, _p(w), _p(bias), _p(residual), _p(out), N, M, N, K, code(a.dtype),
                                code(out.dtype), act, st), "mumpy_linear")
    return out


def linear_dual(a, w, bias, residual):
    """16-bit operands only: returns (a @ w.T + bias + residual as fp32, (operand dtype)(a @ w.T + bias))."""
    K = a.shape[-1]
    N = w.shape[0]
    M = a.numel() // K
    if w.shape[1] != K or a.dtype != w.dtype or a.dtype not in (torch.bfloat16, torch.float16):
        raise _lib.MumpyError("linear_dual: 16-bit operands of one type required")
    out = torch.empty(a.shape[:-1] + (N,), dtype=torch.float32, device=a.device)
    aux = torch.empty(a.shape[:-1] + (N,), dtype=a.dtype, device=a.device)
    lib, st = _prep(a, w, bias, residual, out, aux)
    _lib.check(lib.mumpy_linear_dual(_p(a), K, _p(w), _p(bias), _p(residual), _p(out), _p(aux), N, M, N, K, code(a.dtype), ACT_NONE, st),
               "mumpy_linear_dual")
    return out, aux


def linear_into(a, lda, w, bias, out, ldo, M, N, K, act=ACT_NONE, residual=None):
    """Strided form: a/out are base tensors (possibly wider matrices) with explicit row strides."""
    lib, st = _prep(a, w, bias, out)
    _lib.check(lib.mumpy_linear(_p(a), lda, _p(w), _p(bias), _p(residual), _p(out), ldo, M, N, K, code(a.dtype),
                                code(out.dtype), act, st), "mumpy_linear")
    return out


def layernorm(x, gamma, beta, eps=1e-5, out_dtype=None):
    C = x.shape[-1]
    out = torch.empty(x.shape, dtype=out_dtype or act_dtype(), device=x.device)
    lib, st = _prep(x, gamma, beta, out)
    _lib.check(lib.mumpy_layernorm(_p(x), _p(gamma), _p(beta), _p(out), code(out.dtype), x.numel() // C, C, eps, st),
               "mumpy_layernorm")
    return out


# LayerNorm inside the consuming GEMM (mumpy_ln_linear): bit-identical to mumpy_layernorm + mumpy_linear.  On for the K = 512 blocks
# (stage 2 of view 3: one CTA pair per pair of row tiles, the weight ring at the tensor floor; 2361-2380 vs 2355 clips/s on one box
# and 36 LayerNorm launches fewer per forward); the narrower widths measured slower (K = 384: 2322 clips/s -- the un-overlapped
# LayerNorm prologue costs more than the removed kernel) and stay on the unfused kernels unless listed in
# MUMPY_FUSED_LN_WIDTHS.
FUSED_LN = os.environ.get("MUMPY_FUSED_LN", "1") != "0"
FUSED_LN_WIDTHS = tuple(int(v) for v in os.environ.get("MUMPY_FUSED_LN_WIDTHS", "512").split(",") if v)


def set_fused_ln(enabled: bool, widths=None):
    """Switches the fused LayerNorm + GEMM path; `widths` (optional) replaces the set of K it applies to."""
    global FUSED_LN, FUSED_LN_WIDTHS
    FUSED_LN = bool(enabled)
    if widths is not None:
        FUSED_LN_WIDTHS = tuple(int(w) for w in widths)


def set_ln_linear_pair_mode(mode: int):
    """CTA-pair policy of ln_linear: 0 never, 1 the cost model decides (default), 2 whenever the shape allows."""
    _lib.check(_lib.load().mumpy_set_ln_linear_pair_mode(int(mode)), "mumpy_set_ln_linear_pair_mode")


def ln_linear_fits(N, K) -> bool:
    """True when ln_linear() can run: a 16-bit operand mode, the switch on, and a width the fused kernel holds on chip."""
    return FUSED_LN and tensor_cores() and K in (96, 128, 192, 256, 384, 512) and K in FUSED_LN_WIDTHS and (N % 64 == 0 or N % 96 == 0)


def ln_linear(x, gamma, beta, eps, w, bias=None, act=ACT_NONE):
    """act(LayerNorm(x) @ w.T + bias) in operand precision; x (..., K) fp32, w (N, K) 16-bit.  One kernel: the normalised rows only
    exist in shared memory (mumpy_ln_linear)."""
    K = x.shape[-1]
    N = w.shape[0]
    M = x.numel() // K
    if w.shape[1] != K or x.dtype != torch.float32 or w.dtype not in (torch.bfloat16, torch.float16):
        raise _lib.MumpyError("ln_linear: x%s %s w%s %s" % (tuple(x.shape), x.dtype, tuple(w.shape), w.dtype))
    out = torch.empty(x.shape[:-1] + (N,), dtype=w.dtype, device=x.device)
    lib, st = _prep(x, gamma, beta, w, bias, out)
    _lib.check(lib.mumpy_ln_linear(_p(x), _p(gamma), _p(beta), float(eps), _p(w), _p(bias), _p(out), N, M, N, K, code(w.dtype), act, st),
               "mumpy_ln_linear")
    return out


FUSED_MLP = os.environ.get("MUMPY_FUSED_MLP", "1") != "0"      # LayerNorm + fc1 + GELU + fc2 + residual in one kernel for C <= 256


def set_fused_mlp(enabled: bool):
    global FUSED_MLP
    FUSED_MLP = bool(enabled)


# widths that take the fused kernel (the kernel supports 96, 128, 192 and 256)
FUSED_MLP_WIDTHS = tuple(int(v) for v in os.environ.get("MUMPY_FUSED_MLP_WIDTHS", "96,128,256").split(",") if v)


def set_mlp_fused_shape(shape: int):
    """Kernel shape of mlp_fused: 0 pipelined (default), 1 serial, 2 serial with two CTAs per SM (C <= 128); bit-identical results."""
    _lib.check(_lib.load().mumpy_set_mlp_fused_shape(int(shape)), "mumpy_set_mlp_fused_shape")


def mlp_fused_fits(C) -> bool:
    return FUSED_MLP and tensor_cores() and C in (96, 128, 192, 256) and C in FUSED_MLP_WIDTHS


def mlp_fused(x, gamma, beta, eps, w1, b1, w2, b2):
    """x + fc2(gelu(fc1(LayerNorm(x)))) in one kernel (mumpy_mlp_fused); x (..., C) fp32, w1 (4C, C), w2 (C, 4C) 16-bit."""
    C = x.shape[-1]
    M = x.numel() // C
    if tuple(w1.shape) != (4 * C, C) or tuple(w2.shape) != (C, 4 * C) or x.dtype != torch.float32 or w1.dtype != w2.dtype:
        raise _lib.MumpyError("mlp_fused: x%s w1%s w2%s" % (tuple(x.shape), tuple(w1.shape), tuple(w2.shape)))
    out = torch.empty_like(x)
    lib, st = _prep(x, gamma, beta, w1, b1, w2, b2, out)
    _lib.check(lib.mumpy_mlp_fused(_p(x), _p(gamma), _p(beta), float(eps), _p(w1), _p(b1), _p(w2), _p(b2), _p(out), M, C, code(w1.dtype), st),
               "mumpy_mlp_fused")
    return out


def patch_merge_norm(x, gamma, beta, B, TH, W, C, eps=1e-5, out_dtype=None):
    out = torch.empty((B, (TH // 2) * (W // 2), 4 * C), dtype=out_dtype or act_dtype(), device=x.device)
    lib, st = _prep(x, gamma, beta, out)
    _lib.check(lib.mumpy_patch_merge_norm(_p(x), _p(gamma), _p(beta), _p(out), code(out.dtype), B, TH, W, C, eps, st),
               "mumpy_patch_merge_norm")
    return out


# ------------------------------------------------------------------------------------------------ attention
def window_attention(qkv, bias, mask, B, TH, W, C, heads, ws, shift, rel_table=None, standard_mask=False):
    out = torch.empty((B, TH * W, C), dtype=qkv.dtype, device=qkv.device)
    lib, st = _prep(qkv, bias, mask, rel_table, out)
    _lib.check(lib.mumpy_window_attention(_p(qkv), _p(bias), _p(mask), _p(rel_table), int(standard_mask), _p(out), code(qkv.dtype),
                                          B, TH, W, C, heads, ws, shift, st), "mumpy_window_attention")
    return out


def mha_short(qkv, Bn, N, C, heads):
    out = torch.empty((Bn, N, C), dtype=qkv.dtype, device=qkv.device)
    lib, st = _prep(qkv, out)
    _lib.check(lib.mumpy_mha_short(_p(qkv), _p(out), code(qkv.dtype), Bn, N, C, heads, st), "mumpy_mha_short")
    return out


def mha_short_probs(qkv, Bn, N, C, heads):
    """Attention map of mha_short: (Bn, heads, N, N) fp32 (diagnostic output of blocks.py:66-68)."""
    probs = torch.empty((Bn, heads, N, N), dtype=torch.float32, device=qkv.device)
    lib, st = _prep(qkv, probs)
    _lib.check(lib.mumpy_mha_short_probs(_p(qkv), _p(probs), code(qkv.dtype), Bn, N, C, heads, st), "mumpy_mha_short_probs")
    return probs


# ------------------------------------------------------------------------------------------------ front end
def tokenize(x, w_kc, bias, gamma, beta, kt, eps=1e-5):
    B, T, _, S, _ = x.shape
    C = w_kc.shape[1]
    To = T // kt
    out = torch.empty((B, To * (S // 4) ** 2, C), dtype=torch.float32, device=x.device)
    lib, st = _prep(x, w_kc, bias, gamma, beta, out)
    _lib.check(lib.mumpy_tokenize(_p(x), _p(w_kc), _p(bias), _p(gamma), _p(beta), _p(out), B, T, S, kt, C, eps, st),
               "mumpy_tokenize")
    return out


def patchify16(x, kt, dtype=None):
    """Patches of Conv3d(k=s=(kt,4,4)) as a split GEMM operand: (B,T,3,S,S) fp32 -> (B*To*(S/4)^2, 3*K) [hi | hi | lo], K = 3*kt*16."""
    B, T, _, S, _ = x.shape
    dtype = dtype or act_dtype()
    K = 3 * kt * 16
    out = torch.empty((B * (T // kt) * (S // 4) ** 2, 3 * K), dtype=dtype, device=x.device)
    lib, st = _prep(x, out)
    _lib.check(lib.mumpy_patchify16(_p(x), _p(out), code(dtype), B, T, S, kt, st), "mumpy_patchify16")
    return out


def faf(x, dct, bands, frame=1):
    B, T, _, S, _ = x.shape
    ws = torch.empty((_lib.load().mumpy_faf_workspace_floats(B, S),), dtype=torch.float32, device=x.device)
    out = torch.empty((B, 9, S, S), dtype=torch.float32, device=x.device)
    lib, st = _prep(x, dct, ws, out)
    arr = (ctypes.c_int * 6)(*[int(v) for lohi in bands for v in lohi])
    global launch_count
    launch_count += 3            # four GEMM passes per call
    _lib.check(lib.mumpy_faf(_p(x), _p(dct), _p(ws), _p(out), B, T, frame, S, arr, st), "mumpy_faf")
    return out


def faf16(x, dcat, dtcat, bands, frame=1):
    """Tensor-core FAF (16-bit modes): dcat / dtcat (S, 3S) = [D_hi | D_lo | D_hi] for D and D^T in the operand type."""
    B, T, _, S, _ = x.shape
    ws16 = torch.empty((_lib.load().mumpy_faf16_workspace_bytes(B, S) // 2,), dtype=dcat.dtype, device=x.device)      # the four passes ping-pong between two halves
    out = torch.empty((B, 9, S, S), dtype=torch.float32, device=x.device)
    lib, st = _prep(x, dcat, dtcat, ws16, out)
    arr = (ctypes.c_int * 6)(*[int(v) for lohi in bands for v in lohi])
    global launch_count
    launch_count += 4            # the input split + four GEMMs per call (transposes / band masks / splits live in their epilogues)
    _lib.check(lib.mumpy_faf16(_p(x), _p(dcat), _p(dtcat), _p(ws16), None, _p(out), B, T, frame, S, arr, code(dcat.dtype), st),
               "mumpy_faf16")
    return out


# ------------------------------------------------------------------------------------------------ deformable cross-view attention
def cva_offsets(q, dw_w, dw_b, ln_g, ln_b, pw, B, TH1, W, C, groups, ws):
    N1 = B * (TH1 // ws) * (W // ws)
    pix = torch.empty((N1, groups, ws * ws, 2), dtype=torch.float32, device=q.device)
    lib, st = _prep(q, dw_w, dw_b, ln_g, ln_b, pw, pix)
    _lib.check(lib.mumpy_cva_offsets(_p(q), _p(dw_w), _p(dw_b), _p(ln_g), _p(ln_b), _p(pw), _p(pix), B, TH1, W, C, groups,
                                     ws, st), "mumpy_cva_offsets")
    return pix


def cva_sample(x2, pix, B, TH1, TH2, W, C, groups, ws, per_clip, out_dtype):
    N2 = B * (TH2 // ws) * (W // ws)
    out = torch.empty((N2 * ws * ws, C), dtype=out_dtype, device=x2.device)
    lib, st = _prep(x2, pix, out)
    _lib.check(lib.mumpy_cva_sample(_p(x2), code(x2.dtype), _p(pix), _p(out), code(out_dtype), B, TH1, TH2, W, C, groups, ws,
                                    int(per_clip), st), "mumpy_cva_sample")
    return out


def cva_attention(q, kv, B, TH1, TH2, W, C, heads, ws, per_clip):
    N1 = B * (TH1 // ws) * (W // ws)
    out = torch.empty((N1 * ws * ws, C), dtype=kv.dtype, device=q.device)
    lib, st = _prep(q, kv, out)
    _lib.check(lib.mumpy_cva_attention(_p(q), _p(kv), code(kv.dtype), _p(out), code(out.dtype), B, TH1, TH2, W, C, heads,
                                       ws, int(per_clip), st), "mumpy_cva_attention")
    return out


def cva_attention_probs(q, kv, B, TH1, TH2, W, C, heads, ws, per_clip):
    """Attention map of cva_attention in the reference's layout (N1, r * heads, P, P) fp32 (deformableAttention.py:364,389,399)."""
    N1 = B * (TH1 // ws) * (W // ws)
    r = TH2 // TH1
    probs = torch.empty((N1, r * heads, ws * ws, ws * ws), dtype=torch.float32, device=q.device)
    lib, st = _prep(q, kv, probs)
    _lib.check(lib.mumpy_cva_attention_probs(_p(q), _p(kv), code(kv.dtype), _p(probs), B, TH1, TH2, W, C, heads, ws, int(per_clip), st),
               "mumpy_cva_attention_probs")
    return probs


def cva_residual(h, y, B, TH1, W, C, ws):
    out = torch.empty_like(h)
    lib, st = _prep(h, y, out)
    _lib.check(lib.mumpy_cva_residual(_p(h), _p(y), _p(out), B, TH1, W, C, ws, st), "mumpy_cva_residual")
    return out


# ------------------------------------------------------------------------------------------------ data movement / decoder
def gather_rows(src, C, dst, dst_ld, dst_col, B, rows_out, rows_src, div=1, mul_hi=1, mul_lo=0, add=0):
    lib, st = _prep(src, dst)
    _lib.check(lib.mumpy_gather_rows(_p(src), C, _p(dst), code(dst.dtype), dst_ld, dst_col, B, rows_out, rows_src, div,
                                     mul_hi, mul_lo, add, st), "mumpy_gather_rows")
    return dst


def conv2d_nhwc(x, w_ohwi, bias, B, H, W, Cin, Cout, kh, kw, ph, pw, ld_in=None, out=None, ld_out=None):
    if out is None:
        out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x.device)
    lib, st = _prep(x, w_ohwi, bias, out)
    _lib.check(lib.mumpy_conv2d_nhwc(_p(x), ld_in or Cin, _p(w_ohwi), _p(bias), _p(out), ld_out or Cout, B, H, W, Cin,
                                     Cout, kh, kw, ph, pw, st), "mumpy_conv2d_nhwc")
    return out


CONV_SPLIT_K = os.environ.get("MUMPY_CONV_SPLITK", "1") != "0"      # split-K for small-map / long-reduction convolutions


def conv2d_nhwc_bf16(x, w_packed, bias, B, H, W, Cin, Cout, kh, kw, ph, pw, ld_in=None, act=ACT_NONE, out_dtype=torch.float32,
                     residual=None, split_k=None):
    """Tensor-core implicit GEMM: x (B,H,W,Cin[ld_in]) NHWC and w_packed (Cout, kh*kw*ceil(Cin/64)*64) of one 16-bit type."""
    if x.dtype != w_packed.dtype:
        raise _lib.MumpyError("conv2d_nhwc_bf16: activation %s vs filter %s" % (x.dtype, w_packed.dtype))
    out = torch.empty((B, H, W, Cout), dtype=out_dtype, device=x.device)
    # split-K workspace for small maps with a long reduction (few 128-row tiles, >= 32 k-blocks): one fp32 partial per k range
    ws = None
    m_tiles, kblocks = (B * H * W + 127) // 128, kh * kw * ((Cin + 63) // 64)
    if split_k is None:
        split_k = CONV_SPLIT_K
    if split_k and m_tiles * 2 <= 148 and kblocks >= 32 and Cout % 16 == 0 and Cout <= 256:
        ws = torch.empty((min(148 // m_tiles, kblocks // 8), B * H * W, Cout), dtype=torch.float32, device=x.device)
        if ws.shape[0] >= 2:
            global launch_count
            launch_count += 1        # + the reduce kernel
    lib, st = _prep(x, w_packed, bias, residual, out)
    _lib.check(lib.mumpy_conv2d_nhwc_bf16(_p(x), ld_in or Cin, _p(w_packed), _p(bias), _p(residual), _p(out), Cout, B, H, W, Cin,
                                          Cout, kh, kw, ph, pw, code(x.dtype), code(out_dtype), act, _p(ws),
                                          0 if ws is None else ws.numel() * 4, st), "mumpy_conv2d_nhwc_bf16")
    return out


def conv2d_nhwc_cout1(x, w_hwi, bias, B, H, W, Cin, kh, kw, ph, pw):
    out = torch.empty((B, H, W, 1), dtype=torch.float32, device=x.device)
    lib, st = _prep(x, w_hwi, bias, out)
    _lib.check(lib.mumpy_conv2d_nhwc_cout1(_p(x), _p(w_hwi), _p(bias), _p(out), B, H, W, Cin, kh, kw, ph, pw, st),
               "mumpy_conv2d_nhwc_cout1")
    return out


def im2col_nhwc(x, B, H, W, Cin, kh, kw, ph, pw, Kpad, ld_in=None):
    out = torch.empty((B * H * W, Kpad), dtype=act_dtype(), device=x.device)
    lib, st = _prep(x, out)
    _lib.check(lib.mumpy_im2col_nhwc(_p(x), ld_in or Cin, _p(out), code(out.dtype), B, H, W, Cin, kh, kw, ph, pw, Kpad, st),
               "mumpy_im2col_nhwc")
    return out


def groupnorm_nhwc(x, gamma, beta, B, HW, C, groups, act, eps=1e-5, out=None, ld_out=None, out_col=0, quad_mean=False):
    """GroupNorm + activation on an NHWC map; quad_mean=True returns the mean of every 4 consecutive activated channels (the
    decoder's DAP, decoder.py:140-143) as (..., C/4) without materialising the normalised map."""
    Co = C // 4 if quad_mean else C
    if out is None:
        out = torch.empty(tuple(x.shape[:-1]) + (Co,), dtype=torch.float32, device=x.device)
    n_ws = _lib.load().mumpy_groupnorm_workspace_floats(B, HW, C, groups)      # the library's own layout (statistics + chunk partials)
    if n_ws <= 0:
        raise _lib.MumpyError("groupnorm_nhwc: unsupported shape B=%d HW=%d C=%d groups=%d" % (B, HW, C, groups))
    stats = torch.empty((n_ws,), dtype=torch.float32, device=x.device)
    lib, st = _prep(x, gamma, beta, stats, out)
    global launch_count
    launch_count += 2            # partial statistics + finalize + apply
    _lib.check(lib.mumpy_groupnorm_nhwc(_p(x), _p(gamma), _p(beta), _p(stats), _p(out), ld_out or Co, out_col, B, HW, C,
                                        groups, eps, act, int(quad_mean), st), "mumpy_groupnorm_nhwc")
    return out


def resample_nhwc(x, B, H, W, C, mode, scale=2, mul=None, add=None, out=None, ld_out=None, out_col=0, out_dtype=torch.float32):
    """out_dtype: torch.float32, or the operand type when the result only feeds a tensor-core convolution (no separate cast pass)."""
    Ho, Wo, Co = H, W, C
    if mode in (RS_UP_ALIGNED, RS_UP_HALFPIX):
        Ho, Wo = H * scale, W * scale
    elif mode == RS_AVGPOOL2:
        Ho, Wo = H // 2, W // 2
    elif mode == RS_PIXEL_SHUFFLE2:
        Ho, Wo, Co = 2 * H, 2 * W, C // 4
    if out is None:
        out = torch.empty((B, Ho, Wo, Co), dtype=out_dtype, device=x.device)
    lib, st = _prep(x, mul, add, out)
    _lib.check(lib.mumpy_resample_nhwc(_p(x), _p(mul), _p(add), _p(out), code(out.dtype), ld_out or Co, out_col, B, H, W, C, mode, scale, st),
               "mumpy_resample_nhwc")
    return out


def mul_add(a, b, c=None, out_dtype=torch.float32):
    out = torch.empty(a.shape, dtype=out_dtype, device=a.device)
    lib, st = _prep(a, b, c, out)
    _lib.check(lib.mumpy_mul_add(_p(a), _p(b), _p(c), _p(out), code(out.dtype), a.numel(), st), "mumpy_mul_add")
    return out


def add(a, b):
    out = torch.empty_like(a)
    lib, st = _prep(a, b, out)
    _lib.check(lib.mumpy_add(_p(a), _p(b), _p(out), a.numel(), st), "mumpy_add")
    return out


def nchw_to_nhwc(x, pool2=False, out=None, ld_out=None, out_col=0):
    B, C, H, W = x.shape
    Ho, Wo = (H // 2, W // 2) if pool2 else (H, W)
    if out is None:
        out = torch.empty((B, Ho, Wo, C), dtype=torch.float32, device=x.device)
    lib, st = _prep(x, out)
    _lib.check(lib.mumpy_nchw_to_nhwc(_p(x), _p(out), ld_out or C, out_col, B, C, H, W, int(pool2), st), "mumpy_nchw_to_nhwc")
    return out


def nhwc_to_nchw(x, B, H, W, C, ld_in=None):
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    lib, st = _prep(x, out)
    _lib.check(lib.mumpy_nhwc_to_nchw(_p(x), ld_in or C, _p(out), B, C, H, W, st), "mumpy_nhwc_to_nchw")
    return out


def channel_group_mean(x, pixels, C, k):
    out = torch.empty((pixels, C // k), dtype=torch.float32, device=x.device)
    lib, st = _prep(x, out)
    _lib.check(lib.mumpy_channel_group_mean(_p(x), _p(out), pixels, C, k, st), "mumpy_channel_group_mean")
    return out


def assemble_clips(frames, clip_frames, mean, std, out=None):
    """frames (n,H,W,3) uint8, clip_frames (B,T) int32 -> normalised clips (B,T,3,H,W) fp32 (ToTensor + Normalize on device)."""
    if frames.dtype != torch.uint8 or clip_frames.dtype != torch.int32 or frames.shape[-1] != 3:
        raise _lib.MumpyError("assemble_clips: frames must be uint8 HWC, clip_frames int32")
    B, T = clip_frames.shape
    H, W = frames.shape[1], frames.shape[2]
    if out is None:
        out = torch.empty((B, T, 3, H, W), dtype=torch.float32, device=frames.device)
    lib, st = _prep(frames, clip_frames, out)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s = (ctypes.c_float * 3)(*[float(v) for v in std])
    _lib.check(lib.mumpy_assemble_clips(_p(frames), _p(clip_frames), _p(out), B, T, H, W, m, s, st), "mumpy_assemble_clips")
    return out


RESIZE_NEAREST, RESIZE_BICUBIC = 0, 3          # PIL's filter codes
_taps_cache = {}


def resize_taps(in_size, out_size, filter=RESIZE_BICUBIC):
    """Per-axis resampling tables exactly as Pillow builds them (host computation inside the library, no GPU needed).
    bicubic -> (bounds int32 (out,2) = [first source index, taps], coefs int32 (out,ksize), ksize); nearest -> (index int32 (out,), None, 1)."""
    lib = _lib.load()
    ks = ctypes.c_int(0)
    if filter == RESIZE_NEAREST:
        idx = torch.empty(out_size, dtype=torch.int32)
        _lib.check(lib.mumpy_resize_taps(int(in_size), int(out_size), int(filter), idx.data_ptr(), None, 0, ctypes.addressof(ks)), "mumpy_resize_taps")
        return idx, None, 1
    bounds = torch.empty((out_size, 2), dtype=torch.int32)
    _lib.check(lib.mumpy_resize_taps(int(in_size), int(out_size), int(filter), bounds.data_ptr(), None, 0, ctypes.addressof(ks)), "mumpy_resize_taps")
    coefs = torch.empty((out_size, ks.value), dtype=torch.int32)
    _lib.check(lib.mumpy_resize_taps(int(in_size), int(out_size), int(filter), bounds.data_ptr(), coefs.data_ptr(), coefs.numel(),
                                     ctypes.addressof(ks)), "mumpy_resize_taps")
    return bounds, coefs, ks.value


def _device_taps(in_size, out_size, filter, device):
    key = (in_size, out_size, filter, str(device))
    if key not in _taps_cache:
        b, c, k = resize_taps(in_size, out_size, filter)
        _taps_cache[key] = (b.to(device), None if c is None else c.to(device), k)
    return _taps_cache[key]


def resize_u8(frames, out_h, out_w, filter=RESIZE_BICUBIC):
    """frames (n,H,W,C) uint8 on the device -> (n,out_h,out_w,C) uint8, bit-identical to PIL.Image.resize((out_w,out_h), filter)
    of every frame (universaldataset.py:68-79)."""
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] > 4 or not frames.is_cuda:
        raise _lib.MumpyError("resize_u8: frames must be a (n,H,W,C<=4) uint8 CUDA tensor")
    frames = frames.contiguous()
    n, H, W, C = frames.shape
    out = torch.empty((n, out_h, out_w, C), dtype=torch.uint8, device=frames.device)
    bh, ch, kh = _device_taps(W, out_w, filter, frames.device)
    bv, cv, kv = _device_taps(H, out_h, filter, frames.device)
    tmp = torch.empty((n, H, out_w, C), dtype=torch.uint8, device=frames.device) if (filter != RESIZE_NEAREST and H != out_h and W != out_w) else None
    if tmp is not None:
        global launch_count
        launch_count += 1            # horizontal + vertical pass
    lib, st = _prep(frames, out)
    _lib.check(lib.mumpy_resize_u8(_p(frames), _p(out), _p(tmp), n, H, W, out_h, out_w, C, int(filter), _p(bh), _p(ch), kh, _p(bv), _p(cv), kv, st),
               "mumpy_resize_u8")
    return out


def mask_counts(logits, gt=None, want_mask=True):
    """logits (B,1,H,W) fp32 -> (mask uint8 (B,H,W) in {0,255}, counts int64 (B,4) = [TP, n_pred, n_gt, n_union])."""
    B = logits.shape[0]
    HW = logits.numel() // B
    mask = torch.empty((B,) + tuple(logits.shape[-2:]), dtype=torch.uint8, device=logits.device) if want_mask else None
    counts = torch.zeros((B, 4), dtype=torch.int64, device=logits.device)
    lib, st = _prep(logits, gt, mask, counts)
    _lib.check(lib.mumpy_mask_counts(_p(logits), _p(gt), _p(mask), _p(counts), B, HW, st), "mumpy_mask_counts")
    return mask, counts


def cast16(x, dtype=None):
    """fp32 -> 16-bit operand (`dtype`, default: the operand type of the current precision mode)."""
    dtype = dtype or act_dtype()
    if dtype == torch.float32:
        raise _lib.MumpyError("cast16 called in fp32 mode")
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    lib, st = _prep(x, out)
    _lib.check(lib.mumpy_cast16(_p(x), _p(out), code(dtype), x.numel(), st), "mumpy_cast16")
    return out


def cast_bf16(x):
    return cast16(x, torch.bfloat16)


def set_attention_tc(enabled: bool):
    """16-bit table-mode window attention: True (default) the tcgen05 / TMEM kernel, False the per-warp mma.sync kernel."""
    _lib.check(_lib.load().mumpy_set_attention_tc(int(bool(enabled))), "mumpy_set_attention_tc")


def set_gemm_pair_mode(mode: int):
    """CTA-pair (cta_group::2) policy of the tensor-core GEMM: 0 never, 1 cost model, 2 whenever legal, 3 cost model for K >= 1024,
    4 (default) cost model except for the 16-bit outputs that have the lean 1-CTA kernel."""
    _lib.check(_lib.load().mumpy_set_gemm_pair_mode(int(mode)), "mumpy_set_gemm_pair_mode")
