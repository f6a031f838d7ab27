"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch fp32/fp64 tensor arithmetic) of the
reference's inference forward (test.py:94-95: Encoder -> Decoder).

Nothing in the product package may import this file.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs use it, and only as the checker / timed baseline.

Parity pin: the reference ships no tests or golden vectors.  This restatement
is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by
oracle/make_golden.py (imports /root/reference unmodified through oracle/ref_harness.py) and committed
as tests/golden/*.pt; tests/test_oracle_golden.py re-checks the restatement against those fixtures.

Every function cites the reference lines it restates (paths relative to /root/reference).
It is deliberately written without vmap, einops, grid_sample, nn.Upsample or nn.Module so that the
index maps the CUDA kernels implement are spelled out here.
"""
import math

import torch

# --------------------------------------------------------------------------------------------
# hyper-parameters: models/factory/modelFactory.py:38-62, models/encoder/multiTemporalViewEncoder.py:676
# --------------------------------------------------------------------------------------------
VIEW_DIMS = ([96, 192, 384, 768], [96, 192, 384, 768], [128, 256, 512, 1024])
VIEW_HEADS = ([3, 6, 12, 24], [3, 6, 12, 24], [4, 8, 16, 32])
VIEW_DEPTHS = ([2, 2, 6, 2], [2, 2, 18, 2], [2, 2, 18, 2])
VIEW_KT = (3, 2, 1)          # temporal kernel = stride of the Conv3d tokenizer
VIEW_T = (1, 1, 3)           # temporal tokens per view (frames stacked vertically on the canvas)
STAGE_DEPTHS = (2, 2, 18, 2)
GLOBAL_DIM, GLOBAL_HEADS, GLOBAL_LAYERS = 768, 12, 12
LN_EPS = 1e-5


def default_config(res=(56, 28, 14, 7), ws=7, img=224):
    return dict(res=tuple(res), ws=ws, img=img)


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps=LN_EPS):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu(x):
    """nn.GELU() default = exact erf form.  erf is evaluated in float64 and rounded once: the result is then independent
    of the host's single-precision erf implementation (a B200 box whose CPU path differed by 3e-4 on a 49x96 tensor was
    observed), and still within 1 ulp of what the reference's ATen kernel returns."""
    xd = x.double()
    return (0.5 * xd * (1.0 + torch.erf(xd / math.sqrt(2.0)))).to(x.dtype)


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def group_norm_nchw(x, groups, w, b, eps=1e-5):
    B, C, H, W = x.shape
    xg = x.reshape(B, groups, (C // groups) * H * W)
    mu = xg.mean(-1, keepdim=True)
    var = ((xg - mu) ** 2).mean(-1, keepdim=True)
    xn = ((xg - mu) / torch.sqrt(var + eps)).reshape(B, C, H, W)
    return xn * w.view(1, C, 1, 1) + b.view(1, C, 1, 1)


def conv2d(x, w, b, pad):
    """Direct NCHW convolution as shifted matmuls (stride 1)."""
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    ph, pw = pad
    xp = torch.zeros(B, Cin, H + 2 * ph, W + 2 * pw, dtype=x.dtype)
    xp[:, :, ph:ph + H, pw:pw + W] = x
    out = torch.zeros(B, Cout, H, W, dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, :, i:i + H, j:j + W]                       # B Cin H W
            out += torch.einsum('bchw,oc->bohw', patch, w[:, :, i, j])
    if b is not None:
        out = out + b.view(1, Cout, 1, 1)
    return out


def avg_pool2(x):
    B, C, H, W = x.shape
    return x.reshape(B, C, H // 2, 2, W // 2, 2).mean(dim=(3, 5))


def pixel_shuffle2(x):
    B, C, H, W = x.shape
    r = 2
    x = x.reshape(B, C // (r * r), r, r, H, W)
    return x.permute(0, 1, 4, 2, 5, 3).reshape(B, C // (r * r), H * r, W * r)


def _bilinear_axis(n_in, n_out, align_corners, dtype):
    """source index/weights of nn.Upsample(mode='bilinear') along one axis."""
    o = torch.arange(n_out, dtype=dtype)
    if align_corners:
        src = o * ((n_in - 1) / (n_out - 1)) if n_out > 1 else torch.zeros_like(o)
    else:
        src = (o + 0.5) * (n_in / n_out) - 0.5
        src = src.clamp(min=0.0)
    i0 = src.floor().long().clamp(max=n_in - 1)
    i1 = (i0 + 1).clamp(max=n_in - 1)
    w1 = src - i0.to(dtype)
    return i0, i1, 1.0 - w1, w1


def upsample_bilinear(x, scale, align_corners):
    B, C, H, W = x.shape
    y0, y1, wy0, wy1 = _bilinear_axis(H, H * scale, align_corners, x.dtype)
    x0, x1, wx0, wx1 = _bilinear_axis(W, W * scale, align_corners, x.dtype)
    rows = x[:, :, y0, :] * wy0.view(1, 1, -1, 1) + x[:, :, y1, :] * wy1.view(1, 1, -1, 1)
    return rows[:, :, :, x0] * wx0.view(1, 1, 1, -1) + rows[:, :, :, x1] * wx1.view(1, 1, 1, -1)


# --------------------------------------------------------------------------------------------
# a1. FAF: models/modules/dct.py:42-49,56-79 ; caller keeps frame 1 only (multiTemporalViewEncoder.py:734)
# --------------------------------------------------------------------------------------------
def dct_matrix(n, dtype=torch.float32):
    i = torch.arange(n, dtype=torch.float64).view(n, 1)
    j = torch.arange(n, dtype=torch.float64).view(1, n)
    m = math.sqrt(2.0 / n) * torch.cos((j + 0.5) * math.pi * i / n)
    m[0, :] = math.sqrt(1.0 / n)
    return m.to(dtype)          # dct.py:59 casts the float64 list to .float()


def band_filters(n, dtype=torch.float32):
    """dct.py:66-68 : (start,end) = (0, n//2.82), (n//2.82, n//2), (n, 2n); keep start <= i+j <= end."""
    s = torch.arange(n).view(n, 1) + torch.arange(n).view(1, n)
    bands = [(0, n // 2.82), (n // 2.82, n // 2), (n * 1, n * 2)]
    return [((s >= lo) & (s <= hi)).to(dtype) for lo, hi in bands]


def faf_middle(x):
    """x (B,3,3,S,S) -> (B,9,S,S): FAF applied to frame 1 (frames are independent in FAF)."""
    S = x.shape[-1]
    D = dct_matrix(S, x.dtype)
    xm = x[:, 1]                                   # (B,3,S,S)
    xf = D @ xm @ D.t()
    ys = [D.t() @ (xf * f) @ D for f in band_filters(S, x.dtype)]
    return torch.cat(ys, dim=1)                    # channel = band*3 + rgb


# --------------------------------------------------------------------------------------------
# a2. tokenizer: multiTemporalViewEncoder.py:574-618
# --------------------------------------------------------------------------------------------
def tokenize(sd, x, pre="base.tokenize."):
    """x (B,3,3,S,S) [b t c h w] -> list of canvases (B, T*H*W, C) with token order (t,h,w)."""
    B, T, Cc, S, _ = x.shape
    Hp = S // 4
    outs = []
    for v in range(3):
        w = sd[pre + "project%d.weight" % (v + 1)]          # (C,3,kt,4,4)
        b = sd[pre + "project%d.bias" % (v + 1)]
        kt = VIEW_KT[v]
        To = (T - kt) // kt + 1                                # Conv3d stride=kernel, no padding
        C = w.shape[0]
        xt = x[:, :To * kt].reshape(B, To, kt, Cc, Hp, 4, Hp, 4)
        # -> (B,To,Hp,Hp, c,kt,4,4) matching weight (C, c, kt, 4, 4)
        patches = xt.permute(0, 1, 4, 6, 3, 2, 5, 7).reshape(B, To * Hp * Hp, Cc * kt * 16)
        tok = patches @ w.reshape(C, -1).t() + b
        tok = layer_norm(tok, sd[pre + "norm%d.weight" % (v + 1)], sd[pre + "norm%d.bias" % (v + 1)])
        outs.append(tok)
    return outs


# --------------------------------------------------------------------------------------------
# a4. window maps: models/modules/swinTransformer.py:54-83
# --------------------------------------------------------------------------------------------
def window_partition(x, ws):
    """(B,TH,W,C) -> (B*nW, ws*ws, C)"""
    B, TH, W, C = x.shape
    x = x.reshape(B, TH // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws * ws, C)


def window_reverse(xw, ws, TH, W):
    C = xw.shape[-1]
    B = xw.shape[0] // ((TH // ws) * (W // ws))
    x = xw.reshape(B, TH // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, TH, W, C)


def relative_position_bias(table, ws):
    """swinTransformer.py:114-123,148-150 -> (nH, ws*ws, ws*ws)"""
    idx = torch.arange(ws * ws)
    r, c = idx // ws, idx % ws
    rel = (r.view(-1, 1) - r.view(1, -1) + ws - 1) * (2 * ws - 1) + (c.view(-1, 1) - c.view(1, -1) + ws - 1)
    return table[rel.reshape(-1)].reshape(ws * ws, ws * ws, -1).permute(2, 0, 1).contiguous()


def shifted_window_mask(TH, W, ws, shift, dtype=torch.float32):
    """swinTransformer.py:233-252 on the stacked T*H x W canvas -> (nW, ws*ws, ws*ws) of 0/-100."""
    def region(n):
        r = torch.zeros(n, dtype=torch.long)
        r[n - ws:n - shift] = 1
        r[n - shift:] = 2
        return r
    img = (region(TH).view(-1, 1) * 3 + region(W).view(1, -1)).to(dtype)
    mw = window_partition(img.view(1, TH, W, 1), ws).squeeze(-1)          # nW, ws*ws
    diff = mw.unsqueeze(1) - mw.unsqueeze(2)
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


# --------------------------------------------------------------------------------------------
# a5. WindowAttention.forward: swinTransformer.py:134-166
# --------------------------------------------------------------------------------------------
def window_attention(sd, pre, xw, nH, ws, mask):
    Bn, N, C = xw.shape
    d = C // nH
    qkv = linear(xw, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"]).reshape(Bn, N, 3, nH, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (d ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1) + relative_position_bias(sd[pre + "relative_position_bias_table"], ws).unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(Bn // nW, nW, nH, N, N) + mask.view(1, nW, 1, N, N)).view(Bn, nH, N, N)
    attn = torch.softmax(attn, dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(Bn, N, C)
    return linear(out, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def mlp(sd, pre, x):
    """Mlp.forward swinTransformer.py:45-51 / FeedForward.forward blocks.py:28-34"""
    return linear(gelu(linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"])), sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def _attn_branch(sd, pre, x, TH, W, nH, ws, shift):
    """LN1 -> roll -> partition -> W-MSA -> reverse -> roll back (swinTransformer.py:265-301)."""
    B, L, C = x.shape
    xn = layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"]).view(B, TH, W, C)
    mask = None
    if shift > 0:
        xn = torch.roll(xn, shifts=(-shift, -shift), dims=(1, 2))
        mask = shifted_window_mask(TH, W, ws, shift, x.dtype)
    aw = window_attention(sd, pre + "attn.", window_partition(xn, ws), nH, ws, mask)
    a = window_reverse(aw, ws, TH, W)
    if shift > 0:
        a = torch.roll(a, shifts=(shift, shift), dims=(1, 2))
    return a.reshape(B, L, C)


# a6. SwinTransformerBlock.forward: swinTransformer.py:259-307
def swin_block(sd, pre, x, TH, W, nH, ws, shift):
    # (swinTransformer.py:217-220: resolution <= window forces shift 0 -- applied by the caller, stages())
    x = x + _attn_branch(sd, pre, x, TH, W, nH, ws, shift)
    return x + mlp(sd, pre + "mlp.", layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"]))


# --------------------------------------------------------------------------------------------
# a9. SwinDAttention.forward: models/modules/deformableAttention.py:324-405 (closed form)
# --------------------------------------------------------------------------------------------
def bilinear_sample_zeros(img, py, px):
    """img (N,C,H,W); py/px (N,P) pixel coordinates -> (N,C,P). F.grid_sample(bilinear, zeros padding)."""
    N, C, H, W = img.shape
    y0 = torch.floor(py)
    x0 = torch.floor(px)
    out = torch.zeros(N, C, py.shape[1], dtype=img.dtype)
    flat = img.reshape(N, C, H * W)
    for dy in (0, 1):
        for dx in (0, 1):
            yy = y0 + dy
            xx = x0 + dx
            wgt = (1 - (py - yy).abs()) * (1 - (px - xx).abs())
            ok = (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
            idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).long()
            g = torch.gather(flat, 2, idx.unsqueeze(1).expand(N, C, -1))
            out = out + g * (wgt * ok.to(img.dtype)).unsqueeze(1)
    return out


def swin_dattention(sd, pre, x1w, x2w, nH, ws=7, groups=3, clip_windows=None, return_attn=False):
    """x1w (N1,P,C) query windows, x2w (N2=r*N1,P,C) key/value windows (already through `pre`).

    out[i] = raw_reshape( proj_out( sum_t A(x1w[q(i,t)], x2w[r*i+t]) ) ),  q(i,t) = (r*i+t) mod N1
    (deformableAttention.py:329-330 repeat + :394-395 '(b t)' sum).  clip_windows=None reproduces the
    reference's batch-global modulo; clip_windows=nW1 applies the B=1 map inside every clip.
    """
    N1, P, C = x1w.shape
    N2 = x2w.shape[0]
    r = N2 // N1
    Cg = C // groups
    d = C // nH
    j = torch.arange(N2)
    if clip_windows is None:
        qidx = j % N1
    else:
        i_out = j // r
        clip = i_out // clip_windows
        qidx = clip * clip_windows + ((i_out % clip_windows) * r + j % r) % clip_windows

    q_tok = linear(x1w, sd[pre + "proj_q.weight"].reshape(C, C), sd[pre + "proj_q.bias"])     # (N1,P,C)
    # ---- offset network on q, per group: dw 5x5 (pad 2) -> LN(Cg) -> GELU -> 1x1 (Cg->2)  (:253-258,334-340)
    qmap = q_tok.reshape(N1, ws, ws, groups, Cg)
    dw_w = sd[pre + "conv_offset.0.weight"].reshape(Cg, 5, 5)
    dw_b = sd[pre + "conv_offset.0.bias"]
    padded = torch.zeros(N1, ws + 4, ws + 4, groups, Cg, dtype=x1w.dtype)
    padded[:, 2:2 + ws, 2:2 + ws] = qmap
    acc = torch.zeros_like(qmap)
    for a in range(5):
        for b in range(5):
            acc = acc + padded[:, a:a + ws, b:b + ws] * dw_w[:, a, b]
    acc = acc + dw_b
    acc = gelu(layer_norm(acc, sd[pre + "conv_offset.1.norm.weight"], sd[pre + "conv_offset.1.norm.bias"]))
    off = acc @ sd[pre + "conv_offset.3.weight"].reshape(2, Cg).t()                          # (N1,ws,ws,g,2) (y,x)
    off = torch.tanh(off) * (1.0 / ws) * 2.0
    ref = (torch.arange(ws, dtype=x1w.dtype) + 0.5) / ws * 2.0 - 1.0                          # :313-319
    pos_y = off[..., 0] + ref.view(1, ws, 1, 1)
    pos_x = off[..., 1] + ref.view(1, 1, ws, 1)
    # align_corners=True un-normalisation (:353-356)
    pix_y = ((pos_y + 1.0) * 0.5 * (ws - 1)).permute(0, 3, 1, 2).reshape(N1, groups, P)
    pix_x = ((pos_x + 1.0) * 0.5 * (ws - 1)).permute(0, 3, 1, 2).reshape(N1, groups, P)

    # ---- sample kv windows with the offsets of their paired query window
    x2img = x2w.reshape(N2, ws, ws, groups, Cg).permute(0, 3, 4, 1, 2).reshape(N2 * groups, Cg, ws, ws)
    samp = bilinear_sample_zeros(x2img, pix_y[qidx].reshape(N2 * groups, P), pix_x[qidx].reshape(N2 * groups, P))
    samp = samp.reshape(N2, C, P).transpose(1, 2)                                             # (N2,P,C)
    k = linear(samp, sd[pre + "proj_k.weight"].reshape(C, C), sd[pre + "proj_k.bias"])
    v = linear(samp, sd[pre + "proj_v.weight"].reshape(C, C), sd[pre + "proj_v.bias"])
    qh = q_tok[qidx].reshape(N2, P, nH, d).permute(0, 2, 1, 3)
    kh = k.reshape(N2, P, nH, d).permute(0, 2, 1, 3)
    vh = v.reshape(N2, P, nH, d).permute(0, 2, 1, 3)
    attn = torch.softmax((qh @ kh.transpose(-2, -1)) * (d ** -0.5), dim=-1)                   # :364,390
    o = (attn @ vh).permute(0, 2, 1, 3).reshape(N2, P, C)
    o = o.reshape(N1, r, P, C).sum(1)                                                         # :394-395
    y = linear(o, sd[pre + "proj_out.weight"].reshape(C, C), sd[pre + "proj_out.bias"])       # (N1,P,C) token-major
    # :403 `.reshape(B, H*W, C)` of a (B,C,H,W) tensor: reinterpret channel-major memory, no transpose
    out = y.transpose(1, 2).contiguous().reshape(N1, P, C)
    if return_attn:                                                                           # :399 (N2*nH,P,P) -> (N1, r*nH, P, P)
        return out, attn.reshape(N1, r * nH, P, P)
    return out


# a8. CrossSwinBlock.forward: multiTemporalViewEncoder.py:228-291 (+ CVAModule :134-139), SURVEY A4
def cross_swin_block(sd, pre, x1, x2, TH1, TH2, W, nH, ws, last_view, per_clip_pairing=False):
    B, L1, C1 = x1.shape
    out = _attn_branch(sd, pre, x1, TH1, W, nH, ws, 0)
    h = x1 + out
    if not last_view:
        hw = window_partition(h.view(B, TH1, W, C1), ws)
        x2w = window_partition(x2.view(B, TH2, W, x2.shape[-1]), ws)
        x2w = linear(x2w, sd[pre + "pre.weight"], sd[pre + "pre.bias"])
        nW1 = (TH1 // ws) * (W // ws)
        y = swin_dattention(sd, pre + "cva.crossattn.", hw, x2w, nH, ws, 3,
                            clip_windows=nW1 if per_clip_pairing else None)
        # CVAModule returns x1w + y; '(b n) ws c -> b (n ws) c' is window-major, no window_reverse (:284-286)
        h = h + (hw + y).reshape(B, L1, C1)
    return h + mlp(sd, pre + "mlp.", layer_norm(h, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])), out


# a10. PatchMerging.forward: swinTransformer.py:344-367
def patch_merging(sd, pre, x, TH, W):
    B, L, C = x.shape
    x = x.view(B, TH, W, C)
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).reshape(B, L // 4, 4 * C)
    x = layer_norm(x, sd[pre + "norm.weight"], sd[pre + "norm.bias"])
    return linear(x, sd[pre + "reduction.weight"])


# a11. CreateStages / MultiViewBasicLayer / Cross+OriginalThreeViewSwinBlock: multiTemporalViewEncoder.py:294-350,390-450,489-571
def stages(sd, toks, cfg, pre="base.layers.layers.", per_clip_pairing=False):
    ws = cfg["ws"]
    xs = list(toks)
    outs = []
    for s in range(4):
        H = cfg["res"][s]
        TH = [VIEW_T[v] * H for v in range(3)]
        nH = [VIEW_HEADS[v][s] for v in range(3)]
        eff_ws = min(ws, H)
        for lyr in range(STAGE_DEPTHS[s]):
            bp = pre + "%d.blocks.%d." % (s, lyr)
            if lyr == 0:
                xs[2], out2 = cross_swin_block(sd, bp + "block3.", xs[2], xs[2], TH[2], TH[2], H, nH[2], eff_ws, True)
                xs[1], out1 = cross_swin_block(sd, bp + "block2.", xs[1], out2, TH[1], TH[2], H, nH[1], eff_ws, False, per_clip_pairing)
                xs[0], _ = cross_swin_block(sd, bp + "block1.", xs[0], out1, TH[0], TH[1], H, nH[0], eff_ws, False, per_clip_pairing)
            else:
                shift = 0 if (lyr % 2 == 0 or H <= ws) else ws // 2
                for v in range(3):
                    if lyr < VIEW_DEPTHS[v][s]:                     # else nn.Identity (:415)
                        xs[v] = swin_block(sd, bp + "block%d." % (v + 1), xs[v], TH[v], H, nH[v], eff_ws, shift)
        outs.append([t.unsqueeze(1) for t in xs])                   # (B,1,L,C) as the vmap'd reference returns
        if s < 3:
            for v in range(3):
                xs[v] = patch_merging(sd, pre + "%d.downsample.downsample%d." % (s, v + 1), xs[v], TH[v], H)
    return xs, outs


# a13. Block / Attention / FeedForward: models/modules/blocks.py:37-92 on (B*n, 3, 768)  
def vit_block(sd, pre, x, heads, return_attention=False):
    Bn, N, C = x.shape
    d = C // heads
    xn = layer_norm(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
    qkv = linear(xn, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]).reshape(Bn, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    attn = torch.softmax((qkv[0] @ qkv[1].transpose(-2, -1)) * (d ** -0.5), dim=-1)
    if return_attention:                                                                      # blocks.py:88-89
        return attn
    y = (attn @ qkv[2]).transpose(1, 2).reshape(Bn, N, C)
    x = x + linear(y, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    return x + mlp(sd, pre + "mlp.", layer_norm(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"]))


def encoder_forward(sd, x, cfg=None, per_clip_pairing=False):
    """Encoder.forward (encoder.py:11-18) o ThreeViewSwinTransformer.forward (multiTemporalViewEncoder.py:732-746).

    Returns final_x (B,2304,n,n), view_x [4 stages][3 views] of (B,1,L,C), ffinfo (B,9,S,S).
    """
    cfg = cfg or default_config()
    B = x.shape[0]
    ffinfo = faf_middle(x)
    toks = tokenize(sd, x)
    xs, view_x = stages(sd, toks, cfg, per_clip_pairing=per_clip_pairing)
    n = cfg["res"][3] ** 2
    # a12. merge_views_along_channel_axis (:710-718) + globalembedding (:740); rows ordered (b, n, t)
    v3 = xs[2].reshape(B, 3, n, -1)
    merged = torch.cat([xs[0].unsqueeze(1).expand(-1, 3, -1, -1), xs[1].unsqueeze(1).expand(-1, 3, -1, -1), v3], dim=-1)
    g = linear(merged, sd["base.globalembedding.weight"], sd["base.globalembedding.bias"])      # (B,3,n,768)
    g = g.permute(0, 2, 1, 3).reshape(B * n, 3, GLOBAL_DIM)
    for i in range(GLOBAL_LAYERS):
        g = vit_block(sd, "base.globalblocks.blocks.%d." % i, g, GLOBAL_HEADS)
    final = g.reshape(B, n, 3 * GLOBAL_DIM)                                                   # :745 cat over t
    side = cfg["res"][3]
    final_x = final.reshape(B, side, side, 3 * GLOBAL_DIM).permute(0, 3, 1, 2).contiguous()     # encoder.py:16-17
    return final_x, view_x, ffinfo


# --------------------------------------------------------------------------------------------
# a15-a19. Decoder.forward: models/decoder/decoder.py:183-225
# --------------------------------------------------------------------------------------------
def _gcm(sd, pre, x):      # _GlobalConvModule.forward decoder.py:33-39
    l = conv2d(conv2d(x, sd[pre + "conv_l1.weight"], sd[pre + "conv_l1.bias"], (3, 0)), sd[pre + "conv_l2.weight"], sd[pre + "conv_l2.bias"], (0, 3))
    r = conv2d(conv2d(x, sd[pre + "conv_r1.weight"], sd[pre + "conv_r1.bias"], (0, 3)), sd[pre + "conv_r2.weight"], sd[pre + "conv_r2.bias"], (3, 0))
    return l + r


def _seb(sd, pre, x1, x2):  # SEB.forward decoder.py:12-14 ; nn.Upsample default align_corners=False
    return x1 * upsample_bilinear(conv2d(x2, sd[pre + "conv.weight"], sd[pre + "conv.bias"], (1, 1)), 2, False)


def _dec_stage(sd, pre, x, groups=8):   # decoder.py:67-95: Conv3x3 -> GN -> ReLU -> Up x2 (align_corners=True)
    y = conv2d(x, sd[pre + "0.weight"], sd[pre + "0.bias"], (1, 1))
    y = torch.relu(group_norm_nchw(y, groups, sd[pre + "1.weight"], sd[pre + "1.bias"]))
    return upsample_bilinear(y, 2, True)


def _freq_stage(sd, pre, x, groups):    # decoder.py:147-181: AvgPool2 -> Conv3x3 -> GN -> Sigmoid
    y = conv2d(avg_pool2(x), sd[pre + "1.weight"], sd[pre + "1.bias"], (1, 1))
    return torch.sigmoid(group_norm_nchw(y, groups, sd[pre + "2.weight"], sd[pre + "2.bias"]))


def decoder_forward(sd, final_x, view_x, ffinfo, shape=(56, 28, 14, 7)):
    B = final_x.shape[0]
    rgb = []
    for s in range(4):
        h = shape[s]
        # merge_views_along_channel_axis (decoder.py:43-53) + Conv3d k=s=(3,1,1) (:98-120)
        parts = []
        for v in range(3):
            t = view_x[s][v]
            t = t.reshape(B, VIEW_T[v], -1, t.shape[-1])
            parts.append(t.expand(-1, 3, -1, -1) if VIEW_T[v] == 1 else t)
        merged = torch.cat(parts, dim=-1)                                                      # (B,3,hw,Cin)
        w = sd["rgb_decoder_%d.0.weight" % (s + 1)]                                            # (256,Cin,3,1,1)
        y = torch.einsum('btnc,oct->bno', merged, w[:, :, :, 0, 0]) + sd["rgb_decoder_%d.0.bias" % (s + 1)]
        y = y.permute(0, 2, 1).reshape(B, -1, h, h)
        rgb.append(torch.relu(group_norm_nchw(y, 16, sd["rgb_decoder_%d.1.weight" % (s + 1)], sd["rgb_decoder_%d.1.bias" % (s + 1)])))
    rgb1, rgb2, rgb3, rgb4 = rgb
    freq0 = _freq_stage(sd, "decoder_frequency_0.", ffinfo, 8)
    freq1 = _freq_stage(sd, "decoder_frequency_1.", freq0, 8)
    freq2 = _freq_stage(sd, "decoder_frequency_2.", freq1, 8)
    freq3 = _freq_stage(sd, "decoder_frequency_3.", freq2, 4)
    freq4 = _freq_stage(sd, "decoder_frequency_4.", freq3, 8)

    gcn0 = _gcm(sd, "gcm1.", torch.cat([rgb4, final_x], dim=1))
    out1 = pixel_shuffle2(gcn0 * freq4)
    seb1 = _seb(sd, "seb1.", rgb3, rgb4)
    gcn1 = _gcm(sd, "gcm2.", seb1)
    seb2 = _seb(sd, "seb2.", rgb2, torch.cat([rgb3, upsample_bilinear(rgb4, 2, False)], dim=1))
    gcn2 = _gcm(sd, "gcm3.", seb2)
    seb3 = _seb(sd, "seb3.", rgb1, torch.cat([rgb2, upsample_bilinear(rgb3, 2, False), upsample_bilinear(rgb4, 4, False)], dim=1))
    gcn3 = _gcm(sd, "gcm4.", seb3)

    x = _dec_stage(sd, "decoder_2.", gcn1 * freq3 + out1)
    x = _dec_stage(sd, "decoder_3.", x + gcn2 * freq2)
    x = _dec_stage(sd, "decoder_4.", x + gcn3 * freq1)
    x = _dec_stage(sd, "decoder_5.", x * freq0)
    x_feats = avg_pool2(pixel_shuffle2(x))                                                      # DAP :140-143
    mask = conv2d(x_feats, sd["final_out.weight"], sd["final_out.bias"], (1, 1))
    return mask, x_feats


def forward(enc_sd, dec_sd, x, cfg=None, per_clip_pairing=False):
    """test.py:94-95 : logits (B,1,S,S), x_feats (B,32,S,S)."""
    cfg = cfg or default_config()
    final_x, view_x, ffinfo = encoder_forward(enc_sd, x, cfg, per_clip_pairing)
    return decoder_forward(dec_sd, final_x, view_x, ffinfo, cfg["res"])


# --------------------------------------------------------------------------------------------
# a20 + measure.py:46-91 : mask and per-clip metric
# --------------------------------------------------------------------------------------------
def assemble_clips(frames_u8, seq_lengths, length_clip=3, mean=(0.4776, 0.479, 0.4465), std=(0.230, 0.2085, 0.2324)):
    """One clip per frame, neighbours clamped inside the sequence (universaldataloader.py:45-48), each frame through
    ToTensor (uint8 HWC -> float CHW / 255) and Normalize ((x - mean) / std per channel) (test.py:22-25).
    frames_u8 (n,H,W,3) uint8 -> (n, length_clip, 3, H, W) fp32."""
    k = int(length_clip / 2)
    mean_t = torch.tensor(mean, dtype=torch.float32).view(3, 1, 1)
    std_t = torch.tensor(std, dtype=torch.float32).view(3, 1, 1)
    clips, base = [], 0
    for n in seq_lengths:
        for idx in range(n):
            ids = [max(0, min(n - 1, i)) for i in range(idx - k, idx + k + 1)]
            fr = [((frames_u8[base + j].permute(2, 0, 1).to(torch.float32).div(255)) - mean_t) / std_t for j in ids]
            clips.append(torch.stack(fr, 0))
        base += n
    return torch.stack(clips, 0)


def threshold_mask(logits):
    """test.py:100-106: sigmoid(x) > 0.5  <=>  x > 0 ; uint8 {0,255}."""
    return (logits > 0).to(torch.uint8) * 255


def clip_counts(pred_mask, gt_mask):
    """integer counts [TP, n_pred, n_gt, n_union] per clip; masks (B,H,W) bool."""
    p = pred_mask.bool().flatten(1)
    g = gt_mask.bool().flatten(1)
    return torch.stack([(p & g).sum(1), p.sum(1), g.sum(1), (p | g).sum(1)], dim=1).to(torch.int64)


def f1_iou_from_counts(counts, n_pixels):
    """measure.py:77-91 and :46-62, float64 like numpy. counts (N,4) int64 -> (f1, iou) each (N,)"""
    c = counts.to(torch.float64)
    tp, n_pred, n_gt, n_union = c[:, 0], c[:, 1], c[:, 2], c[:, 3]
    recall = tp / (n_gt + n_pixels * 1e-6)          # np.sum(gt_mask + 1e-6): +1e-6 per pixel (:86)
    precision = tp / (n_pred + 1e-6)
    f1 = 2 * (precision * recall) / (precision + recall + 1e-6)
    iou = (tp + 1e-5) / (n_union + 1e-5)
    return f1, iou
