"""Training path of the Moment-Query model on the sm_100a kernels: forward with a tape, hand-written backward.

Every operator of vilco_b200/engine.py has a differentiable twin here (`Tape.*`); the forward records a closure per
operator, `Tape.backward()` replays them in reverse.  Gradients are fp32 token-major tensors; parameter gradients are
accumulated in the packed-weight layout (`tape.G`) and mapped back to the nn.Parameter shapes by `unpack_grads`.
Dropout / stochastic depth are not built yet: the training forward uses the evaluation semantics of those layers
(AffineDropPath = its per-channel scale), which is exactly what the gradient-parity tests compare against.
"""
import ctypes as CT
import math

import os

import torch

from . import backward as BW
from . import engine as E
from . import lib as L
from . import ops
from .ops import _i64, _p, bf16, f32, lo


class V:
    """value + gradient; `p16` caches the 16-bit operand planes of `v`; const = no gradient wanted."""
    __slots__ = ("v", "g", "p16", "const")

    def __init__(self, v=None, p16=None, const=False):
        self.v, self.g, self.p16, self.const = v, None, p16, const

    @property
    def shape(self):
        return self.v.shape if self.v is not None else self.p16.shape[1:]


def _add(a, b):
    return ops.axpby(a, b, 1.0, 1.0)[0]


FUSED_ATTN_BWD = os.environ.get("VILCO_FUSED_ATTN_BWD", "1") == "1"   # 0: materialised S / softmax / dP chain for every layer


class Tape:
    def __init__(self, W, dropout=0.0, droppath=0.0, xl_dropout=0.0, seed=0, sinks=None):
        """dropout / droppath > 0 only when the module is in training mode (nn.Dropout / AffineDropPath semantics)."""
        self.W = W
        self.G = {}       # packed-layout parameter gradients
        # key -> fp32 buffer in the packed layout that IS the parameter's .grad (zeroed by the trainer): kernels accumulate
        # straight into it, no temporary and no add per parameter
        self.sinks = sinks or {}
        self.touch = {}          # sink key -> index (in backward execution order) of the last node that accumulated into it
        self.cur = -1
        self.after_node = None   # optional callback(i, n_nodes) after backward node i (the trainer's bucketed all-reduce)
        self.nodes = []
        self.p_drop, self.p_path, self.p_xl = float(dropout), float(droppath), float(xl_dropout)
        self.seed = int(seed) << 20

    # ---- plumbing ----
    def acc(self, x, g):
        if x is None or g is None or x.const:
            return
        g = g.reshape(x.shape)
        x.g = g.contiguous() if x.g is None else _add(x.g, g.contiguous())

    def accp(self, key, g):
        if g is None:
            return
        self.G[key] = g if key not in self.G else _add(self.G[key], g)

    def planes(self, x, hi=False):
        """operand planes of the activation x; hi: this consumer is one of the split-operand contractions (E.sensitive_key)"""
        want = ops.PLANES_HI if hi else ops.PLANES
        if x.p16 is None or (x.p16.shape[0] < want and x.v is not None):
            y, _ = BW.to_planes(x.v.reshape(-1, x.v.shape[-1]), grad=False, planes=want)
            x.p16 = y.reshape(y.shape[0], *x.v.shape)
        return x.p16

    def sink(self, key):
        """the .grad view to accumulate parameter `key` into (or None), remembering which backward node used it last"""
        if key is None:
            return None
        t = self.sinks.get(key)
        if t is not None:
            self.touch[key] = self.cur
        return t

    def backward(self):
        if ops.BWD_PRECISION == "bf16" and not ops._single:
            with ops.single_plane():
                return self.backward()
        self.n_nodes = len(self.nodes)
        for i, fn in enumerate(reversed(self.nodes)):
            self.cur = i
            fn()
            if self.after_node is not None:
                self.after_node(i, self.n_nodes)
        self.nodes = []

    # ---- operators ----
    def linear(self, x, wkey, bkey=None, rowmul=None, bias=None, out16=False):
        """y = (x w^T + b) * rowmul.  `bias` overrides W[bkey] (XLNet r_w / r_r).  out16: the result only feeds tensor-core
        kernels, so the GEMM epilogue writes the bf16 operand planes directly and no fp32 copy exists."""
        W = self.W
        hi = E.sensitive_key(wkey)
        x16 = self.planes(x, hi)
        b = bias if bias is not None else (W[bkey] if bkey else None)
        if out16:
            y = V(p16=ops.linear(x16, W[wkey], bf16, bias=b, rowmul=rowmul, planes=ops.PLANES_HI if hi else None))
        else:
            y = V(ops.linear(x16, W[wkey], f32, bias=b, rowmul=rowmul))

        def bwd():
            if y.g is None:
                return
            sw, sb = self.sink(wkey), self.sink(bkey) if bkey else None
            dx, dw, db = BW.linear_bwd(y.g, x16, W[wkey], rowmul=rowmul, need_dx=not x.const, need_db=b is not None,
                                       dw_out=sw, db_out=sb)
            self.acc(x, dx)
            if sw is None:
                self.accp(wkey, dw)
            if bkey and sb is None:
                self.accp(bkey, db)
        self.nodes.append(bwd)
        return y

    def conv3(self, x, wkey, bkey=None, rowmul=None):
        W = self.W
        x16 = self.planes(x, E.sensitive_key(wkey))
        y = V(ops.conv3(x16, W[wkey], f32, bias=W[bkey] if bkey else None, rowmul=rowmul))
        cache = W.setdefault("_cache", {})       # derived from the current weights; dropped when they change
        N = W[wkey].shape[2]
        N8 = (N + 7) // 8 * 8
        fkey = (wkey, "flip", N8)
        if fkey not in cache:                   # tap-reversed weights; rows padded to 8 so the dgrad K keeps 16-byte rows
            wf = W[wkey].flip(1)
            if N8 != N:
                wp = torch.zeros(wf.shape[0], 3, N8, wf.shape[3], device=wf.device, dtype=wf.dtype)
                wp[:, :, :N] = wf
                wf = wp
            cache[fkey] = wf.contiguous()
        wflip = cache[fkey]

        def bwd():
            if y.g is None:
                return
            if N8 != N:
                g8 = torch.zeros(*y.g.shape[:-1], N8, device=y.g.device, dtype=f32)
                g8[..., :N] = y.g
                dx, dw, db = BW.conv3_bwd(g8, x16, None, wflip, rowmul=rowmul, need_dx=not x.const)
                dw, db = dw[:, :N].contiguous(), db[:N].contiguous()
            else:
                dx, dw, db = BW.conv3_bwd(y.g, x16, None, wflip, rowmul=rowmul, need_dx=not x.const)
            self.acc(x, dx)
            self.accp(wkey, dw)
            if bkey:
                self.accp(bkey, db)
        self.nodes.append(bwd)
        return y

    def gelu(self, x, want16=True):
        y = V(*ops.ew(1, x.v, out16=True)) if want16 else V(ops.ew(1, x.v))
        self.nodes.append(lambda: self.acc(x, BW.gelu_bwd(y.g, x.v)) if y.g is not None else None)
        return y

    def relu(self, x):
        y = V(ops.ew(2, x.v))
        self.nodes.append(lambda: self.acc(x, ops.ew(3, y.g, y=y.v)) if y.g is not None else None)
        return y

    def ln(self, x, wkey, bkey, eps=1e-5, relu=False, pe=None, rowmul=None, zero_rows=None, keep_rows=None, want16=True,
           hi=False):
        """channel LayerNorm (+ReLU, + pe*rowmul constant, rows flagged in zero_rows forced to 0).  want16: the kernel also
        writes the operand planes the next GEMM reads (hi: for a split-operand contraction)."""
        W = self.W
        np_ = ops.PLANES_HI if hi else None
        if pe is not None:   # constant positional term, no gradient
            y32, y16 = ops.layernorm(x.v, W[wkey], W[bkey], eps, relu=relu, pe=pe, rowmul=rowmul, out32=True, out16=want16,
                                     rows_per_batch=x.v.shape[1], planes=np_)
            yr, _ = ops.layernorm(x.v, W[wkey], W[bkey], eps, relu=relu, out32=True, out16=False)
        else:
            y32, y16 = ops.layernorm(x.v, W[wkey], W[bkey], eps, relu=relu, out32=True, out16=want16, zero_rows=zero_rows,
                                     planes=np_)
            yr = y32
        y = V(y32, p16=y16)

        def bwd():
            if y.g is None:
                return
            g = y.g if keep_rows is None else ops.ew(0, y.g, rowmul=keep_rows)
            sw, sb = self.sink(wkey), self.sink(bkey)
            if (sw is None) != (sb is None):
                sw = sb = None
            dx, dw, db = BW.layernorm_bwd(g, x.v, W[wkey], eps, y_relu=yr if relu else None, dw_out=sw, db_out=sb)
            self.acc(x, dx)
            if sw is None:
                self.accp(wkey, dw)
                self.accp(bkey, db)
        self.nodes.append(bwd)
        return y

    def dwconv_ln3(self, x, pre, mask, stride):
        W = self.W
        names = ("query", "key", "value")
        wc = [W[pre + f"{n}_conv.conv.weight"] for n in names]
        lw = [W[pre + f"{n}_norm.weight"] for n in names]
        lb = [W[pre + f"{n}_norm.bias"] for n in names]
        outs = [V(p16=o) for o in ops.dwconv_ln(x.v, mask, wc, lw, lb, stride)]

        def bwd():
            B, T, Cc = x.v.shape
            dys = [o.g if o.g is not None else torch.zeros(B, T // stride, Cc, device=x.v.device) for o in outs]
            dx, dwc, dlw, dlb = BW.dwconv_ln_bwd(dys, x.v, mask, wc, lw, stride)
            self.acc(x, dx)
            for n, a, b_, c in zip(names, dwc, dlw, dlb):
                self.accp(pre + f"{n}_conv.conv.weight", a)
                self.accp(pre + f"{n}_norm.weight", b_)
                self.accp(pre + f"{n}_norm.bias", c)
        self.nodes.append(bwd)
        return outs

    def attention(self, q, k, v, kmask, H, scale):
        q16, k16, v16 = self.planes(q), self.planes(k), self.planes(v)
        C = q16.shape[-1]
        if E.SELF_ATTN_SP and q16.shape == k16.shape and q16.shape[0] == 1 and v16.shape[0] == 1 and \
                ops.xl_attention_ok(q16, q16.shape[2], C, H):
            r = ops.self_attention(q16, k16, v16, kmask, H, scale, want_lse=FUSED_ATTN_BWD)   # single-pass kernel
            o16, lse = r if FUSED_ATTN_BWD else (r, None)
            o = V(p16=o16)
        else:
            lse = None
            o = V(p16=ops.attention(q16, k16, v16, kmask, H, scale))

        def bwd():
            if o.g is None:
                return
            if lse is not None:      # softmax recompute / backward fused into the GEMM epilogues (backward.attention_bwd_lse)
                dq, dk, dv = BW.attention_bwd_lse(o.g, o.p16, lse, q16, k16, v16, kmask, H, scale)
            else:
                dq, dk, dv = BW.attention_bwd(o.g, q16, k16, v16, kmask, H, scale)
            self.acc(q, dq); self.acc(k, dk); self.acc(v, dv)
        self.nodes.append(bwd)
        return o

    def proj_branch(self, resid, rm, x, wkey, bkey, skey, mask_rows, B, rows_per_sample, p=None, path=True):
        """One fused residual branch: out = resid * rm + scale * dropout(x W^T + b) * mask * drop_path  (training semantics of
        `pool_skip(x)*mask + drop_path_attn(attn(...))` / `out + drop_path_mlp(mlp(...)*mask)`, blocks.py:567-585).
        Forward = GEMM + vilco_resid_branch_fwd; backward = vilco_resid_branch_bwd (d resid, dZ planes, d bias, d scale in one
        pass) + the two gradient GEMMs."""
        W = self.W
        x16 = self.planes(x)
        y = ops.linear(x16, W[wkey], f32)
        b = W[bkey] if bkey else None
        sc = W.get(skey) if skey else None
        path = self.path_rows(B, rows_per_sample, y.device) if path else None
        ymul = mask_rows if path is None else (mask_rows * path if mask_rows is not None else path)
        p = self.p_drop if p is None else p
        seed = 0
        if p > 0:
            self.seed += 1
            seed = self.seed
        N = y.shape[-1]
        rows = y.numel() // N
        outv = torch.empty_like(y)
        L.check(L.lib().vilco_resid_branch_fwd(_p(resid.v), _p(rm), _p(y), _p(b), _p(sc), _p(ymul), _p(outv), _i64(rows), N,
                                               CT.c_float(p), CT.c_uint64(seed), L.stream_ptr()), "vilco_resid_branch_fwd")
        out = V(outv)

        def bwd():
            if out.g is None:
                return
            g = out.g.contiguous()
            dz = ops.empty16(rows, N, device=g.device, grad=True)
            dres = torch.empty_like(g) if rm is not None else None
            sb = self.sink(bkey) if bkey else None
            ss = self.sink(skey) if sc is not None else None
            db = sb if sb is not None else (torch.zeros(N, device=g.device, dtype=f32) if bkey else None)
            ds = ss if ss is not None else (torch.zeros(N, device=g.device, dtype=f32) if sc is not None else None)
            L.check(L.lib().vilco_resid_branch_bwd(_p(g), _p(rm), _p(y), _p(b), _p(sc), _p(ymul), _p(dres), _p(dz), _i64(lo(dz)),
                                                   _p(db), _p(ds), rows, N, CT.c_float(p), CT.c_uint64(seed), L.stream_ptr()),
                    "vilco_resid_branch_bwd")
            self.acc(resid, dres if rm is not None else g)
            if bkey and sb is None:
                self.accp(bkey, db)
            if sc is not None and ss is None:
                self.accp(skey, ds)
            sw = self.sink(wkey)
            dx, dw = BW.linear_bwd16(dz, x16, W[wkey], need_dx=not x.const, dw_out=sw)
            self.acc(x, dx)
            if sw is None:
                self.accp(wkey, dw)
        self.nodes.append(bwd)
        return out

    def resid_scale(self, resid, rowmul, y, skey):
        """out = resid * rowmul[row] + scale[c] * y"""
        s = self.W.get(skey) if skey else None
        out = V(ops.scale_add(resid.v, rowmul, y.v, s))

        def bwd():
            if out.g is None:
                return
            self.acc(resid, out.g if rowmul is None else ops.ew(0, out.g, rowmul=rowmul))
            self.acc(y, out.g if s is None else ops.ew(0, out.g, colmul=s))
            if s is not None:
                sk = self.sink(skey)
                if sk is None:
                    self.accp(skey, BW.colsum(out.g, y=y.v))
                else:
                    BW.colsum(out.g, y=y.v, out=sk)
        self.nodes.append(bwd)
        return out

    def mix(self, a, b, wa=1.0, wb=1.0):
        out = V(ops.axpby(a.v, b.v, wa, wb)[0])

        def bwd():
            if out.g is None:
                return
            self.acc(a, out.g if wa == 1.0 else ops.axpby(out.g, None, wa, 0.0)[0])
            self.acc(b, out.g if wb == 1.0 else ops.axpby(out.g, None, wb, 0.0)[0])
        self.nodes.append(bwd)
        return out

    def dropout(self, x, p=None, want16=True):
        """nn.Dropout(p) in training mode (identity when p == 0)."""
        p = self.p_drop if p is None else p
        if p <= 0.0:
            return x
        self.seed += 1
        seed = self.seed
        y = V(*ops.dropout(x.v, p, seed, out16=True)) if want16 else V(ops.dropout(x.v, p, seed))
        self.nodes.append(lambda: self.acc(x, ops.dropout(y.g.contiguous(), p, seed)) if y.g is not None else None)
        return y

    def path_rows(self, B, rows_per_sample, device):
        """per-row multiplier of stochastic depth (drop_path, blocks.py:640-652): one Bernoulli(1 - p) / (1 - p) per sample."""
        if self.p_path <= 0.0:
            return None
        keep = 1.0 - self.p_path
        m = torch.empty(B, device=device, dtype=f32).bernoulli_(keep) / keep
        return m.repeat_interleave(rows_per_sample).contiguous()

    def rowscale(self, x, rows):
        if rows is None:
            return x
        y = V(ops.ew(0, x.v, rowmul=rows))
        self.nodes.append(lambda: self.acc(x, ops.ew(0, y.g.contiguous(), rowmul=rows)) if y.g is not None else None)
        return y

    def transpose(self, x):
        """(B,T,C) -> (B,C,T)"""
        y = V(ops.unpack(x.v))
        self.nodes.append(lambda: self.acc(x, ops.unpack(y.g)) if y.g is not None else None)
        return y

    def maxpool(self, x):
        y = V(ops.maxpool3s2(x.v))
        self.nodes.append(lambda: self.acc(x, BW.maxpool3s2_bwd(y.g, x.v)) if y.g is not None else None)
        return y

    def channel_attention(self, qkv, H):
        q16 = self.planes(qkv)
        y16, A16 = ops.channel_attention(q16, H, return_A=True)
        y = V(p16=y16)

        def bwd():
            if y.g is not None:
                self.acc(qkv, ops.channel_attention_bwd(y.g.contiguous(), q16, A16, H))
        self.nodes.append(bwd)
        return y


# ----------------------------------------------------------------------------------------------------
# differentiable blocks (same structure as engine.py; citations there)
# ----------------------------------------------------------------------------------------------------
def adapter(tp, pre, ln1):
    """meta_archs.Adapter.layer over the time axis (meta_archs.py:105-148): (B,T,C) -> (B,T/2,C)."""
    xt = tp.transpose(ln1)                                              # (B,C,T)
    h = tp.gelu(tp.linear(xt, pre + "layer.0.weight", pre + "layer.0.bias"))
    return tp.transpose(tp.linear(h, pre + "layer.2.weight", pre + "layer.2.bias"))


def transformer_block(tp, pre, x, mask, H, stride, cross=None, t_c_alpha=0.8, adapter_pre=None):
    W = tp.W
    C = x.shape[-1]
    scale = 1.0 / math.sqrt(C // H)
    ln1 = tp.ln(x, pre + "ln1.weight", pre + "ln1.bias", want16=(stride == 1), hi=True)   # planes only feed the channel qkv
    qc, kc, vc = tp.dwconv_ln3(ln1, pre + "attn.", mask, stride)
    om = mask[:, ::stride].contiguous() if stride > 1 else mask
    omf = om.reshape(-1)
    q = tp.linear(qc, pre + "attn.query.weight", pre + "attn.query.bias", out16=True)
    k = tp.linear(kc, pre + "attn.key.weight", pre + "attn.key.bias", out16=True)
    v = tp.linear(vc, pre + "attn.value.weight", pre + "attn.value.bias", rowmul=omf, out16=True)
    o = tp.attention(q, k, v, om, H, scale)
    B, To = om.shape
    skip = x if stride == 1 else tp.maxpool(x)
    sa = pre + "drop_path_attn.scale" if (pre + "drop_path_attn.scale") in W else None
    sm = pre + "drop_path_mlp.scale" if (pre + "drop_path_mlp.scale") in W else None
    if adapter_pre is not None:          # out = attn(ln1 x) + adapter(ln1 x), the adapter output is not masked nor dropped
        if tp.p_drop > 0:                # proj_drop sits between the projection and the mask (blocks.py:404-405)
            proj = tp.rowscale(tp.dropout(tp.linear(o, pre + "attn.proj.weight", pre + "attn.proj.bias"), want16=False), omf)
        else:
            proj = tp.linear(o, pre + "attn.proj.weight", pre + "attn.proj.bias", rowmul=omf)
        proj = tp.mix(proj, adapter(tp, adapter_pre, ln1))
        h = tp.resid_scale(skip, omf, tp.rowscale(proj, tp.path_rows(B, To, omf.device)), sa)
    else:
        h = tp.proj_branch(skip, omf, o, pre + "attn.proj.weight", pre + "attn.proj.bias", sa, omf, B, To)
    if cross is not None and (pre + "cross_attn.query.weight") in W:
        text, tmask = cross
        hx = tp.ln(h, pre + "ln3.weight", pre + "ln3.bias")
        hy = tp.ln(text, pre + "ln3.weight", pre + "ln3.bias")
        cq = tp.linear(hx, pre + "cross_attn.query.weight", pre + "cross_attn.query.bias", out16=True)
        ck = tp.linear(hy, pre + "cross_attn.key.weight", pre + "cross_attn.key.bias", out16=True)
        cv = tp.linear(hy, pre + "cross_attn.value.weight", pre + "cross_attn.value.bias", rowmul=tmask.reshape(-1), out16=True)
        c = tp.attention(cq, ck, cv, tmask, H, scale)
        h = tp.proj_branch(h, omf, c, pre + "cross_attn.proj.weight", pre + "cross_attn.proj.bias", sa, omf, B, To)
    h2 = tp.ln(h, pre + "ln2.weight", pre + "ln2.bias")
    m1 = tp.dropout(tp.gelu(tp.linear(h2, pre + "mlp.0.weight", pre + "mlp.0.bias"), want16=tp.p_drop <= 0))
    out = tp.proj_branch(h, None, m1, pre + "mlp.3.weight", pre + "mlp.3.bias", sm, omf, B, To)
    if stride == 1:
        cpre = pre + "channel_attn."
        qkv = tp.linear(ln1, cpre + "attn.qkv.weight", out16=True)
        y = tp.channel_attention(qkv, H)
        T_ = ln1.shape[1]        # ChannelBlock: DropPath on both residual branches (blocks.py:448, 462-464)
        x1 = tp.proj_branch(ln1, None, y, cpre + "attn.proj.weight", cpre + "attn.proj.bias", None, None, B, T_, p=0.0)
        n2 = tp.ln(x1, cpre + "norm2.weight", cpre + "norm2.bias", 1e-5)
        out2 = tp.proj_branch(x1, None, tp.gelu(tp.linear(n2, cpre + "mlp.0.weight", cpre + "mlp.0.bias")),
                              cpre + "mlp.2.weight", cpre + "mlp.2.bias", None, None, B, T_, p=0.0)
        out = tp.mix(out, out2, t_c_alpha, 1.0 - t_c_alpha)
    return out, om


def xlnet_layer(tp, pre, x, mask, H, eps=1e-12):
    """XLNetModel.forward with one XLNetLayer on inputs_embeds (modeling_xlnet_x.py:1121-1283, 440-467, 270-332, 482-490),
    including its dropout sites (input :1201, pos_emb :1228, attn_prob :308, attn_out :327, ff :486-488, output :1280) when
    tp.p_xl > 0."""
    W = tp.W
    B, T, C = x.shape
    d = C // H
    pd = tp.p_xl
    scale = 1.0 / math.sqrt(d)
    kq, kk, kv, ko, kr = (pre + "rel_attn." + n for n in "qkvor")
    rw, rr = pre + "rel_attn.r_w_bias", pre + "rel_attn.r_r_bias"
    x = tp.dropout(x, pd)
    qw = tp.linear(x, kq, rw, out16=True)
    qr = tp.linear(x, kq, rr, out16=True)
    k = tp.linear(x, kk, out16=True)
    v = tp.linear(x, kv, out16=True)
    pos16 = E.xlnet_pos_emb(T, C, x.v.device)                  # constant (2T, C)
    if pd > 0:     # the reference drops the (2T, B, C) expanded table, i.e. an independent mask per clip (:1228)
        posb = V(ops.merge16(pos16).unsqueeze(0).expand(B, 2 * T, C).contiguous(), const=True)
        krel1, krel = None, tp.linear(tp.dropout(posb, pd), kr, out16=True)
    else:
        krel1 = tp.linear(V(p16=pos16, const=True), kr)       # (2T, C) fp32, depends on W_r only
        krel = V(krel1.v.unsqueeze(0).expand(B, 2 * T, C).contiguous())
    qw16, qr16, k16, v16, kr16 = (tp.planes(t) for t in (qw, qr, k, v, krel))
    ac = ops.attn_scores(qw16, k16, H, 1.0)
    bd = ops.attn_scores(qr16, kr16, H, 1.0, band=(T, 2 * T))
    P16, P32 = ops.softmax_rows(ac, mask, mode=1, BD=bd, scale=scale, want32=True)
    del ac, bd
    seed = None
    if pd > 0:                                                 # dropout on the attention probabilities
        tp.seed += 1
        seed = tp.seed
        P16 = ops.dropout(P32, pd, seed, out32=False, out16=True)[1]
    vec = V(ops.attn_pv(P16, v16, H, T, out32=True))

    def bwd():
        if vec.g is None:
            return
        gi = ops.ginv()                                        # gradient planes are stored times GRAD_SCALE
        dvec16, _ = BW.to_planes(vec.g.reshape(-1, C))
        dvec16 = dvec16.reshape(dvec16.shape[0], B, T, C)
        dP = ops.attn_scores(dvec16, v16, H, gi)
        if seed is not None:
            ops.dropout(dP, pd, seed, out=dP)
        dS, dS16 = BW.softmax_bwd(dP, scale, P32=P32, want32=True)
        del dP
        dBD16 = ops.empty16(B, H, T, 2 * T, device=dS.device, grad=True)
        L.check(L.lib().vilco_relshift_bwd(_p(dS), None, _p(dBD16), _i64(lo(dBD16)), _i64(B * H), T, L.stream_ptr()),
                "vilco_relshift_bwd")
        del dS
        tp.acc(qw, ops.attn_pv(dS16, k16, H, T, out32=True, alpha=gi))
        tp.acc(k, ops.attn_pv(dS16, qw16, H, T, out32=True, a_trans=True, alpha=gi))        # dS^T qw: dS read as MN-major A
        del dS16
        tp.acc(v, ops.attn_pv(P16, dvec16, H, T, out32=True, a_trans=True, alpha=gi))       # P^T dvec
        tp.acc(qr, ops.attn_pv(dBD16, kr16, H, 2 * T, out32=True, alpha=gi))
        tp.acc(krel, ops.attn_pv(dBD16, qr16, H, T, out32=True, a_trans=True, alpha=gi))    # dBD^T qr -> (B, 2T, C)
        if krel1 is not None:                                  # krel is the batch broadcast of krel1
            g1 = krel.g[0]
            for b in range(1, B):
                g1 = _add(g1.contiguous(), krel.g[b].contiguous())
            tp.acc(krel1, g1)
    tp.nodes.append(bwd)
    a = tp.proj_branch(x, None, vec, ko, None, None, None, B, T, p=pd, path=False)          # dropout(attn_out) + h
    h1 = tp.ln(a, pre + "rel_attn.layer_norm.weight", pre + "rel_attn.layer_norm.bias", eps)
    f = tp.dropout(tp.gelu(tp.linear(h1, pre + "ff.layer_1.weight", pre + "ff.layer_1.bias"), want16=pd <= 0), pd)
    f = tp.proj_branch(h1, None, f, pre + "ff.layer_2.weight", pre + "ff.layer_2.bias", None, None, B, T, p=pd, path=False)
    return tp.dropout(tp.ln(f, pre + "ff.layer_norm.weight", pre + "ff.layer_norm.bias", eps, want16=False), pd, want16=False)


def backbone(tp, cfg, x16, mask, text16, tmask, pe, pets_prefix="pets."):
    """ConvTransformerBackbone.forward (backbones.py:181-289), training batch semantics (padded text participates).
    Returns (feats, masks, text input V or None)."""
    pre = "backbone."
    H = cfg.n_head
    m = mask.reshape(-1)
    x = tp.linear(V(p16=x16, const=True), pre + "proj.0.conv.weight", pre + "proj.0.conv.bias", rowmul=m)
    n_embd = cfg.arch[0]
    for i in range(n_embd):
        c = tp.conv3(x, pre + f"embd.{i}.conv.weight", None, rowmul=mask)
        last = i == n_embd - 1
        x = tp.ln(c, pre + f"embd_norm.{i}.weight", pre + f"embd_norm.{i}.bias", relu=True, pe=pe if last else None,
                  rowmul=m if last else None, want16=not last, hi=True)
    cross, tin = None, None
    if cfg.use_cross_modal and text16 is not None:
        tm = tmask.reshape(-1)
        t = tin = V(p16=text16)
        for i in range(n_embd):
            c = tp.linear(t, pre + f"txt_embd.{i}.conv.weight", None, rowmul=tm)
            t = tp.ln(c, pre + f"txt_embd_norm.{i}.weight", pre + f"txt_embd_norm.{i}.bias", relu=True)
        for i in range(cfg.arch[1]):
            t, _ = transformer_block(tp, pre + f"txt_stem.{i}.", t, tmask, H, 1, t_c_alpha=0.8)
        cross = (t, tmask)
    for i in range(cfg.arch[1]):
        x, _ = transformer_block(tp, pre + f"stem.{i}.", x, mask, H, 1, t_c_alpha=cfg.t_c_alpha)
    feats, masks = [x], [mask]
    if cfg.use_xl and cfg.arch[2] > 0:
        x = xlnet_layer(tp, pre + "xlnet.layer.0.", x, mask, H)
    for i in range(cfg.arch[2]):
        cr = None if i in (1, 2) else cross
        ad = (pets_prefix + f"{list(cfg.adapt_blocks).index(i)}.") if i in cfg.adapt_blocks else None
        x, mask = transformer_block(tp, pre + f"branch.{i}.", x, mask, H, cfg.scale_factor, cross=cr, t_c_alpha=cfg.t_c_alpha,
                                    adapter_pre=ad)
        feats.append(x)
        masks.append(mask)
    return feats, masks, tin


def neck_heads(tp, cfg, feats, masks):
    W = tp.W
    B, C = feats[0].shape[0], cfg.embd_dim
    dev = feats[0].v.device
    pyr = E.Pyramid.cached([f.shape[1] for f in feats], dev)
    P = pyr.P
    lv = [tp.ln(f, f"neck.fpn_norms.{l}.weight", f"neck.fpn_norms.{l}.bias", want16=False) for l, f in enumerate(feats)]
    fpn = V(torch.zeros(B, P, C, device=dev, dtype=f32))
    pmask = torch.zeros(B, P, device=dev, dtype=f32)
    if getattr(pyr, "lvl_of_row", None) is None:
        lvl = torch.full((P,), -1, dtype=torch.long)
        for l in range(len(lv)):
            lvl[pyr.off[l]:pyr.off[l] + pyr.lens[l]] = l
        pyr.lvl_of_row = lvl.to(dev)
        pyr.lvl_onehot = (pyr.lvl_of_row[None, :] == torch.arange(len(lv), device=dev)[:, None]).float()    # (levels, P)
    lvl_of_row = pyr.lvl_of_row
    for l, (f, mk) in enumerate(zip(lv, masks)):
        o, n = pyr.off[l], pyr.lens[l]
        fpn.v[:, o:o + n] = f.v
        pmask[:, o:o + n] = mk

    def cat_bwd():
        if fpn.g is None:
            return
        for l, f in enumerate(lv):
            o, n = pyr.off[l], pyr.lens[l]
            tp.acc(f, fpn.g[:, o:o + n].contiguous())
    tp.nodes.append(cat_bwd)
    zero_rows = pyr.gap_rows.repeat(B)
    keep = (1.0 - pyr.gap_rows.float()).repeat(B).contiguous()
    outs = []
    for head in ("cls_head.", "reg_head."):
        x = fpn
        for i in range(2):
            c = tp.conv3(x, head + f"head.{i}.conv.weight", None, rowmul=pmask)
            x = tp.ln(c, head + f"norm.{i}.weight", head + f"norm.{i}.bias", relu=True, zero_rows=zero_rows, keep_rows=keep)
        if head == "cls_head.":
            outs.append(tp.conv3(x, "cls_head.cls_head.conv.weight", "cls_head.cls_head.conv.bias", rowmul=pmask))
        else:
            z = tp.conv3(x, "reg_head.offset_head.conv.weight", "reg_head.offset_head.conv.bias", rowmul=pmask)
            scales = torch.stack([W[f"reg_head.scale.{l}.scale"].reshape(()) for l in range(len(feats))])
            sv = torch.where(lvl_of_row >= 0, scales[lvl_of_row.clamp(min=0)], torch.zeros((), device=dev)).repeat(B).contiguous()
            s = V(ops.ew(0, z.v, rowmul=sv))

            def sc_bwd(s=s, z=z, sv=sv):
                if s.g is None:
                    return
                tp.acc(z, ops.ew(0, s.g, rowmul=sv))
                r = (s.g * z.v).sum(-1).sum(0)            # (P,) tiny glue reduction for the per-level Scale parameters
                per_level = pyr.lvl_onehot @ r            # (levels,): no boolean indexing (it synchronises)
                for l in range(len(feats)):
                    tp.accp(f"reg_head.scale.{l}.scale", per_level[l:l + 1])
            tp.nodes.append(sc_bwd)
            outs.append(tp.relu(s))
    return outs[0], outs[1], pmask, pyr, lv


# ----------------------------------------------------------------------------------------------------
# packed gradient -> nn.Parameter layout (inverse of engine.pack_weights)
# ----------------------------------------------------------------------------------------------------
def unpack_grad(key, g, param):
    last = key.rsplit(".", 1)[-1]
    if ".rel_attn." in key and last in ("q", "k", "v", "r"):
        return g.t().reshape(param.shape)
    if key.endswith("_conv.conv.weight"):
        return g.t().reshape(param.shape)
    if last == "weight" and param.dim() == 3 and param.shape[2] == 3 and param.shape[0] != 1:
        return g.permute(1, 2, 0).reshape(param.shape)
    return g.reshape(param.shape)
