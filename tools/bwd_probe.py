"""GPU probe: backward primitives (vilco_b200/backward.py) vs torch autograd in fp32."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vilco_b200 import backward as BW, ops
dev = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
bad = 0
torch.manual_seed(0)

def rep(name, got, ref, tol=2e-4):
    global bad
    err = ((got.double() - ref.double()).abs().max() / (ref.double().abs().max() + 1e-12)).item()
    ok = err < tol and torch.isfinite(got).all().item()
    bad += 0 if ok else 1
    print(("OK " if ok else "BAD"), f"{name:40s} rel err {err:.2e}", flush=True)

# ---- linear
for (R, N, K) in [(256, 128, 192), (2048, 1024, 1024), (114, 256, 96)]:
    x = torch.randn(R, K, device=dev, requires_grad=True); w = (torch.randn(N, K, device=dev) * 0.1).requires_grad_()
    b = torch.randn(N, device=dev, requires_grad=True); rm = (torch.rand(R, device=dev) > 0.3).float()
    y = (F.linear(x, w, b)) * rm[:, None]
    dy = torch.randn_like(y)
    y.backward(dy)
    dx, dw, db = BW.linear_bwd(dy, ops.split16(x.detach()), ops.split16(w.detach()), rowmul=rm)
    rep(f"linear dx {R}x{N}x{K}", dx, x.grad); rep(f"linear dw {R}x{N}x{K}", dw, w.grad); rep(f"linear db {R}x{N}x{K}", db, b.grad)
# ---- conv3
for (B, T, N, K) in [(2, 128, 64, 96), (2, 1024, 256, 256)]:
    x = torch.randn(B, T, K, device=dev, requires_grad=True); w = (torch.randn(N, K, 3, device=dev) * 0.1).requires_grad_()
    rm = (torch.rand(B, T, device=dev) > 0.3).float()
    y = F.conv1d(x.transpose(1, 2), w, padding=1).transpose(1, 2) * rm[:, :, None]
    dy = torch.randn_like(y)
    y.backward(dy)
    w3 = w.detach().permute(2, 0, 1).contiguous()
    dx, dw, db = BW.conv3_bwd(dy, ops.split16(x.detach()), ops.split16(w3), ops.split16(w3.flip(0).contiguous()), rowmul=rm)
    rep(f"conv3 dx {B}x{T}x{N}x{K}", dx, x.grad); rep(f"conv3 dw {B}x{T}x{N}x{K}", dw.permute(1, 2, 0), w.grad)
# ---- layernorm (+relu, +add)
for relu in (False, True):
    R, Cc = 300, 256
    x = torch.randn(R, Cc, device=dev, requires_grad=True); a = torch.randn(R, Cc, device=dev, requires_grad=True)
    w = torch.randn(Cc, device=dev, requires_grad=True); b = torch.randn(Cc, device=dev, requires_grad=True)
    y = F.layer_norm(x + a, (Cc,), w, b, 1e-5)
    if relu: y = F.relu(y)
    dy = torch.randn_like(y); y.backward(dy)
    dx, dw, db = BW.layernorm_bwd(dy, x.detach(), w.detach(), 1e-5, add=a.detach(), y_relu=y.detach() if relu else None)
    rep(f"layernorm dx relu={relu}", dx, x.grad); rep(f"layernorm dw relu={relu}", dw, w.grad); rep(f"layernorm db relu={relu}", db, b.grad)
# ---- gelu, maxpool
x = torch.randn(4, 64, 128, device=dev, requires_grad=True)
y = F.gelu(x); dy = torch.randn_like(y); y.backward(dy)
rep("gelu dx", BW.gelu_bwd(dy, x.detach()), x.grad)
x = torch.randn(2, 64, 128, device=dev, requires_grad=True)
y = F.max_pool1d(x.transpose(1, 2), 3, 2, 1).transpose(1, 2); dy = torch.randn_like(y); y.backward(dy)
rep("maxpool dx", BW.maxpool3s2_bwd(dy, x.detach()), x.grad)
# ---- dwconv + LN (q,k,v)
for stride in (1, 2):
    B, T, Cc = 2, 64, 256
    x = torch.randn(B, T, Cc, device=dev, requires_grad=True)
    mask = (torch.arange(T, device=dev)[None, :] < torch.tensor([T, 41], device=dev)[:, None]).float().contiguous()
    ws = [torch.randn(Cc, 1, 3, device=dev, requires_grad=True) for _ in range(3)]
    lw = [torch.randn(Cc, device=dev, requires_grad=True) for _ in range(3)]; lb = [torch.randn(Cc, device=dev, requires_grad=True) for _ in range(3)]
    om = mask[:, ::stride]
    outs = []
    for i in range(3):
        c = F.conv1d(x.transpose(1, 2), ws[i], stride=stride, padding=1, groups=Cc).transpose(1, 2) * om[:, :, None]
        outs.append(F.layer_norm(c, (Cc,), lw[i], lb[i], 1e-5))
    dys = [torch.randn_like(o) for o in outs]
    torch.autograd.backward(outs, dys)
    wpk = [w.detach()[:, 0, :].t().contiguous() for w in ws]
    dx, dwc, dlw, dlb = BW.dwconv_ln_bwd(dys, x.detach(), mask, wpk, [w.detach() for w in lw], stride)
    rep(f"dwconv_ln dx s{stride}", dx, x.grad)
    for i in range(3):
        rep(f"dwconv_ln dwconv{i} s{stride}", dwc[i].t().unsqueeze(1), ws[i].grad); rep(f"dwconv_ln dlnw{i} s{stride}", dlw[i], lw[i].grad)
# ---- attention (global / cross)
for (B, H, Tq, Tk, valid) in [(2, 2, 128, 128, [128, 77]), (2, 4, 256, 57, [57, 30]), (2, 16, 1024, 1024, [1024, 611])]:
    Cc = H * 64
    q = torch.randn(B, Tq, Cc, device=dev, requires_grad=True); k = torch.randn(B, Tk, Cc, device=dev, requires_grad=True)
    v = torch.randn(B, Tk, Cc, device=dev, requires_grad=True)
    km = (torch.arange(Tk, device=dev)[None, :] < torch.tensor(valid, device=dev)[:, None]).float().contiguous()
    qh, kh, vh = (t.view(B, -1, H, 64).permute(0, 2, 1, 3) for t in (q, k, v))
    att = ((qh * 0.125) @ kh.transpose(-1, -2)).masked_fill(km[:, None, None, :] == 0, float("-inf")).softmax(-1)
    o = (att @ vh).permute(0, 2, 1, 3).reshape(B, Tq, Cc)
    dO = torch.randn_like(o); o.backward(dO)
    dq, dk, dv = BW.attention_bwd(dO, ops.split16(q.detach()), ops.split16(k.detach()), ops.split16(v.detach()), km, H, 0.125)
    rep(f"attention dq {B,H,Tq,Tk}", dq, q.grad); rep(f"attention dk {B,H,Tq,Tk}", dk, k.grad); rep(f"attention dv {B,H,Tq,Tk}", dv, v.grad)
print("bad", bad)
sys.exit(1 if bad else 0)
