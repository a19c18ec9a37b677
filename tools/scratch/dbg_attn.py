import sys, torch
sys.path.insert(0, '/root/repo')
from toc3d_b200 import lib
lib.load()
DEV='cuda'
def bf16_round(x): return x.bfloat16().float()
for seq in (144, 180, 192, 256):
    g = torch.Generator().manual_seed(seq + 1000)
    nW, heads = 150, 6
    C = heads * 64
    rows = [seq, 1, 33, min(seq, 65), seq, min(seq, 129), 9, seq, min(seq, 128), min(seq, 17), seq, min(seq, 161)]
    qr = torch.tensor([rows[i % len(rows)] for i in range(nW)], dtype=torch.int32)
    qkv = bf16_round(torch.randn(nW, seq, 3 * C, generator=g))
    ref_in = qkv.clone()
    vb = torch.randn(C, generator=g) * 0.5
    for w in range(nW):
        ref_in[w, int(qr[w]):, C:2 * C] = 0.0
        ref_in[w, int(qr[w]):, 2 * C:] = vb
        qkv[w, int(qr[w]):, C:] = 7.0
    q, k, v = ref_in.to(DEV).reshape(nW, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-1, -2)).softmax(-1) @ v).transpose(1, 2).reshape(nW, seq, C).cpu()
    out = torch.full((nW * seq, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    lib.window_attention(qkv.reshape(nW * seq, 3 * C).to(DEV).bfloat16(), out, nW, seq, heads, q_rows=qr.to(DEV), kv_rows=qr.to(DEV), pad_v=vb.to(DEV))
    got = out.float().cpu().view(nW, seq, C)
    bad = []
    for w in range(nW):
        r = int(qr[w])
        e = (got[w, :r] - ref[w, :r]).abs()
        if not torch.isfinite(got[w,:r]).all() or e.max() > 3e-2:
            # which heads / rows
            eh = e.view(r, heads, 64).amax(-1)
            bad.append((w, r, float(e.max()), [int(x) for x in (eh > 3e-2).any(0).nonzero().flatten()], [int(x) for x in (eh > 3e-2).any(1).nonzero().flatten()[:6]]))
    print("seq", seq, "bad windows", len(bad))
    for b in bad[:12]: print("  ", b)
    if bad:
        w, r = bad[0][0], bad[0][1]
        print("   got", got[w, 0, :8].tolist()); print("   ref", ref[w, 0, :8].tolist()); print("   vb ", vb[:8].tolist())
