import torch, time, sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import conformer_oracle as O
from _util import build_encoder
cfg = O.conformer_cfg("M")
enc = build_encoder(cfg, 0, compute_dtype=torch.bfloat16)
x = torch.randn(64, 998, 80, device="cuda")
mask = torch.ones(64, 1, 998, dtype=torch.bool, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - s) / n * 1e3
with torch.no_grad():
    print("embed (current bf16 autocast):", t(lambda: enc.embed(x, mask)))
    conv, out = enc.embed.conv, enc.embed.out
    def v_fp32():
        y = conv(x.unsqueeze(1)); b, c, tt, f = y.shape
        return out(y.transpose(1, 2).contiguous().view(b, tt, c * f))
    print("embed fp32/tf32:", t(v_fp32))
    convcl = torch.nn.Sequential(*[m for m in conv]).to(memory_format=torch.channels_last)
    def v_cl():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = convcl(x.unsqueeze(1).contiguous(memory_format=torch.channels_last))
            b, c, tt, f = y.shape
            return out(y.permute(0, 2, 1, 3).reshape(b, tt, c * f))
    print("embed bf16 channels_last:", t(v_cl))
    torch.backends.cudnn.benchmark = True
    print("embed bf16 autocast + cudnn.benchmark:", t(lambda: enc.embed(x, mask)))
    print("embed bf16 channels_last + benchmark:", t(v_cl))
    xh = torch.randn(64, 998, 80).pin_memory()
    print("H2D 20MB:", t(lambda: xh.to("cuda", non_blocking=True)))
    o = torch.randn(64, 248, 256, device="cuda")
    print("D2H 16MB:", t(lambda: o.cpu()))
