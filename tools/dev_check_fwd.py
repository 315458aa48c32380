"""Developer check (GPU): forward LSE kernel vs the CPU oracle."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vlp_b200
from vlp_b200 import _lib
from oracle import clip_oracle as O

lib = _lib.load()
dev = torch.device("cuda:0")
print("version", lib.vlpclip_version(), "sms", lib.vlpclip_sm_count())

def run(n_rows, n_cols, d, ls, shift=0, seed=1):
    n = max(n_rows, n_cols)
    I, T = O.make_embeddings(n, d, rho=0.35, seed=seed)
    X = I[:n_rows].contiguous(); Y = T[:n_cols].contiguous()
    s = min(math.exp(ls), 100.0)
    S = (X.double() @ Y.double().T) * s
    ref_lse = torch.logsumexp(S, dim=1)
    xb = X.to(dev).to(torch.bfloat16); yb = Y.to(dev).to(torch.bfloat16)
    row_m = torch.empty(n_rows, device=dev); row_l = torch.empty(n_rows, device=dev)
    diag = torch.full((n_rows,), float("nan"), device=dev)
    ws_bytes = lib.vlpclip_lse_workspace_bytes(n_rows, n_cols, d)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.vlpclip_lse_fwd(xb.data_ptr(), d, yb.data_ptr(), d, n_rows, n_cols, d, s, shift,
                             row_m.data_ptr(), row_l.data_ptr(), diag.data_ptr(), ws.data_ptr(), ws_bytes, st)
    _lib.check(rc, "lse_fwd")
    torch.cuda.synchronize()
    lse = (row_m.double() + torch.log2(row_l.double())) * math.log(2.0)
    err = (lse.cpu() - ref_lse).abs().max().item()
    # diag
    idx = torch.arange(n_rows) - shift
    ok = (idx >= 0) & (idx < n_cols)
    dref = S[torch.arange(n_rows)[ok], idx[ok]]
    derr = (diag.cpu().double()[ok] - dref).abs().max().item() if ok.any() else 0.0
    print(f"n_rows={n_rows} n_cols={n_cols} d={d} s={s:.2f} shift={shift}: max|lse err|={err:.3e} max|diag err|={derr:.3e}")
    return err, derr

bad = 0
for (nr, nc, d, ls, sh) in [(128, 128, 64, 2.6593, 0), (256, 256, 512, 2.6593, 0), (256, 256, 512, 5.0, 0),
                            (100, 130, 40, 2.6593, 0), (384, 1000, 128, 3.9, 0), (1024, 4096, 512, 2.6593, -1024),
                            (4096, 4096, 512, 2.6593, 0), (4096, 512, 256, 4.0, 512), (8192, 8192, 512, 2.6593, 0)]:
    e, de = run(nr, nc, d, ls, sh)
    if not (e < 2e-4 and de < 2e-4): bad += 1
print("FWD CHECK", "PASS" if bad == 0 else f"FAIL ({bad})")

# timing at the headline size
for n in (8192, 32768):
    d = 512
    xb = torch.nn.functional.normalize(torch.randn(n, d, device=dev)).to(torch.bfloat16)
    yb = torch.nn.functional.normalize(torch.randn(n, d, device=dev)).to(torch.bfloat16)
    row_m = torch.empty(n, device=dev); row_l = torch.empty(n, device=dev); diag = torch.empty(n, device=dev)
    ws_bytes = lib.vlpclip_lse_workspace_bytes(n, n, d); ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def call():
        rc = lib.vlpclip_lse_fwd(xb.data_ptr(), d, yb.data_ptr(), d, n, n, d, 14.29, 0, row_m.data_ptr(), row_l.data_ptr(), diag.data_ptr(), ws.data_ptr(), ws_bytes, st)
        assert rc == 0
    for _ in range(3): call()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"lse_fwd n={n} d={d}: {ms:.3f} ms  -> {2*n*n*d/ms/1e9:.1f} TFLOP/s")
