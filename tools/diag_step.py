"""Dev: per-step CUDA-event and host-enqueue times of the public fwd+bwd call, enqueued back to back."""
import math, os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import vlp_b200  # noqa
from vlp_b200 import functional as VF
dev = torch.device("cuda:0")
n, d = 32768, 512
def make(seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    a = torch.randn(n, d, generator=g, device=dev); c = torch.randn(n, d, generator=g, device=dev)
    c = 0.35 * a + math.sqrt(1 - 0.35 ** 2) * c
    return F.normalize(a).to(torch.bfloat16), F.normalize(c).to(torch.bfloat16)
sets = [make(42), make(43)]
ls = torch.tensor([math.log(1 / 0.07)], device=dev, requires_grad=True)
def step(I, T):
    a = I.detach().requires_grad_(True); b = T.detach().requires_grad_(True); ls.grad = None
    loss, _, _ = VF.fused_clip_loss_from_embeddings(a, b, ls)
    loss.backward()
    return loss
def run(tag, nvml):
    stop = [False]
    def poll():
        import pynvml
        pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not stop[0]:
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h); time.sleep(0.01)
    th = None
    if nvml:
        th = threading.Thread(target=poll, daemon=True); th.start()
    for w in range(5):
        step(*sets[w % 2])
    torch.cuda.synchronize()
    K = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    host = []
    ev[0].record()
    for k in range(K):
        t0 = time.perf_counter()
        step(*sets[k % 2])
        ev[k + 1].record()
        host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    stop[0] = True
    print(tag, "gpu ms:", " ".join(f"{ev[k].elapsed_time(ev[k+1]):.1f}" for k in range(K)))
    print(tag, "host ms:", " ".join(f"{h:.1f}" for h in host), flush=True)
run("plain", False)
run("nvml ", True)
os.environ["X"] = "1"
VF.SINGLE_SWEEP = False
run("2pass", False)
