import torch
dev='cuda'
def t(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)*1e3/n
torch.set_float32_matmul_precision("medium")
N=256
for M in (8448, 7424):
  for K in (602, 604, 608, 640):
    X=torch.randn(M,K,device=dev); W=torch.randn(N,K,device=dev); dY=torch.randn(M,N,device=dev)
    print(M, K, "fwd %.1f us" % t(lambda: X@W.t()), " dW %.1f us" % t(lambda: dY.t()@X))
M,K=8448,608
X=torch.randn(M,K,device=dev); dY=torch.randn(M,N,device=dev)
for S in (4,8,16):
    f=lambda: torch.bmm(dY.view(S,M//S,N).transpose(1,2), X.view(S,M//S,K)).sum(0)
    print("K=608 split-K", S, "%.1f us" % t(f))
