import torch, time
dev='cuda'
def t(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)*1e3/n
M,K,N = 8448, 602, 256
X=torch.randn(M,K,device=dev); W=torch.randn(N,K,device=dev); dY=torch.randn(M,N,device=dev)
for prec in ("highest","high","medium"):
    torch.set_float32_matmul_precision(prec)
    print(prec, "fwd X@W^T %.1f us" % t(lambda: X@W.t()), " dW dY^T@X %.1f us" % t(lambda: dY.t()@X))
torch.set_float32_matmul_precision("medium")
def splitk(S):
    Xs=X.view(S,M//S,K); Ys=dY.view(S,M//S,N)
    return torch.bmm(Ys.transpose(1,2), Xs).sum(0)
for S in (4,8,16,32):
    print("split-K", S, "%.1f us" % t(lambda: splitk(S)), float((splitk(S)-dY.t()@X).abs().max()/ (dY.t()@X).abs().max()))
Xb=X.bfloat16(); Wb=W.bfloat16(); dYb=dY.bfloat16()
print("bf16 fwd %.1f us" % t(lambda: Xb@Wb.t()), " dW %.1f us" % t(lambda: dYb.t()@Xb))
try:
    torch.backends.cuda.preferred_blas_library("cublaslt")
    print("cublaslt fwd %.1f us" % t(lambda: X@W.t()), " dW %.1f us" % t(lambda: dY.t()@X))
except Exception as e: print("lt err", e)
# other shapes in the step
for (m,k,n) in [(8448,256,256),(4608,256,256),(4608,256,41),(1536,256,41)]:
    A=torch.randn(m,k,device=dev); B=torch.randn(n,k,device=dev); G=torch.randn(m,n,device=dev)
    print((m,k,n), "fwd %.1f" % t(lambda: A@B.t()), "dX %.1f" % t(lambda: G@B), "dW %.1f" % t(lambda: G.t()@A))
