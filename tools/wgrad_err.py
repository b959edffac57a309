import sys, os
R=os.environ.get("GRAFT_REPO_ROOT","/root/repo")
sys.path.insert(0, R)
import torch, x2gnn_b200
from x2gnn_b200 import _lib
L=_lib.lib()
for rows in (50000, 200000, 811834):
  for N in (128, 42):
    g=torch.Generator(device="cuda").manual_seed(rows+N)
    Y=torch.randn(rows,128,device="cuda",generator=g); X=torch.randn(rows,N,device="cuda",generator=g)
    _lib.require_cuda(Y, what="t")
    dW=torch.empty(128,N,device="cuda"); db=torch.empty(128,device="cuda")
    ws=_lib.workspace(L.x2_tc_wgrad_workspace_bytes(rows,N),"cuda")
    _lib.check(L.x2_tc_wgrad(_lib.ptr(Y),Y.stride(0),_lib.ptr(X),X.stride(0),rows,N,_lib.ptr(dW),dW.stride(0),_lib.ptr(db),_lib.ptr(ws),ws.numel(),_lib.stream()),"w")
    ref=(Y.double().t()@X.double())
    f32=(Y.t()@X)   # cuBLAS fp32 (may use tf32? default False for matmul)
    e=float((dW.double()-ref).abs().max()/ref.abs().max()); e32=float((f32.double()-ref).abs().max()/ref.abs().max())
    sgn=float(((dW.double()-ref)*torch.sign(ref)).mean()/ref.abs().mean())
    print(rows,N,"tc_wgrad err %.2e"%e,"cublas fp32 err %.2e"%e32,"signed mean rel bias %.2e"%sgn)
