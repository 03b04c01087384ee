import sys, torch, time
sys.path.insert(0, 'compress-robust-vqa_b200')
from crvqa import ops
torch.manual_seed(0)
dev = 'cuda'
def rel(a,b): return float((a-b).abs().max()/b.abs().max())
for (M,N,K) in [(128,128,64),(128,128,128),(256,256,256),(640,768,768)]:
    x = torch.randn(M,K,device=dev).bfloat16(); w = (torch.randn(N,K,device=dev)*0.05).bfloat16()
    s = torch.rand(N,K,device=dev); thr = torch.tensor(0.5,device=dev); dy = torch.randn(M,N,device=dev).bfloat16()
    b = torch.randn(N,device=dev)
    for masked in (False, True):
        wm = w.float()*(s>thr).float() if masked else w.float()
        try:
            y = ops.masked_linear_fwd(x,w,s if masked else None,thr,b); torch.cuda.synchronize()
            print('fwd',M,N,K,masked, rel(y, x.float()@wm.t()+b))
        except Exception as e: print('fwd FAIL',M,N,K,masked,e)
        try:
            dx = ops.masked_linear_bwd_dx(dy,w,s if masked else None,thr); torch.cuda.synchronize()
            print('dx ',M,N,K,masked, rel(dx, dy.float()@wm))
        except Exception as e: print('dx FAIL',M,N,K,masked,e)
    try:
        ds = ops.masked_linear_bwd_ds(dy,x,w.float()); torch.cuda.synchronize()
        print('ds ',M,N,K, rel(ds, (dy.float().t()@x.float())*w.float()))
    except Exception as e: print('ds FAIL',M,N,K,e)
