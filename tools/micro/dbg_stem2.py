import ctypes as C, torch, sys
sys.path.insert(0, '/root/repo')
from fce_yolo_b200 import _lib as L
l = L.load(check_device=True)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
B,H,W,C1 = 2,64,64,128; C0=64
g = torch.Generator().manual_seed(1)
x = torch.randint(0,256,(B,H,W,3),generator=g,dtype=torch.uint8)
w0 = torch.randn(C0,3,3,3,generator=g)*0.3; b0 = torch.randn(C0,generator=g)*0.2
w1 = torch.randn(C1,C0,3,3,generator=g)/(9*C0)**0.5; b1 = torch.randn(C1,generator=g)*0.1
wk = torch.zeros(C0,32); wk[:,:27] = w0.permute(0,2,3,1).reshape(C0,27)/255.0
wk_d,b0_d = wk.bfloat16().cuda(), b0.cuda()
w1_d,b1_d = w1.permute(0,2,3,1).contiguous().bfloat16().cuda(), b1.cuda()
xd = x.cuda()
for act in (0,1):
    y1 = torch.zeros(B,H//4,W//4,C1,dtype=torch.bfloat16,device='cuda'); y2 = y1.clone()
    d = L.Stem2Desc(B=B,H=H,W=W,C0=C0,C1=C1,out_pitch=C1,out_off=0,act0=act,act1=act)
    L.check(l.fce_stem2_conv(C.byref(d),p(xd),p(wk_d),p(b0_d),p(w1_d),p(b1_d),p(y1),st),"s2")
    mid = torch.empty(B,H//2,W//2,C0,dtype=torch.bfloat16,device='cuda')
    ds = L.StemDesc(B=B,H=H,W=W,Cout=C0,out_pitch=C0,out_off=0,act=act,in_dtype=L.U8,in_layout=L.NHWC)
    L.check(l.fce_stem_conv(C.byref(ds),p(xd),p(wk_d),p(b0_d),p(mid),st),"s")
    for impl in (2,3,4,1):
        dc = L.ConvDesc(B=B,H=H//2,W=W//2,Cin=C0,Cout=C1,in_pitch=C0,in_off=0,out_pitch=C1,out_off=0,res_pitch=0,res_off=0,k=3,stride=2,act=act,in_dtype=L.BF16,w_dtype=L.BF16,out_dtype=L.BF16,in_layout=L.NHWC,in_scale=1.0,impl=impl)
        r = l.fce_conv2d(C.byref(dc),p(mid),p(w1_d),p(b1_d),p(None),p(y2),st)
        torch.cuda.synchronize()
        a,b = y1.float(), y2.float()
        diff = (a-b).abs()
        ulp = torch.maximum(a.abs(),b.abs())*2.0**-7+1e-30
        bad = (diff>ulp)
        print("act",act,"impl",impl,"rc",r,"ndiff",int((diff>0).sum()),"of",diff.numel(),"bad",int(bad.sum()),"maxdiff",diff.max().item())
        if impl==2 and bad.any():
            idx = bad.nonzero()[:8]
            for i in idx.tolist():
                print("   ", i, a[tuple(i)].item(), b[tuple(i)].item())
