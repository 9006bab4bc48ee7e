import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import gpu_util as U
def bench(M,N,K,out_f32=False,res=False,act=0,reps=20):
    a=torch.randn(M,K,device='cuda').half(); w=(torch.randn(N,K,device='cuda')*0.05).half()
    bias=torch.randn(N,device='cuda')
    out=torch.empty(M,N,device='cuda',dtype=torch.float32 if out_f32 else torch.float16)
    r = out if res else None
    for _ in range(3): U.gemm_16(a,w,shift=bias,act=act,residual=r,out_f32=out_f32,out=out)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): U.gemm_16(a,w,shift=bias,act=act,residual=r,out_f32=out_f32,out=out)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/reps
    return ms, 2*M*N*K/ms/1e9
M=31744
for name,(N,K,f32,res,act) in dict(qkv=(1536,512,False,False,0), proj=(512,512,True,True,0), fc1=(2048,512,False,False,2), fc2=(512,2048,True,True,0), big=(4096,4096,False,False,0)).items():
    ms,tf=bench(M,N,K,f32,res,act)
    print(f"dbg={os.environ.get('HVIT_DBG','0')} 1cta={os.environ.get('HVIT_IGEMM_1CTA','0')} {name:5s} M={M} N={N} K={K}: {ms*1e3:8.1f} us  {tf:7.1f} TF/s", flush=True)
