# fp64 / int8 / bf16 library peaks on this box via torch (cuBLAS); test infrastructure only.
import torch, time, json
def bench(fn, n=5):
    fn(); torch.cuda.synchronize()
    best=1e9
    for _ in range(n):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
out={}
for n in (4096,8192):
    a=torch.randn(n,n,dtype=torch.float64,device='cuda'); b=torch.randn(n,n,dtype=torch.float64,device='cuda')
    ms=bench(lambda: torch.matmul(a,b)); out[f'dgemm_{n}_tflops']=2*n**3/ms*1e-9
    print(n,'dgemm ms',ms,'TF',2*n**3/ms*1e-9)
a=torch.randn(8192,8192,dtype=torch.float64,device='cuda'); a=a@a.T+8192*torch.eye(8192,dtype=torch.float64,device='cuda')
for n in (4608,):
    m=a[:n,:n].contiguous()
    ms=bench(lambda: torch.linalg.cholesky(m),3); print('cusolver potrf',n,ms,'ms'); out['potrf_4608_ms']=ms
    ms=bench(lambda: torch.linalg.ldl_factor(m),2); print('cusolver sytrf',n,ms,'ms'); out['sytrf_4608_ms']=ms
    ms=bench(lambda: torch.linalg.eigvalsh(m),1); print('cusolver syevd',n,ms,'ms'); out['syevd_4608_ms']=ms
try:
    n=8192
    ai=torch.randint(-128,127,(n,n),dtype=torch.int8,device='cuda'); bi=torch.randint(-128,127,(n,n),dtype=torch.int8,device='cuda')
    ms=bench(lambda: torch._int_mm(ai,bi)); print('int8 mm ms',ms,'TOPS',2*n**3/ms*1e-9); out['int8_tops']=2*n**3/ms*1e-9
except Exception as e: print('int8 fail',e)
x=torch.empty(1<<28,dtype=torch.float64,device='cuda'); y=torch.empty_like(x)
ms=bench(lambda: y.copy_(x)); print('copy GB/s', 2*x.numel()*8/ms*1e-6); out['copy_gbs']=2*x.numel()*8/ms*1e-6
json.dump(out,open('gpurun_out/probe_torch.json','w'),indent=1)
