import torch, time
a=torch.randn(8192,8192,device='cuda',dtype=torch.bfloat16); b=torch.randn(8192,8192,device='cuda',dtype=torch.bfloat16)
for _ in range(3): (a@b)
torch.cuda.synchronize()
best=1e9
for _ in range(10):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); c=a@b; e1.record(); torch.cuda.synchronize()
    best=min(best,e0.elapsed_time(e1))
print("bf16 8192^3 best ms", best, "TFLOP/s", 2*8192**3/best/1e9)
t0=time.time(); n=0
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(300): c=a@b
e1.record(); torch.cuda.synchronize()
print("sustained TFLOP/s", 300*2*8192**3/e0.elapsed_time(e1)/1e9)
x=torch.empty(1<<30,dtype=torch.bfloat16,device='cuda'); y=torch.empty_like(x)
best=1e9
for _ in range(10):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
print("copy GB/s", 2*x.numel()*2/best/1e6)
import subprocess; print(subprocess.run(["nvidia-smi","--query-gpu=name,driver_version,vbios_version,power.limit,clocks.max.sm,clocks.max.mem","--format=csv"],capture_output=True,text=True).stdout)
