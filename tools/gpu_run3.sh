#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s --durations=8 > gpurun_out/tests_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -25 gpurun_out/tests_gpu.log
python - <<'PY'
import torch, time, sys
sys.path.insert(0,'.')
import pyvbmp_b200 as V
from pyvbmp_b200 import _lib
dev=torch.device('cuda:0')
# d = 128 timing, K = 256, N = 1Mi
N,K,d=1<<20,256,128
g=torch.Generator(device=dev).manual_seed(1)
mu=1.0*torch.randn(K,d,generator=g,device=dev)
X=mu[torch.randint(K,(N,),generator=g,device=dev)]+torch.randn(N,d,generator=g,device=dev)
torch.manual_seed(0); m=V.GaussianMixtureModel(K,d).to(dev); m.dist.mu=X[:K].clone()
for _ in range(3): m.update(X,1)
torch.cuda.synchronize(); _lib.PROFILE={}
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
for _ in range(5): m.update(X,1)
b.record(); b.synchronize(); prof,_lib.PROFILE=_lib.PROFILE,None
print('GMM d=128 K=256 N=1Mi: %.2f ms/iter'%(a.elapsed_time(b)/5), {k: round(sum(x.elapsed_time(y) for x,y in v)/5,3) for k,v in prof.items()})
fl=4*d*d*N*K/(a.elapsed_time(b)/5/1e3)/1e12; print('algorithmic TFLOP/s', fl)
# isotropic GMM timing d=64 K=256 N=4Mi
N,K,d=1<<22,256,64
X=torch.randn(N,d,generator=g,device=dev)*1.5
torch.manual_seed(0); m=V.GaussianMixtureModel(K,d,isotropic=True).to(dev); m.dist.mu=X[:K].clone()
for _ in range(3): m.update(X,1)
torch.cuda.synchronize(); _lib.PROFILE={}
a.record()
for _ in range(5): m.update(X,1)
b.record(); b.synchronize(); prof,_lib.PROFILE=_lib.PROFILE,None
print('GMM isotropic d=64 K=256 N=4Mi: %.2f ms/iter'%(a.elapsed_time(b)/5), {k: round(sum(x.elapsed_time(y) for x,y in v)/5,3) for k,v in prof.items()})
PY
