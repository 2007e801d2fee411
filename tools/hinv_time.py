import sys, os, time
sys.path[:0]=["/root/repo","/root/repo/ai-edge-quantizer_b200"]
import torch
from aeq_b200 import device
dev=torch.device("cuda:0")
for k in (4096, 11008):
  x=torch.randn(16384,k,device=dev)
  h=device.xtx(x,2.0/8); del x
  for _ in range(2): device.hessian_inverse(h,0.01)
  torch.cuda.synchronize()
  e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
  e0.record(); device.hessian_inverse(h,0.01); e1.record(); torch.cuda.synchronize()
  print(os.environ.get("AEQB_CHOL_TWO_LEVEL_MIN_K","default"), k, round(e0.elapsed_time(e1),2),"ms", flush=True)
