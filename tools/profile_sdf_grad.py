import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nefii_b200 import ops
from oracle import mlp
dev = torch.device("cuda:0")
params = mlp.sdf_init(seed=1, bumps=0.3)
net = ops.SdfMlp(device=dev)
net.set_weights([t.to(dev) for t in params.W], [t.to(dev) for t in params.b])
x = torch.rand(131072, 3, device=dev) * 1.8 - 0.9
for _ in range(2):
    net.eval(x, want_feat=True, want_grad=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
net.eval(x, want_feat=True, want_grad=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
