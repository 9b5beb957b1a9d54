import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, zfp_b200 as zb
from test_gpu_fullsize import device_field
x = device_field((1024,1024,1024), torch.float64)
c = zb.compress(x, accuracy=1e-6)
c = zb.compress(x, reuse=c, accuracy=1e-6)
torch.cuda.synchronize()
print("launches", zb.launch_count())
