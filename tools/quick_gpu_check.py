import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import zfp_b200 as zb
from oracle.oracle import Port
from helpers import analytic_field
P=Port()
for dt,shape in [(np.float64,(64,64,64)),(np.float32,(64,64,64)),(np.float64,(30,33,35)),(np.int32,(40,40)),(np.float64,(100,)),(np.float64,(8,8,8,8))]:
    a=analytic_field(shape,dt); x=torch.from_numpy(a).cuda()
    for mode in ({"rate":8},{"rate":5.3},{"precision":20},{"accuracy":1e-4},{"reversible":True}):
        if np.dtype(dt).kind!='f' and 'accuracy' in mode: continue
        try:
            c=zb.compress(x,**mode); got=c.to_numpy(); want=P.compress(a,**mode)
            ok=got.tobytes()==want.tobytes()
            back=zb.decompress(c).cpu().numpy(); ok2=back.tobytes()==P.decompress(want,a.shape,a.dtype,**mode).tobytes()
            print(np.dtype(dt).name,shape,mode,got.nbytes,want.nbytes,ok,ok2)
        except Exception as e:
            print(np.dtype(dt).name,shape,mode,"EXC",e)
# timing 512^3 fp64 rate 8
a=analytic_field((512,512,512),np.float64); x=torch.from_numpy(a).cuda()
for rate in (4,8,16):
    c=zb.compress(x,rate=rate)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): c=zb.compress(x,out=c.words,rate=rate)
    e1.record(); torch.cuda.synchronize(); t=e0.elapsed_time(e1)/5
    y=zb.decompress(c); torch.cuda.synchronize(); e0.record()
    for _ in range(5): zb.decompress(c,out=y)
    e1.record(); torch.cuda.synchronize(); t2=e0.elapsed_time(e1)/5
    print("512^3 fp64 rate",rate,"compress %.3f ms %.1f GB/s; decompress %.3f ms %.1f GB/s"%(t,a.nbytes/t/1e6,t2,a.nbytes/t2/1e6), (y.cpu().numpy()==P.decompress(P.compress(a[:64],rate=rate),(64,512,512),np.float64,rate=rate)).all() if False else "")
