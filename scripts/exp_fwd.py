import sys, ctypes; sys.path.insert(0,'.')
import torch
from efficientvideoclassification_youtube8m_b200 import ops, _lib
from efficientvideoclassification_youtube8m_b200.params import ModelConfig, HLstmParams
from efficientvideoclassification_youtube8m_b200.engine import HLstmEngine
dbg = ctypes.CDLL(_lib.LIB_PATH).evc_debug_set
cfg = ModelConfig(); B=256
p = HLstmParams("model", cfg, "cuda", seed=0)
t = HLstmEngine(p, B, 300, 20, training=True)
x = torch.randn(B,300,1152,device="cuda"); nf = torch.full((B,),300,dtype=torch.int32,device="cuda")
t.forward(x, None, True, nf); torch.cuda.synchronize()
H,D,R1,ell = 1024,1152,t.R1,t.ell; lay=t.l1[0]
def run():
    ops.lstm_seq_fwd(t.x, R1*D, D, p.shadow[p.kernel(0,0)], p.w[p.bias(0,0)], R1, H, ell, t.len_l1, lay.h_all, lay.c_all, lay.gates)
def timeit(n=5):
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n/ell*1e3
for flag in [0,4,24,28,2,1]:
    dbg(flag); print("debug",flag,"us/launch", timeit())
dbg(0)
