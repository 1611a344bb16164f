import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, logps as ologps
from open_o3_video_b200 import _lib, logprob
T,H,V = 1000,512,5000
hidden, weight, targets = synth.head_inputs(T,H,V,seed=T+V,planted=True)
h,w,t = hidden.cuda().bfloat16(), weight.cuda().bfloat16(), targets.cuda()
torch.backends.cuda.matmul.allow_tf32=False
zref = (h.float()@w.float().T)
for cta in (1,2):
  for groups in (1,3,0):
    _lib.set_tunable("cta_pair",cta); _lib.set_tunable("fwd_groups",groups)
    z = torch.zeros(T,V,dtype=torch.bfloat16,device="cuda")
    st = logprob.lmhead_stats(h,w,t,0,z)
    rows=[0,676, 5]
    print("cta",cta,"groups",groups)
    for r in rows:
        zr = zref[r]
        print("  row",r,"tgt",int(t[r]),"M,s,z=",[float(st[i,r]) for i in range(3)],"ref max %.7f argmax %d ztgt %.7f lse %.7f"%(float(zr.max()),int(zr.argmax()),float(zr[t[r]]),float(torch.logsumexp(zr,0))),
              "zstore[tgt]", float(z[r,t[r]]), "top2", [float(x) for x in zr.topk(2).values])
