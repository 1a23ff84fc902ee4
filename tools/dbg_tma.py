import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'oracle')
import numpy as np, torch
import livevideo_b200 as lv
w,h,ns,nf = 352,288,1,2
ctx = lv.Context(w,h,ns)
d = lv.gen_synthetic("camera",1,0,0,ns,nf,w,h)
out = ctx.alloc_outputs(ns*nf)
try:
    ctx.convert_diff_batch(d, ns, nf, out); torch.cuda.synchronize(); print("dbg", os.environ.get("LV_TMA_DEBUG"), "OK")
except Exception as e:
    print("dbg", os.environ.get("LV_TMA_DEBUG"), "FAIL", str(e)[:80])
