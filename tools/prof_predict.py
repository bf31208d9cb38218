"""Small driver for ncu captures: a few device-resident predict steps of the bench workload.
usage: python tools/prof_predict.py [B] [N] [M] [steps] [kernel] [f64|f32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from corenav_gp_b200 import synthetic as syn
from corenav_gp_b200.api import GpContext

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
M = int(sys.argv[3]) if len(sys.argv) > 3 else 600
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
kern = sys.argv[5] if len(sys.argv) > 5 else "rbf+stdperiodic"
prec = sys.argv[6] if len(sys.argv) > 6 else "f64"
ctx = GpContext(0, precision=prec)
x, y = syn.slip_windows(0, B, N)
xs = syn.test_grid(x[0], M)
th = syn.theta_for(kern)
dx, dy, dxs, dth = (torch.from_numpy(a).cuda() for a in (x, y, xs, th))
out = ctx.predict(kern, dth, dx, dy, dxs)   # warm-up (module load, attributes)
torch.cuda.synchronize()
ctx.set_profiling(True)
for _ in range(steps):
    out = ctx.predict(kern, dth, dx, dy, dxs)
torch.cuda.synchronize()
print("fit ms", ctx.profile_read(0), "var ms", ctx.profile_read(1))
