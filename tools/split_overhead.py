"""One middle z-slab (ghost layer below and above) on one GPU, no neighbours attached: what does splitting a
step into boundary part + interior part cost against the single launch?   python tools/split_overhead.py [edge] [planes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voxelyze_b200 import capi, slab

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 512
planes = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lib = capi.load_product()
r = slab.SlabRunner(lib, edge, edge, 3 * planes, 1, 3, peer=False)
sim = r.sim
sim.set_stream(torch.cuda.current_stream().cuda_stream)
dt = sim.recommended_dt()

def timed(fn, n=40):
    fn(5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("voxels", sim.n_voxels, "dt", dt)
print("vx_step       (one launch per step)        %.3f ms/step" % timed(lambda n: sim.step(dt, n)))
print("vx_slab_step  (boundary + interior parts)  %.3f ms/step" % timed(lambda n: sim.slab_step(dt, n)))
def parts(n, which):
    sim.step_begin(dt)
    for _ in range(n):
        if which == "all":
            sim.step_enqueue(0)
        else:
            sim.step_enqueue(1); sim.step_enqueue(2)
    sim.step_end()
print("step_enqueue(ALL)                          %.3f ms/step" % timed(lambda n: parts(n, "all")))
print("step_enqueue(BOUNDARY)+(INTERIOR)          %.3f ms/step" % timed(lambda n: parts(n, "split")))
