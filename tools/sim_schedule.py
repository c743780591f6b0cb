"""Scheduling model of a frame from the oracle's per-pixel cost map (analysis only; test infrastructure).
usage: python tools/sim_schedule.py ggs120:refl_transp 480 270
Greedy list scheduling of 8x4-pixel blocks over 4,736 warps per GPU (148 SMs x 32 resident warps), in queue order and
longest-first, against the ideal sum/P; block cost = slowest lane (SIMT upper bound) or mean lane (lower bound)."""
import sys, time, heapq, numpy as np
sys.path.insert(0,'/root/repo')
from tests import fixtures as fx, oracle_lib as ol
import ctypes as C
from ntracer_b200 import _capi
name,var=sys.argv[1].split(':'); w,h=int(sys.argv[2]),int(sys.argv[3])
sc,g=fx.load(name); sc=fx.variant(sc,g,var)
# fast per-pixel costs: call the oracle's window renderer directly per pixel
d,keep=_capi.make_desc(sc)
o=np.ascontiguousarray(sc['cam_origin'],np.float32); a=np.ascontiguousarray(sc['cam_axes'],np.float32)
rgb=np.zeros((h,w,3),np.float32); mask=np.zeros((h,w),np.uint8)
lib=ol.lib(); cnt=_capi.Counters()
cost=np.zeros((h,w),np.float64)
t=time.time()
p=lambda x: x.ctypes.data_as(C.c_void_p)
for y in range(h):
    for x in range(w):
        lib.oracle_render_float_window(C.byref(d),p(o),p(a),w,h,x,y,x+1,y+1,p(rgb),p(mask),C.byref(cnt))
        cost[y,x]=cnt.simplex_tests+0.3*cnt.node_steps
print('cost map',round(time.time()-t,1),'s; mean',cost.mean(),'max',cost.max())
np.save('/root/repo/gpurun_out/costmap_%s_%dx%d.npy'%(name,w,h),cost)
# blocks of 8x4, tile-major order (32x32 tiles)
bw,bh=8,4
blocks=[]
for ty in range(0,h,32):
    for tx in range(0,w,32):
        for by in range(ty,min(ty+32,h),bh):
            for bx in range(tx,min(tx+32,w),bw):
                blk=cost[by:by+bh,bx:bx+bw]
                blocks.append((blk.max(), blk.sum()/32))   # SIMT: a warp pays for its slowest lane (upper bound) / mean (lower bound)
blocks=np.array(blocks)
def makespan(costs,P):
    heap=[0.0]*P
    for c in costs:
        t=heapq.heappop(heap); heapq.heappush(heap,t+c)
    return max(heap)
scale=(3840*2160)/(w*h)   # replicate the block population to the 4K block count
for label,col in (('warp pays max lane',0),('warp pays mean lane',1)):
    c=np.tile(blocks[:,col],int(round(scale)))
    for gpus in (1,8):
        P=4736*gpus
        ideal=c.sum()/P
        rm=makespan(c,P); lpt=makespan(np.sort(c)[::-1],P)
        print('%s | %d GPU(s): ideal %.0f  queue order %.0f (x%.2f)  LPT %.0f (x%.2f)  largest block %.0f'%(label,gpus,ideal,rm,rm/ideal,lpt,lpt/ideal,c.max()))
