set -x
python - <<'P' > gpurun_out/slab_check.txt 2>&1
import os, numpy as np, torch
import sys; sys.path.insert(0,'.')
from tests import fixtures as fx
from ntracer_b200 import _capi
from ntracer_b200.backend import DeviceScene
for name,(w,h) in [('cell120',(1920,1080)),('box4',(640,480)),('solids6',(1000,777)),('cell120',(333,129))]:
    sc,g=fx.load(name)
    fmt=_capi.make_image_format(w,h,_capi.RGB8,pitch=w*3+32)
    os.environ['NTR_NO_SLABS']='1'
    with DeviceScene(sc) as ds:
        a=ds.render(fmt, torch.full((fmt.pitch*h,),7,dtype=torch.uint8).pin_memory().numpy()).copy(); ca=ds.counters()
    del os.environ['NTR_NO_SLABS']
    with DeviceScene(sc) as ds:
        b=ds.render(fmt, torch.full((fmt.pitch*h,),7,dtype=torch.uint8).pin_memory().numpy()).copy(); cb=ds.counters()
        c=ds.render(fmt, np.full(fmt.pitch*h,7,np.uint8)).copy()     # pageable: plain path
    print(name,w,h,'slabs == plain:',np.array_equal(a,b),np.array_equal(a,c), ca['shadow_rays'],cb['shadow_rays'], ca['primary_rays'], cb['primary_rays'])
P
cat gpurun_out/slab_check.txt
for c in c2 c1 c3; do
NTR_NO_SLABS=1 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --stream-frames 0 > gpurun_out/slab0_$c.json 2>gpurun_out/slab0_$c.err
python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --stream-frames 0 > gpurun_out/slab1_$c.json 2>gpurun_out/slab1_$c.err
done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/gpu_tests.txt
cat gpurun_out/gpu_tests.txt
