set -x
python - <<'P' > gpurun_out/zc_check.txt 2>&1
import os, numpy as np, torch
import sys; sys.path.insert(0,'.')
from tests import fixtures as fx
from ntracer_b200 import _capi
from ntracer_b200.backend import DeviceScene
for name in ['cell120','box4','solids6']:
    sc,g=fx.load(name)
    w,h=1920,1080
    fmt=_capi.make_image_format(w,h,_capi.RGB8,pitch=w*3+32)
    os.environ['NTR_ZEROCOPY']='0'
    with DeviceScene(sc) as ds:
        a=ds.render(fmt, torch.full((fmt.pitch*h,),7,dtype=torch.uint8).pin_memory().numpy()).copy()
    os.environ['NTR_ZEROCOPY']='1'
    with DeviceScene(sc) as ds:
        b=ds.render(fmt, torch.full((fmt.pitch*h,),7,dtype=torch.uint8).pin_memory().numpy()).copy()
        c=ds.render(fmt, np.full(fmt.pitch*h,7,np.uint8)).copy()     # pageable: falls back to the copy
    print(name,'zero-copy == copy:',np.array_equal(a,b),np.array_equal(a,c))
P
for c in c2 c1 c3; do
for z in 0 1; do
NTR_ZEROCOPY=$z python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --stream-frames 0 > gpurun_out/zc${z}_$c.json 2>gpurun_out/zc${z}_$c.err
done; done
cat gpurun_out/zc_check.txt
