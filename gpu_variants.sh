for v in ntracer_b200/libntracer_b200.so ntracer_b200/variants/libntr_mb0.so ntracer_b200/variants/libntr_mb16.so; do
  echo "== $v"
  NTR_B200_LIB=$PWD/$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 ms',round(d['ms_per_step'],4),'Mrays/s',round(d['value'],1),'e2e ms',round(d['e2e']['ms_per_step'],4),'frac',round(d['roofline']['frac'],4))"
  NTR_B200_LIB=$PWD/$v python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4 ms',round(d['ms_per_step'],3),'Mrays/s',round(d['value'],1))"
done
