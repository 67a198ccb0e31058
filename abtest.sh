for lib in voxelis_b200/alt/lib_min3.so voxelis_b200/alt/lib_min4.so voxelis_b200/alt/lib_noinl4.so; do
  echo "== lib=$lib"
  VX_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --no-cpu --no-e2e --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('perlin', round(d['roofline']['kernel_ms'],4))
for k,v in d['others'].items(): print(k, round(v['kernel_ms'],4))
"
done
