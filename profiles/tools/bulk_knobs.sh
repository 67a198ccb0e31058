for cfg in "" "VX_BULK_NO_PDL=1" "VX_PLAN_UPW=28" "VX_PLAN_UPW=1" "VX_PLAN_UPW=2" "VX_BULK_DENSE_MIN=64" "VX_BULK_DENSE_MIN=256"; do
  echo "== $cfg"; env $cfg timeout 200 python profiles/tools/bulk_ab.py perlin 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('perlin'):
        d = json.loads(l.split(' ',1)[1]); print(d['bulk']['ms'], d['bulk']['stages'])
    elif 'Error' in l: print(l)
"
done
