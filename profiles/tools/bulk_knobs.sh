run() { echo "== $*"; env "$@" timeout 200 python profiles/tools/bulk_ab.py perlin 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('perlin'):
        d = json.loads(l.split(' ',1)[1]); print(round(d['bulk']['ms'],4), d['bulk']['nodes'], [(a.replace('bulk_','').replace('_kernel',''), b) for a,b in d['bulk']['stages']])
    elif 'Error' in l: print(l)
"; }
run A=1
run VX_LIB=/root/repo/voxelis_b200/libvoxelis_b200_c4.so
run VX_PLAN_UPW=4
run A=2
