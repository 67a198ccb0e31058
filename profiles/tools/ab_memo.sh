# A/B runs of the busy-unit path (profiles/README.md): bash profiles/tools/ab_memo.sh "cfg1" "cfg2" ...
for wlname in below random255 checkerboard sum; do
for cfg in "$@"; do
  echo "== $wlname $cfg"
  env $cfg python bench.py --workload $wlname --steps 100 --no-others --no-e2e --no-cpu --no-d7 --no-traffic 2>&1 >/dev/null | grep -E "device-res|stages"
done; done
