# A/B of experimental builds of the library (same ABI): bash profiles/tools/ab_libs.sh "wl1 wl2 ..." lib1.so lib2.so ...
wls="$1"; shift
for wl in $wls; do for lib in "$@"; do
  echo "== $wl $lib"
  VX_LIB=$PWD/$lib timeout 200 python bench.py --workload $wl --steps 100 --no-others --no-e2e --no-cpu --no-d7 --no-traffic 2>&1 >/dev/null | grep -E "device-res|stages"
done; done
