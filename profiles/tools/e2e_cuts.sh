# e2e leg with different slice cuts (VX_STAGE_CUTS) — profiles/README.md
for c in "0.125,0.5" "0.125,0.5,0.8" "0.1,0.4,0.7,0.9" "0.2,0.6" "0.15,0.45,0.75" "0.25,0.6,0.85"; do
  echo "== cuts $c"; VX_STAGE_JCTAS=1 VX_STAGE_CUTS=$c timeout 120 python profiles/tools/e2e_workload.py 12 2>&1 | tail -1 | cut -c1-330
done
