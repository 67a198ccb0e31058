for c in "0.125,0.5" "0.1,0.35,0.65,0.9" "0.125,0.375,0.625,0.875" "0.06,0.2,0.4,0.6,0.8,0.93" "0.25,0.5,0.75"; do
  echo "== cuts $c"; VX_STAGE_CUTS=$c timeout 120 python profiles/tools/e2e_workload.py 12 2>&1 | tail -1
done
