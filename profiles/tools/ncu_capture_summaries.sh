# On the GPU box: ncu --set full captures of one apply call of each workload, summarised to text (the .ncu-rep files stay
# in /tmp: gpurun brings back at most 64 MiB).  usage: bash profiles/tools/ncu_capture_summaries.sh <tag> <workload>...
tag=$1; shift
for w in "$@"; do
  name=${w%%:*}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"bulk_" -s 5 -c 5 -o /tmp/${tag}_$name python profiles/tools/ncu_traffic.py $w > /dev/null 2>&1
  { python profiles/tools/ncu_summary.py /tmp/${tag}_$name.ncu-rep
    echo; echo "== stall samples per source line: bulk_blocks_kernel<u8> (+ busy units)"
    NCU_ARGS="-k regex:bulk_blocks" python profiles/tools/ncu_lines.py /tmp/${tag}_$name.ncu-rep voxelis_b200/libvoxelis_b200.so bulk_blocks_kernelIh 28
    if [ "$name" = perlin ]; then
      echo; echo "== stall samples per source line: bulk_level_kernel<u8>, level 1"
      NCU_ARGS="-k regex:bulk_level" NCU_CHUNK=0 python profiles/tools/ncu_lines.py /tmp/${tag}_$name.ncu-rep voxelis_b200/libvoxelis_b200.so bulk_level_kernelIh 24
      echo; echo "== stall samples per source line: bulk_upper_kernel<u8> (units .. roots)"
      NCU_ARGS="-k regex:bulk_upper" python profiles/tools/ncu_lines.py /tmp/${tag}_$name.ncu-rep voxelis_b200/libvoxelis_b200.so bulk_upper_kernelIh 24
      echo; echo "== stall samples per source line: bulk_plan_kernel"
      NCU_ARGS="-k regex:bulk_plan" python profiles/tools/ncu_lines.py /tmp/${tag}_$name.ncu-rep voxelis_b200/libvoxelis_b200.so bulk_plan_kernel 16
    fi
  } > gpurun_out/${tag}_ncu_${name}_summary.txt 2>&1
done
ls -la gpurun_out/
