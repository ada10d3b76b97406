mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 --gemm-table > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_kernel_table.txt
cut -c1-200 gpurun_out/r02h_bench.json | tail -1
for tool in memcheck racecheck synccheck; do
  timeout 110 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/r02h_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|finished" gpurun_out/r02h_sanitizer_$tool.log | tail -3
done
