mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02e_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 --gemm-table > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_kernel_table.txt
tail -3 gpurun_out/r02e_pytest_gpu.txt; cut -c1-300 gpurun_out/r02e_bench.json; grep -E "ms " gpurun_out/r02e_kernel_table.txt | head -30
