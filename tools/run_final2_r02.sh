mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_optim.py -m gpu -q 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 5 --gemm-table > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_kernel_table.txt
cut -c1-160 gpurun_out/r02i_bench.json | tail -1
