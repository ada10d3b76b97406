set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:gemm_tcgen05|sattn|xattn|norm_bwd|layernorm' -f -o gpurun_out/r02_targets python tools/ncu_targets.py > gpurun_out/r02_targets_ncu.log 2>&1
tail -5 gpurun_out/r02_targets_ncu.log
