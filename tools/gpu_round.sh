set -x
mkdir -p gpurun_out
python tools/kernel_check.py > gpurun_out/kernel_check_3.log 2>&1; echo "kernel_check rc=$?"
grep -E "FAIL|GROUP|Error|error" gpurun_out/kernel_check_3.log | head -30
python tools/bench_igemm.py > gpurun_out/bench_igemm_6.log 2>&1
python tools/bench_epi.py > gpurun_out/bench_epi_2.log 2>&1
python tools/ablate_unet.py --batch 1 --per-op > gpurun_out/ablate_b1_v9.log 2>&1
python tools/ablate_unet.py --batch 8 > gpurun_out/ablate_b8_v9.log 2>&1
head -60 gpurun_out/ablate_b1_v9.log
