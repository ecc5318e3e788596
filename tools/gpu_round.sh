set -x
mkdir -p gpurun_out
for pm in 0 1 2; do
LDMSEG_ATTN_POLY=$pm timeout 300 python tools/kernel_check.py --group attn > gpurun_out/kernel_check_attn_pm$pm.log 2>&1; echo "pm=$pm kernel_check rc=$?"; grep -E "FAIL|GROUP|PASS attn nb=1 ntok=4096" gpurun_out/kernel_check_attn_pm$pm.log
LDMSEG_ATTN_POLY=$pm timeout 300 python tools/bench_attn.py > gpurun_out/bench_attn_pm$pm.log 2>&1; cat gpurun_out/bench_attn_pm$pm.log
done
