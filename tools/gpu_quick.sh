# quick validation: norm kernel checks, in-graph ablation, bench batch 1 / 8.  Usage: bash tools/gpu_quick.sh <tag>
TAG=${1:-quick}
O=gpurun_out/$TAG
mkdir -p $O
for g in gn_fused norm; do timeout 240 python tools/kernel_check.py --group $g 2>&1 | grep -E "FAIL|GROUP"; done
timeout 300 python tools/ablate_unet.py --batch 1 > $O/ablate_b1.log 2>&1; head -8 $O/ablate_b1.log
timeout 300 python tools/ablate_unet.py --batch 8 > $O/ablate_b8.log 2>&1; head -8 $O/ablate_b8.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_b1.json 2> $O/bench_b1.err; echo "bench rc=$?"; cut -c1-700 $O/bench_b1.json
timeout 600 python bench.py --steps 3 --warmup 3 --batch 8 --no-cpu-baseline > $O/bench_b8.json 2> $O/bench_b8.err; cut -c1-300 $O/bench_b8.json
