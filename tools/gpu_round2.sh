# Round-2 GPU session: kernel checks, parity tests (errors -> r02_parity.json), smoke, bench (batch 1 + config3 +
# library baseline), in-graph A/Bs of the round-2 switches.  Usage (from the repo root, under gpurun):
#   bash tools/gpu_round2.sh <tag> [quick]
TAG=${1:-r02a}
MODE=${2:-full}
O=gpurun_out/$TAG
mkdir -p $O
export LDMSEG_PARITY_OUT=$PWD/$O/r02_parity.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
timeout 1200 python tools/kernel_check.py > $O/kernel_check.log 2>&1; echo "kernel_check rc=$?"; grep -E "FAIL|GROUP|TIMEOUT|Error|error" $O/kernel_check.log | head -60
timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_gpu_parity.py::test_kernel_checks > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "parity\]|passed|failed|Error|assert" $O/pytest_gpu.log | head -60
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -5 $O/bench.err
if [ "$MODE" = "full" ]; then
  for b in 1 8; do
    echo "== ablate batch $b (default)"; timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
    echo "== ablate batch $b LDMSEG_NEXTW=0"; LDMSEG_NEXTW=0 timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
    echo "== ablate batch $b LDMSEG_S2_TMA=0"; LDMSEG_S2_TMA=0 timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
    echo "== ablate batch $b LDMSEG_ATTN_POLY=1"; LDMSEG_ATTN_POLY=1 timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
    echo "== ablate batch $b LDMSEG_ATTN_POLY=0"; LDMSEG_ATTN_POLY=0 timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
    echo "== ablate batch $b LDMSEG_RESID_F32=1"; LDMSEG_RESID_F32=1 timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
  done
  timeout 300 python tools/ablate_unet.py --batch 1 --per-op > $O/ablate_b1.log 2>&1; head -34 $O/ablate_b1.log
  timeout 300 python tools/ablate_unet.py --batch 8 --per-op > $O/ablate_b8.log 2>&1; head -34 $O/ablate_b8.log
fi
du -sh $O; ls $O
