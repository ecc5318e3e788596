# Round-2 follow-up session: kernel checks, the full GPU test-suite, in-graph A/B of the LayerNorm fold.
TAG=${1:-r02b}
O=gpurun_out/$TAG
mkdir -p $O
export LDMSEG_PARITY_OUT=$PWD/$O/r02_parity.json
timeout 1200 python tools/kernel_check.py > $O/kernel_check.log 2>&1; echo "kernel_check rc=$?"; grep -E "FAIL|FAILED|TIMEOUT|Error|error|panoptic image|VAE encoder|ln-fold" $O/kernel_check.log | head -60
timeout 1800 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_parity.py::test_kernel_checks > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "parity\]|passed|failed|Error|^E  |FAILED" $O/pytest_gpu.log | head -80
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
for b in 1 8; do
  echo "== ablate batch $b (default)"; timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
  echo "== ablate batch $b LDMSEG_LN_FOLD=0"; LDMSEG_LN_FOLD=0 timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-library-baseline --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("b1", d["value"], d["e2e"]["value"], d["roofline"]["unet_forward_ms_graph"], d["roofline"]["frac"], d["roofline_norm"]["frac"], d["roofline_norm"]["avg_launch_us"])
c=d["config3"]; print("b8", c["value"], c["e2e"]["value"], c["roofline"]["unet_forward_ms_graph"], c["roofline"]["frac"], c["roofline_norm"]["frac"])
PY
tail -3 $O/bench.err
timeout 300 python tools/ablate_unet.py --batch 1 --per-op > $O/ablate_b1.log 2>&1; head -12 $O/ablate_b1.log
ls $O
