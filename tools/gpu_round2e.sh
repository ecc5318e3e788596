# Round-2 final session: full kernel checks, GPU parity tests, smoke, bench (batch 1 + config3 + library baseline).
TAG=${1:-r02h}
O=gpurun_out/$TAG
mkdir -p $O
export LDMSEG_PARITY_OUT=$PWD/$O/r02_parity.json
timeout 1200 python tools/kernel_check.py > $O/kernel_check.log 2>&1; echo "kernel_check rc=$?"; grep -E "FAIL|GROUP|TIMEOUT|Error|error" $O/kernel_check.log | head -40
timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_gpu_parity.py::test_kernel_checks > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|assert" $O/pytest_gpu.log | head -20
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-1500 $O/bench.json; tail -5 $O/bench.err
du -sh $O; ls $O
