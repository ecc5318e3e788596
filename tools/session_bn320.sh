#!/bin/bash
O=gpurun_out/r04b; mkdir -p $O
timeout 300 python tools/kernel_check.py --group igemm_bn320 > $O/kernel_check_bn320.log 2>&1; echo "bn320 check rc=$?"
tail -40 $O/kernel_check_bn320.log
timeout 300 python tools/kernel_check.py --group igemm_pair > $O/kernel_check_pair.log 2>&1; echo "pair check rc=$?"; grep -c PASS $O/kernel_check_pair.log; grep FAIL $O/kernel_check_pair.log
timeout 300 python tools/kernel_check.py --group igemm_streamk > $O/kernel_check_streamk.log 2>&1; echo "streamk check rc=$?"; grep -c PASS $O/kernel_check_streamk.log; grep FAIL $O/kernel_check_streamk.log
timeout 400 python tools/bench_bn320.py > $O/bench_bn320.log 2>&1; echo "bench rc=$?"; cat $O/bench_bn320.log
