#!/bin/bash
O=gpurun_out/r04d; mkdir -p $O
timeout 300 python tools/bench_bn320.py --stages 2>&1 | tee $O/bench_bn320_stages.log
