#!/bin/bash
O=gpurun_out/r04e; mkdir -p $O
timeout 400 ncu --set full --import-source on --clock-control none -k regex:igemm -s 3 -c 1 -o $O/bn320_src -f python tools/bench_bn320.py --profile 8 64 320 320 320 > $O/ncu.log 2>&1; tail -3 $O/ncu.log
ncu -i $O/bn320_src.ncu-rep --page source --csv > $O/bn320_src.csv 2>/dev/null
python tools/ncu_source_hot.py $O/bn320_src.csv 70 > $O/bn320_src_hot.txt 2>&1; head -90 $O/bn320_src_hot.txt
python tools/ncu_summary.py $O/bn320_src.ncu-rep > $O/bn320_summary.txt 2>&1; cat $O/bn320_summary.txt | cut -c1-300 | head -20
rm -f $O/bn320_src.csv
ls -la $O
