#!/bin/bash
O=gpurun_out/r04f; mkdir -p $O
L=$PWD/latent-diffusion-segmentation_b200/lib
timeout 300 python tools/kernel_check.py --group igemm_bn320 > $O/kernel_check_bn320.log 2>&1; echo "bn320 check rc=$?"; grep -c PASS $O/kernel_check_bn320.log; grep FAIL $O/kernel_check_bn320.log
echo "== 16 epilogue warps"; timeout 300 python tools/bench_bn320.py --batches 8 4 2>&1 | tee $O/bench_bn320_epi16.log | grep -v "^nb=. 16x16"
echo "== 8 epilogue warps"; LDMSEG_LIB=$L/libldmseg_b200_epi8.so timeout 300 python tools/bench_bn320.py --batches 8 2>&1 | tee $O/bench_bn320_epi8.log | grep -v "^nb=. 16x16"
echo "== stages, 16 warps"; timeout 300 python tools/bench_bn320.py --stages 2>&1 | tee $O/bench_bn320_stages_epi16.log | grep "^nb"
for f in 16 8 16 8; do
  lib=$L/libldmseg_b200.so; [ $f = 8 ] && lib=$L/libldmseg_b200_epi8.so
  LDMSEG_LIB=$lib timeout 300 python tools/ablate_unet.py --batch 8 --full-only 2>&1 | tail -1 | sed "s/^/b8 epi=$f: /" | tee -a $O/ab_epi.log
done
