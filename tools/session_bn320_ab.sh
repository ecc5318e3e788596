#!/bin/bash
# A/B of the 320-wide pair tiles (LDMSEG_BN320) inside the UNet forward graph, interleaved on one box
O=gpurun_out/r04c; mkdir -p $O
timeout 300 python tools/kernel_check.py --group igemm_up2 > $O/kernel_check_up2.log 2>&1; echo "up2 check rc=$?"; grep -c PASS $O/kernel_check_up2.log; grep -E "FAIL|320" $O/kernel_check_up2.log
for b in 8 4; do for f in 0 1 0 1; do
  LDMSEG_BN320=$f timeout 300 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -1 | sed "s/^/b$b bn320=$f: /" | tee -a $O/ab_bn320.log
done; done
LDMSEG_BN320=1 timeout 300 python tools/ablate_unet.py --batch 8 --per-op > $O/ablate_b8.log 2>&1; head -12 $O/ablate_b8.log; grep -c "bn320" $O/ablate_b8.log
LDMSEG_BN320=1 timeout 300 python tools/ablate_unet.py --batch 2 --full-only 2>&1 | tail -1 | sed "s/^/b2 bn320=1: /" | tee -a $O/ab_bn320.log
LDMSEG_BN320=0 timeout 300 python tools/ablate_unet.py --batch 2 --full-only 2>&1 | tail -1 | sed "s/^/b2 bn320=0: /" | tee -a $O/ab_bn320.log
