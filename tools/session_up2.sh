#!/bin/bash
# A/B of the folded up-sampling (LDMSEG_UP2_FOLD) on one box: kernel checks, smoke, forward time at batch 1 / 8
O=gpurun_out/r04a; mkdir -p $O
timeout 300 python tools/kernel_check.py --group igemm_up2 > $O/kernel_check_up2.log 2>&1; echo "up2 check rc=$?"
tail -45 $O/kernel_check_up2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
for b in 1 8; do for f in 0 1 0 1; do
  LDMSEG_UP2_FOLD=$f timeout 300 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -1 | sed "s/^/b$b fold=$f: /" | tee -a $O/ab_up2.log
done; done
LDMSEG_UP2_FOLD=1 timeout 300 python tools/ablate_unet.py --batch 1 --per-op > $O/ablate_b1.log 2>&1; grep -E "\.up:|upsample|full graph" $O/ablate_b1.log
LDMSEG_UP2_FOLD=1 timeout 300 python tools/ablate_unet.py --batch 8 --per-op > $O/ablate_b8.log 2>&1; grep -E "\.up:|upsample|full graph" $O/ablate_b8.log
