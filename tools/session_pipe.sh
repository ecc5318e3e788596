#!/bin/bash
# software-pipelined DIRECT epilogue + early TMEM release: checks, then A/B against the previous build (lib ..._epi8.so)
O=gpurun_out/r04g; mkdir -p $O
L=$PWD/latent-diffusion-segmentation_b200/lib
for g in igemm_plain igemm_conv igemm_epi igemm_splitk igemm_streamk igemm_pair igemm_bn320 igemm_s2 igemm_up2 igemm_f32stream igemm_lnfold gn_fused vae_pdl; do
  timeout 300 python tools/kernel_check.py --group $g > $O/kc_$g.log 2>&1; echo "$g rc=$? pass=$(grep -c PASS $O/kc_$g.log) fail=$(grep -c FAIL $O/kc_$g.log)"
done
cat $O/kc_*.log > $O/kernel_check_igemm.log; rm -f $O/kc_*.log
echo "== new"; timeout 300 python tools/bench_bn320.py --batches 8 2>&1 | tee $O/bench_bn320_new.log | grep -v "^nb=. 16x16"
echo "== old"; LDMSEG_LIB=$L/libldmseg_b200_epi8.so timeout 300 python tools/bench_bn320.py --batches 8 2>&1 | tee $O/bench_bn320_old.log | grep -v "^nb=. 16x16"
for b in 1 8; do for f in new old new old; do
  lib=$L/libldmseg_b200.so; [ $f = old ] && lib=$L/libldmseg_b200_epi8.so
  LDMSEG_LIB=$lib timeout 300 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -1 | sed "s/^/b$b $f: /" | tee -a $O/ab_pipe.log
done; done
