set -x
mkdir -p gpurun_out
timeout 900 python tools/kernel_check.py > gpurun_out/kernel_check_7.log 2>&1; echo "kernel_check rc=$?"; grep -E "FAIL|GROUP" gpurun_out/kernel_check_7.log
for v in prev new prev new; do
  if [ $v = prev ]; then export LDMSEG_LIB=$PWD/latent-diffusion-segmentation_b200/lib/prev_libldmseg_b200.so; else unset LDMSEG_LIB; fi
  timeout 600 python tools/ablate_unet.py --batch 8 > gpurun_out/ablate_b8_$v.log 2>&1
  timeout 600 python tools/ablate_unet.py --batch 1 > gpurun_out/ablate_b1_$v.log 2>&1
  echo "== $v"; grep -E "full|family igemm" gpurun_out/ablate_b8_$v.log gpurun_out/ablate_b1_$v.log
done
