# Attention A/B: share of exponentials on the FMA pipe; skeleton (all stages removed) and per-phase clock attribution.
TAG=${1:-r02d}
O=gpurun_out/$TAG
mkdir -p $O
L=$PWD/latent-diffusion-segmentation_b200/lib
for g in attn xattn; do
  timeout 300 python tools/kernel_check.py --group $g > $O/kc_$g.log 2>&1; echo "kernel_check $g rc=$?"; grep -E "FAIL|OK" $O/kc_$g.log | tail -12
done
for lib in libldmseg_b200.so; do for pm in 2 1 0; do
  echo "== $lib LDMSEG_ATTN_POLY=$pm"
  LDMSEG_LIB=$L/$lib LDMSEG_ATTN_POLY=$pm timeout 200 python tools/bench_attn.py 2>&1 | tail -6
done; done 2>&1 | tee $O/attn_ab4.log
echo "== skeleton (ablation mask 127)"; LDMSEG_LIB=$L/libldmseg_b200_abl127.so timeout 200 python tools/bench_attn.py 2>&1 | tail -6 | tee -a $O/attn_ab4.log
for lib in libldmseg_b200_tim.so; do for pm in 2; do
  echo "== $lib LDMSEG_ATTN_POLY=$pm"
  LDMSEG_LIB=$L/$lib LDMSEG_ATTN_POLY=$pm timeout 200 python tools/attn_timing.py 2>&1
done; done 2>&1 | tee $O/attn_timing3.log
