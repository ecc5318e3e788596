# Which part of the attention kernel bounds it (garbage results, real timing): compile-time ablation builds
# (python build_native.py --variant abl<mask> -DLDMSEG_ATTN_ABLATE=<mask>).
TAG=${1:-r02d}
O=gpurun_out/$TAG
mkdir -p $O
L=$PWD/latent-diffusion-segmentation_b200/lib
for m in ${MASKS:-0 1 4 8 16 32 12 63 64 127}; do
  for alt in 1 0; do
  echo "== ablation mask $m alternate $alt"
  if [ $m = 0 ]; then lib=$L/libldmseg_b200.so; else lib=$L/libldmseg_b200_abl$m.so; fi
  LDMSEG_ATTN_ALTERNATE=$alt LDMSEG_LIB=$lib timeout 200 python tools/bench_attn.py 2>&1 | grep -E "d=40|ntok=1024"
  done
done | tee $O/attn_ablate${SUFFIX}.log
