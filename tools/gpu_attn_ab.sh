# Attention A/B: column halves (16 softmax warps) against one thread per score row; share of exponentials on the FMA pipe.
TAG=${1:-r02h}
O=gpurun_out/$TAG
mkdir -p $O
for g in attn xattn; do
  timeout 300 python tools/kernel_check.py --group $g > $O/kc_$g.log 2>&1; echo "kernel_check $g rc=$?"; grep -E "FAIL|OK" $O/kc_$g.log | tail -12
done
for hs in 1 2; do for pm in 2 0 1; do
  echo "== LDMSEG_ATTN_HALVES=$hs LDMSEG_ATTN_POLY=$pm"
  LDMSEG_ATTN_HALVES=$hs LDMSEG_ATTN_POLY=$pm timeout 200 python tools/bench_attn.py 2>&1 | tail -6
done; done 2>&1 | tee $O/attn_ab_halves.log
