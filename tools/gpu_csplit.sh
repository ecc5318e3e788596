# cluster split-K: kernel checks, then in-graph A/B of the UNet forward.
TAG=${1:-r02g}
O=gpurun_out/$TAG
mkdir -p $O
for g in igemm_splitk igemm_plain igemm_streamk igemm_pair; do
  timeout 300 python tools/kernel_check.py --group $g > $O/kc_$g.log 2>&1; echo "kernel_check $g rc=$?"; grep -E "FAIL|max clusters" $O/kc_$g.log | head -40
done
grep -E "cluster split" $O/kc_igemm_splitk.log | cut -c1-160
for b in 1 2 4 8; do for c in 0 1; do
  echo "== batch $b LDMSEG_SPLIT_CLUSTER=$c"; LDMSEG_SPLIT_CLUSTER=$c timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
done; done 2>&1 | tee $O/ab_csplit.log
timeout 300 python tools/ablate_unet.py --batch 1 --per-op > $O/ablate_b1.log 2>&1; head -12 $O/ablate_b1.log
