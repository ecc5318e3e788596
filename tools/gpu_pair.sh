# Validate the CTA-pair igemm and A/B it against the 1-CTA path.  Usage: bash tools/gpu_pair.sh <tag>
TAG=${1:-pair}
O=gpurun_out/$TAG
mkdir -p $O
for g in igemm_pair igemm_epi igemm_splitk; do
  timeout 240 python tools/kernel_check.py --group $g > $O/check_$g.log 2>&1; echo "$g rc=$?"
  grep -E "FAIL|PASS" $O/check_$g.log | cut -c1-150
  tail -3 $O/check_$g.log | cut -c1-300
done
if grep -q FAIL $O/check_igemm_pair.log || ! grep -q PASS $O/check_igemm_pair.log; then echo "pair kernel broken: stopping"; exit 1; fi
for b in 8 1; do
  for pair in 1 0; do
    LDMSEG_PAIR=$pair timeout 300 python tools/ablate_unet.py --batch $b > $O/ablate_b${b}_pair$pair.log 2>&1
    echo "== batch $b pair=$pair"; grep -E "full graph|family igemm|igemm rows" $O/ablate_b${b}_pair$pair.log
  done
done
