# ncu launch lists + small full captures (summaries only) + per-op ablation.  Usage: bash tools/gpu_profile.sh <tag>
TAG=${1:-prof}
O=gpurun_out/$TAG
mkdir -p $O
for g in gn_fused norm igemm_pair; do timeout 240 python tools/kernel_check.py --group $g 2>&1 | grep -E "FAIL|GROUP"; done
timeout 300 python tools/ablate_unet.py --batch 1 --per-op > $O/ablate_b1.log 2>&1; head -8 $O/ablate_b1.log
timeout 300 python tools/ablate_unet.py --batch 8 --per-op > $O/ablate_b8.log 2>&1; head -8 $O/ablate_b8.log
NCU="ncu --profile-from-start off --clock-control none"
for b in 1 8; do
  timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_unet_b$b.csv python tools/profile_unet.py --batch $b > /dev/null 2>&1
  python tools/summarize_launches.py $O/launches_unet_b$b.csv > $O/launches_unet_b$b.txt 2>&1
done
for w in encode decode; do
  timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_$w.csv python tools/profile_unet.py --what $w > /dev/null 2>&1
  python tools/summarize_launches.py $O/launches_$w.csv > $O/launches_$w.txt 2>&1
done
timeout 400 $NCU --set full -k regex:igemm -s 2 -c 14 -o $O/igemm_full_b8 -f python tools/profile_unet.py --batch 8 > /dev/null 2>&1
timeout 400 $NCU --set full -k regex:igemm -s 2 -c 14 -o $O/igemm_full_b1 -f python tools/profile_unet.py --batch 1 > /dev/null 2>&1
timeout 400 $NCU --set full -k "regex:attn|gn_|layernorm" -c 10 -o $O/attn_gn_full_b8 -f python tools/profile_unet.py --batch 8 > /dev/null 2>&1
timeout 400 $NCU --set full -k "regex:attn|gn_|layernorm" -c 10 -o $O/attn_gn_full_b1 -f python tools/profile_unet.py --batch 1 > /dev/null 2>&1
for f in igemm_full_b8 igemm_full_b1 attn_gn_full_b8 attn_gn_full_b1; do
  python tools/ncu_summary.py $O/$f.ncu-rep > $O/$f.txt 2>&1
  rm -f $O/$f.ncu-rep
done
du -sh $O
