# Round-2 profiles: ncu launch lists (UNet forward b1/b8, VAE encode, seg decode, the bench's timed region), --set full
# summaries of the dominant kernels, a source-level capture of the d=40 attention kernel, in-graph ablations, the
# drop-in API loop timing and configs[3]/[4] on one GPU.   Usage: bash tools/gpu_profile2.sh <tag>
TAG=${1:-r02p}
O=gpurun_out/$TAG
mkdir -p $O
NCU="ncu --profile-from-start off --clock-control none"
for b in 1 8; do
  timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_unet_b$b.csv python tools/profile_unet.py --batch $b > /dev/null 2>&1
  python tools/summarize_launches.py $O/launches_unet_b$b.csv > $O/launches_unet_b$b.txt 2>&1; head -30 $O/launches_unet_b$b.txt
done
for w in encode decode; do
  timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_$w.csv python tools/profile_unet.py --what $w > /dev/null 2>&1
  python tools/summarize_launches.py $O/launches_$w.csv > $O/launches_$w.txt 2>&1; head -16 $O/launches_$w.txt
done
LDMSEG_PROFILE=1 timeout 400 $NCU --metrics gpu__time_duration.sum --csv -c 1400 --log-file $O/launches_bench_b1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config3 > /dev/null 2>&1
python tools/summarize_launches.py $O/launches_bench_b1.csv > $O/launches_bench_b1.txt 2>&1; head -30 $O/launches_bench_b1.txt
head -400 $O/launches_bench_b1.csv > $O/launches_bench_b1_head.csv; rm -f $O/launches_bench_b1.csv
timeout 400 $NCU --set full -k regex:igemm -s 2 -c 14 -o $O/igemm_full_b8 -f python tools/profile_unet.py --batch 8 > /dev/null 2>&1
timeout 400 $NCU --set full -k regex:igemm -s 2 -c 14 -o $O/igemm_full_b1 -f python tools/profile_unet.py --batch 1 > /dev/null 2>&1
timeout 400 $NCU --set full -k "regex:attn|gn_" -c 8 -o $O/attn_gn_full_b8 -f python tools/profile_unet.py --batch 8 > /dev/null 2>&1
timeout 400 $NCU --set full -k "regex:attn|gn_" -c 8 -o $O/attn_gn_full_b1 -f python tools/profile_unet.py --batch 1 > /dev/null 2>&1
for f in igemm_full_b8 igemm_full_b1 attn_gn_full_b8 attn_gn_full_b1; do
  python tools/ncu_summary.py $O/$f.ncu-rep > $O/$f.txt 2>&1
  rm -f $O/$f.ncu-rep
done
cat $O/igemm_full_b1.txt | cut -c1-260 | head -20
# source-level view of the d = 40 attention kernel (batch 8, one launch)
timeout 400 $NCU --set full --import-source on -k regex:attn_kernel -c 1 -o $O/attn_src -f python tools/profile_unet.py --batch 8 > /dev/null 2>&1
ncu -i $O/attn_src.ncu-rep --page source --csv > $O/attn_src.csv 2>/dev/null
python tools/ncu_source_hot.py $O/attn_src.csv 60 > $O/attn_src_hot.txt 2>&1; head -70 $O/attn_src_hot.txt
rm -f $O/attn_src.ncu-rep $O/attn_src.csv
# the drop-in API loop (unet(...) + scheduler.step(...) as TrainerDiffusion.sample drives them) vs the fused sampler
timeout 300 python tools/bench_dropin.py > $O/dropin.json 2>&1; cat $O/dropin.json
# BASELINE configs[3] (inpainting, 4 per GPU) and configs[4] (1024x1024, 100-step DDPM, 2 per GPU) on one GPU
timeout 600 python bench.py --config inpaint --batch 4 --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > $O/bench_config4_inpaint_b4.json 2> $O/bench_config4.err; cat $O/bench_config4_inpaint_b4.json | cut -c1-900
timeout 900 python bench.py --config ddpm --size 1024 --ddim-steps 100 --batch 2 --steps 2 --warmup 3 --no-cpu-baseline --no-config3 > $O/bench_config5_ddpm_1024_b2.json 2> $O/bench_config5.err; cat $O/bench_config5_ddpm_1024_b2.json | cut -c1-900; tail -3 $O/bench_config5.err
du -sh $O; ls $O
