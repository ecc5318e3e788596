# Round-2 closing profiles of the final kernels: ncu launch lists of the UNet forward (batch 1 / 8) and of the bench's
# timed region, --set full summaries of the igemm launches (batch 1 / 8; the batch-8 one holds the 320-wide pair tile).
# Usage: bash tools/gpu_profile3.sh <tag>
TAG=${1:-r02q}
O=gpurun_out/$TAG
mkdir -p $O
NCU="ncu --profile-from-start off --clock-control none"
for b in 1 8; do
  timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_unet_b$b.csv python tools/profile_unet.py --batch $b > /dev/null 2>&1
  python tools/summarize_launches.py $O/launches_unet_b$b.csv > $O/launches_unet_b$b.txt 2>&1; head -24 $O/launches_unet_b$b.txt
done
LDMSEG_PROFILE=1 timeout 400 $NCU --metrics gpu__time_duration.sum --csv -c 1400 --log-file $O/launches_bench_b1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config3 > /dev/null 2>&1
python tools/summarize_launches.py $O/launches_bench_b1.csv > $O/launches_bench_b1.txt 2>&1; head -20 $O/launches_bench_b1.txt
head -400 $O/launches_bench_b1.csv > $O/launches_bench_b1_head.csv; rm -f $O/launches_bench_b1.csv
timeout 400 $NCU --set full -k regex:igemm -s 2 -c 14 -o $O/igemm_full_b8 -f python tools/profile_unet.py --batch 8 > /dev/null 2>&1
timeout 400 $NCU --set full -k regex:igemm -s 2 -c 14 -o $O/igemm_full_b1 -f python tools/profile_unet.py --batch 1 > /dev/null 2>&1
for f in igemm_full_b8 igemm_full_b1; do
  python tools/ncu_summary.py $O/$f.ncu-rep > $O/$f.txt 2>&1
  rm -f $O/$f.ncu-rep
done
cut -c1-260 $O/igemm_full_b8.txt | head -20
du -sh $O; ls $O
# bench line of the final state (batch 1 + config3 + library baseline), then BASELINE configs[3] / configs[4] on one GPU
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-400 $O/bench.json
timeout 600 python bench.py --config inpaint --batch 4 --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > $O/bench_config4_inpaint_b4.json 2> $O/bench_config4.err; cut -c1-300 $O/bench_config4_inpaint_b4.json
timeout 900 python bench.py --config ddpm --size 1024 --ddim-steps 100 --batch 2 --steps 2 --warmup 3 --no-cpu-baseline --no-config3 > $O/bench_config5_ddpm_1024_b2.json 2> $O/bench_config5.err; cut -c1-300 $O/bench_config5_ddpm_1024_b2.json; tail -3 $O/bench_config5.err
ls $O
