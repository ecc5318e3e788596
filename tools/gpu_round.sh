# One GPU round: parity tests, smoke, bench (batch 1 + batch 8), in-graph ablation, ncu launch lists + small full captures.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>     (keeps gpurun_out/<tag> well below 64 MiB)
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_b1.json 2> $O/bench_b1.err; echo "bench rc=$?"; cat $O/bench_b1.json
timeout 600 python bench.py --steps 3 --warmup 3 --batch 8 --no-cpu-baseline > $O/bench_b8.json 2> $O/bench_b8.err; cat $O/bench_b8.json
timeout 300 python tools/ablate_unet.py --batch 1 --per-op > $O/ablate_b1.log 2>&1; head -12 $O/ablate_b1.log
timeout 300 python tools/ablate_unet.py --batch 8 --per-op > $O/ablate_b8.log 2>&1; head -12 $O/ablate_b8.log
NCU="ncu --profile-from-start off --clock-control none"
for b in 1 8; do
  timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_unet_b$b.csv python tools/profile_unet.py --batch $b > /dev/null 2>&1
  python tools/summarize_launches.py $O/launches_unet_b$b.csv > $O/launches_unet_b$b.txt 2>&1
done
# launch list of the bench command itself: the first 1400 launches of its timed region (encode + ~5 DDIM steps)
LDMSEG_PROFILE=1 timeout 400 $NCU --metrics gpu__time_duration.sum --csv -c 1400 --log-file $O/launches_bench_b1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_launches.py $O/launches_bench_b1.csv > $O/launches_bench_b1.txt 2>&1
for w in encode decode; do
  timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_$w.csv python tools/profile_unet.py --what $w > /dev/null 2>&1
  python tools/summarize_launches.py $O/launches_$w.csv > $O/launches_$w.txt 2>&1
done
# full captures: a window of the forward that holds 64x64-level convs, a transformer block and its attention
timeout 400 $NCU --set full --import-source on -k regex:igemm -s 2 -c 14 -o $O/igemm_full_b8 -f python tools/profile_unet.py --batch 8 > /dev/null 2>&1
timeout 400 $NCU --set full -k regex:igemm -s 2 -c 14 -o $O/igemm_full_b1 -f python tools/profile_unet.py --batch 1 > /dev/null 2>&1
timeout 400 $NCU --set full -k "regex:attn|gn_|layernorm" -c 10 -o $O/attn_gn_full_b8 -f python tools/profile_unet.py --batch 8 > /dev/null 2>&1
timeout 400 $NCU --set full -k "regex:attn|gn_|layernorm" -c 10 -o $O/attn_gn_full_b1 -f python tools/profile_unet.py --batch 1 > /dev/null 2>&1
for f in igemm_full_b8 igemm_full_b1 attn_gn_full_b8 attn_gn_full_b1; do
  python tools/ncu_summary.py $O/$f.ncu-rep > $O/$f.txt 2>&1
  rm -f $O/$f.ncu-rep      # ~1.7 MB per captured launch: only the summaries travel back (64 MiB limit)
done
du -sh $O; ls -la $O
