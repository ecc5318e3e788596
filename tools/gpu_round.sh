set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_2.log 2>&1; echo "pytest rc=$?"
python bench.py > gpurun_out/bench_3.json 2> gpurun_out/bench_3.err; echo "bench rc=$?"
python tools/profile_unet.py --events --batch 1 > gpurun_out/events_unet_b1_v8.log 2>&1
python tools/profile_unet.py --events --batch 8 > gpurun_out/events_unet_b8_v8.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_unet_b1_v8.csv python tools/profile_unet.py --batch 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_unet_b8_v8.csv python tools/profile_unet.py --batch 8 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:igemm_kernel -s 20 -c 12 -o gpurun_out/igemm_full_b8_v8 -f python tools/profile_unet.py --batch 8 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'attn_kernel|gn_apply' -c 8 -o gpurun_out/attn_gn_full_b8_v8 -f python tools/profile_unet.py --batch 8 > /dev/null 2>&1
tail -3 gpurun_out/pytest_gpu_2.log; cat gpurun_out/bench_3.json | head -c 600
