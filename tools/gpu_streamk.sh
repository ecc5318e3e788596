# stream-K tail: in-graph A/B of the UNet forward (interleaved, clocks sampled), per-op times.
TAG=${1:-r02e}
O=gpurun_out/$TAG
mkdir -p $O
for rep in 1 2; do for b in 8 4 2 1; do for sk in 0 1; do
  echo "== batch $b LDMSEG_STREAM_K=$sk"; LDMSEG_STREAM_K=$sk timeout 120 python tools/power_probe.py --batch $b --seconds 2 2>&1 | grep "t= 1"
done; done; done 2>&1 | tee $O/ab_streamk2.log
timeout 300 python tools/ablate_unet.py --batch 8 --per-op > $O/ablate_b8_v2.log 2>&1; head -4 $O/ablate_b8_v2.log
