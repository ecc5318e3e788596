"""Is the UNet forward clock- / power-limited?  Replays the captured forward back to back for a few seconds while a
thread samples SM clock, board power and throttle reasons (NVML), and prints replay time against clock over time.

Usage: python tools/power_probe.py --batch 8 [--seconds 4]
"""
import argparse
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=4.0)
    args = ap.parse_args()
    import pynvml
    from bench import build_models
    from ldmseg import _native as nat
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    unet, _, _, _ = build_models(dev)
    plan = unet._get_engine().plan(args.batch, args.size)
    plan.x_in.normal_()
    plan.run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        plan.run()
    samples, stop = [], False

    def poll():
        while not stop:
            samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            time.sleep(0.02)

    th = threading.Thread(target=poll)
    th.start()
    time.sleep(0.3)     # idle baseline
    t_start = time.time()
    times = []
    while time.time() - t_start < args.seconds:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append((time.time() - t_start, e0.elapsed_time(e1) / 10))
    stop = True
    th.join()
    print(f"batch {args.batch}: {len(times)} x 10 replays")
    n = len(times)
    for lo, hi in [(0, n // 4), (n // 4, n // 2), (n // 2, 3 * n // 4), (3 * n // 4, n)]:
        seg = times[lo:hi]
        ta, tb = seg[0][0] + t_start, seg[-1][0] + t_start
        ss = [s for s in samples if ta <= s[0] <= tb] or samples[-1:]
        clk = sorted(s[1] for s in ss)
        pw = sorted(s[2] for s in ss)
        reasons = 0
        for s in ss:
            reasons |= s[3]
        print(f"  t={seg[0][0]:4.1f}-{seg[-1][0]:4.1f}s  forward {sum(t for _, t in seg) / len(seg) * 1e3:8.1f} us   "
              f"SM clock median {clk[len(clk) // 2]} MHz (min {clk[0]})   power median {pw[len(pw) // 2]:.0f} W (max {pw[-1]:.0f})   "
              f"throttle reasons 0x{reasons:x}")
    print(f"  max SM clock {pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)} MHz, power limit "
          f"{pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0:.0f} W")


if __name__ == "__main__":
    main()
