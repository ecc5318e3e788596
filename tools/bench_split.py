import sys, os
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, os.path.join(os.environ.get('GRAFT_REPO_ROOT', '/root/repo'), 'tools'))
from bench_igemm import run_case
import torch
from ldmseg import _native as nat
torch.cuda.set_device(0)
DBG = int(os.environ.get('DBG', '0'))
nat.load().ldmseg_set_debug(DBG)
print('debug flags', DBG)
for (nb, h, w, cin, n, taps) in [(1, 8, 8, 1280, 1280, 9), (1, 16, 16, 1280, 1280, 9), (1, 1, 1024, 640, 640, 1), (1, 32, 32, 640, 640, 9), (1, 64, 64, 320, 320, 9)]:
    line = f"nb={nb} {h}x{w} cin={cin} n={n} taps={taps}:"
    for bn in (64, 128, 160):
        for split in (1, 2, 3, 4, 7, 12):
            us, tf = run_case(nb, h, w, cin, n, taps, bn, split)
            line += f"  bn{bn}/s{split}: {us:6.1f}us"
    print(line, flush=True)
