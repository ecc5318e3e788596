#!/bin/bash
# last check of the committed state: UNet forward parity at batch 8 (pair / 320-wide plan) and at a 128x128 latent
O=gpurun_out/r04z; mkdir -p $O
export LDMSEG_PARITY_OUT=$PWD/$O/parity.json
timeout 150 python -m pytest tests/test_gpu_configs.py -x -q -s -k "unet_forward_batch8_and_latent128" > $O/pytest_unet_b8.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_unet_b8.log
