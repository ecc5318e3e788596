# Round-2 session c: re-check what changed (split-K cooperative final pass, LN-fold epilogue), failing tests, A/Bs.
TAG=${1:-r02c}
O=gpurun_out/$TAG
mkdir -p $O
export LDMSEG_PARITY_OUT=$PWD/$O/r02_parity.json
for g in igemm_splitk igemm_pair igemm_lnfold igemm_epi igemm_f32stream gn_fused panoptic; do
  timeout 300 python tools/kernel_check.py --group $g > $O/kc_$g.log 2>&1; echo "kernel_check $g rc=$?"; grep -E "FAIL|panoptic image" $O/kc_$g.log | head -20
done
timeout 1800 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_parity.py::test_kernel_checks > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "parity\]|passed|failed|Error|^E  |FAILED" $O/pytest_gpu.log | head -80
for b in 1 8; do
  echo "== ablate batch $b (default)"; timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
  echo "== ablate batch $b LDMSEG_SPLITK_COOP=0"; LDMSEG_SPLITK_COOP=0 timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
  echo "== ablate batch $b LDMSEG_LN_FOLD=0"; LDMSEG_LN_FOLD=0 timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
  echo "== ablate batch $b LDMSEG_LN_FOLD=0 lean-epilogue build"; LDMSEG_LN_FOLD=0 LDMSEG_LIB=$PWD/latent-diffusion-segmentation_b200/lib/libldmseg_b200_lean.so timeout 200 python tools/ablate_unet.py --batch $b --full-only 2>&1 | tail -2
done
timeout 300 python tools/ablate_unet.py --batch 1 --per-op > $O/ablate_b1.log 2>&1; head -34 $O/ablate_b1.log
timeout 300 python tools/ablate_unet.py --batch 8 --per-op > $O/ablate_b8.log 2>&1; head -34 $O/ablate_b8.log
ls $O
