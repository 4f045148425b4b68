# compute-sanitizer over the CTA-pair kernels (small shapes): memcheck, synccheck, initcheck.  usage: gpu_sanitize.sh <outdir>
out=gpurun_out/$1; mkdir -p $out
for tool in memcheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_xgemm.py -x -q -m gpu -k "tc2_pair and (300 or 256-256 or 1000)" > $out/${tool}_xgemm_pair.txt 2>&1; echo "xgemm pair $tool rc=$?"; tail -2 $out/${tool}_xgemm_pair.txt
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "model_F or stage" > $out/${tool}_seg_pair.txt 2>&1; echo "seg pair $tool rc=$?"; tail -2 $out/${tool}_seg_pair.txt
done
