# compute-sanitizer memcheck over the new CTA-pair kernels (small shapes).  usage: gpu_sanitize.sh <outdir>
out=gpurun_out/$1; mkdir -p $out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_xgemm.py -x -q -m gpu -k "tc2_pair and (300 or 256-256 or 1000)" > $out/memcheck_xgemm_pair.txt 2>&1; echo "xgemm pair memcheck rc=$?"; tail -4 $out/memcheck_xgemm_pair.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "model_F or stage" > $out/memcheck_parity.txt 2>&1; echo "parity memcheck rc=$?"; tail -4 $out/memcheck_parity.txt
grep -c "Invalid\|out of bounds\|misaligned" $out/memcheck_*.txt
