python tools/_bkdiag.py > gpurun_out/bkdiag.log 2>&1
python -m pytest tests/test_gpu_linsys.py tests/test_gpu_dist.py tests/test_gpu_schur.py tests/test_gpu_scale_parity.py -q -m gpu 2>&1 | tail -15 > gpurun_out/linsys_tests.log
for cfg in "8192 256 4" "4096 128 4" "1536 128 4"; do
  timeout 300 python tools/trace_potrf.py $cfg 2>&1 | tail -1
done > gpurun_out/leaf_probe.log 2>&1
