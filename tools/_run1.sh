python -m pytest tests/test_gpu_linsys.py tests/test_gpu_dist.py tests/test_gpu_schur.py tests/test_gpu_scale_parity.py -q -m gpu -k "indefinite or ldl or bunch or nan or alternative" 2>&1 | tail -40 > gpurun_out/bk_tests.log
for cfg in "8192 128,256 4 -1 -1 -1 0" "8192 128,256 4 -1 -1 -1 1" "4096 128 4 -1 -1 0 0" "4096 128 4 -1 -1 0 1" "4096 128 4 -1 -1 1 0" "6144 128 4 -1 -1 0 0" "6144 128 4 -1 -1 0 1" "6144 128 4 -1 -1 1 0" "10240 256 4 -1 -1 -1 0" "10240 256 4 -1 -1 -1 1" "2048 128 4 -1 -1 0 1" "2048 128 4 -1 -1 1 0"; do
  timeout 300 python tools/trace_potrf.py $cfg 2>&1 | tail -3
done > gpurun_out/part_probe.log 2>&1
