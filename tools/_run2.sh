python tools/_bkdiag.py > gpurun_out/bkdiag.log 2>&1
python -m pytest tests/test_gpu_linsys.py -q -m gpu -x 2>&1 | tail -5 > gpurun_out/linsys_tests.log
for cfg in "8192 128,256 4 -1 -1 -1 0" "4096 128 4" "6144 128 4" "2048 128 4" "1536 128 4" "10240 256 4" "20000 512 4"; do
  timeout 300 python tools/trace_potrf.py $cfg 2>&1 | tail -2
done > gpurun_out/leaf_probe.log 2>&1
