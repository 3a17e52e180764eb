set -x
timeout 900 python -m pytest tests/test_celeba_gpu.py tests/test_free_running_gpu.py tests/test_routes_gpu.py -q -k "celeba or Celeba or free" 2>&1 | tail -8 > gpurun_out/celeba_pytest.log
timeout 300 python tools/celeba_bench.py 16 20 > gpurun_out/celeba_b16.log 2>&1
timeout 300 python tools/celeba_bench.py 64 8 > gpurun_out/celeba_b64.log 2>&1
tail -8 gpurun_out/celeba_pytest.log; tail -2 gpurun_out/celeba_b16.log | cut -c1-300; tail -1 gpurun_out/celeba_b64.log | cut -c1-300
