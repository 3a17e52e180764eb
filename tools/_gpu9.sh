set -x
timeout 600 python -m pytest tests/test_routes_gpu.py tests/test_tabular_gpu.py -x -q -k "tvae" 2>&1 | tail -25 > gpurun_out/mma_pytest.log
timeout 300 python tools/tabular_bench.py > gpurun_out/tab_bench_1m.log 2>&1
tail -25 gpurun_out/mma_pytest.log
