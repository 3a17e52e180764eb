set -x
timeout 600 python -m pytest tests/test_routes_gpu.py tests/test_tabular_gpu.py -x -q 2>&1 | tail -3 > gpurun_out/tab_pytest.log
timeout 300 python tools/tabular_bench.py 4194304 > gpurun_out/tab_bench_4m.log 2>&1
tail -3 gpurun_out/tab_pytest.log
