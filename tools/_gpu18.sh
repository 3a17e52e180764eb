set -x
timeout 600 python -m pytest tests/test_routes_gpu.py tests/test_tabular_gpu.py tests/test_fullsize_properties_gpu.py tests/test_free_running_gpu.py tests/test_inference_gpu.py -q -k "tvae or TVAE or benchmarked or free" 2>&1 | tail -4 > gpurun_out/tab_pytest.log
timeout 300 python tools/tabular_bench.py > gpurun_out/tab_bench_1m.log 2>&1
tail -4 gpurun_out/tab_pytest.log
