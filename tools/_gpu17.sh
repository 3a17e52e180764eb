set -x
for i in 1 2 3 4; do timeout 300 python -m pytest tests/test_routes_gpu.py -q -k celeba 2>&1 | grep -E "AssertionError|passed|failed" | head -3; done
timeout 300 python tools/tabular_bench.py 2097152 > gpurun_out/tab_bench_2m.log 2>&1
timeout 600 python -m pytest tests/test_tabular_gpu.py tests/test_routes_gpu.py -q 2>&1 | tail -3
