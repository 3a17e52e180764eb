set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_final_v2.log
timeout 900 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_bench_ref.json 2> gpurun_out/r02e_bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu_final_v2.log; tail -2 gpurun_out/r02e_smoke.log
