set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_check.py > gpurun_out/dp_check_oneshot.log 2>&1
tail -12 gpurun_out/dp_check_oneshot.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
tail -4 gpurun_out/r02d_bench_n2.err
