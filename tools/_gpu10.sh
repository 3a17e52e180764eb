set -x
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tvae_ -s 1 -c 1 -o gpurun_out/r02_tvae_mma python tools/tab_one.py tvae > gpurun_out/ncu_tvae.log 2>&1
ls -la gpurun_out/r02_tvae_mma.ncu-rep
