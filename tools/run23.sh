set -x
timeout 300 ncu --set full --import-source on --clock-control none -k regex:nmtf_s_sweep --launch-skip 4 -c 2 -o gpurun_out/r02_s_sweep -f python tools/bench_configs.py --only C4 --its 3 > gpurun_out/r02_ncu23.log 2>&1
tail -3 gpurun_out/r02_ncu23.log
ls -la gpurun_out/r02_s_sweep.ncu-rep
