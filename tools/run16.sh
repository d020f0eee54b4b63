set -x
timeout 600 python -m pytest tests/test_gpu_nmtf.py tests/test_gpu_configs.py -m gpu -q -x --durations=5 -k "nmtf or c4 or vb" > gpurun_out/r02_pytest16.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest16.log
tail -15 gpurun_out/r02_pytest16.log
timeout 300 python tools/bench_configs.py --only C4 > gpurun_out/r02_configs16.log 2>&1; tail -3 gpurun_out/r02_configs16.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c4_launches16.csv python tools/bench_configs.py --only C4 --its 3 > gpurun_out/r02_c4_ncu16.log 2>&1
python tools/launch_summary.py gpurun_out/r02_c4_launches16.csv 24
