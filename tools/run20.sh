set -x
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r02_pytest20.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest20.log
tail -12 gpurun_out/r02_pytest20.log
timeout 300 python tools/bench_configs.py --only C4 > gpurun_out/r02_configs20.log 2>&1; grep "^C4" gpurun_out/r02_configs20.log | cut -c1-150
