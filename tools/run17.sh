set -x
timeout 600 python -m pytest tests/test_gpu_nmtf.py -m gpu -q -x --durations=5 > gpurun_out/r02_pytest17.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest17.log
tail -6 gpurun_out/r02_pytest17.log
timeout 300 python tools/bench_configs.py --only C4 > gpurun_out/r02_configs17.log 2>&1; tail -4 gpurun_out/r02_configs17.log | cut -c1-300
BNMTF_NP_S_PASSES=1 timeout 300 python tools/bench_configs.py --only C4np --its 1 > gpurun_out/r02_configs17_old.log 2>&1; tail -2 gpurun_out/r02_configs17_old.log | cut -c1-300
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c1_launches17.csv python tools/bench_configs.py --only C1 --its 20 > gpurun_out/r02_c1_ncu17.log 2>&1
python tools/launch_summary.py gpurun_out/r02_c1_launches17.csv 12
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c2_launches17.csv python tools/bench_configs.py --only C2 --its 20 > gpurun_out/r02_c2_ncu17.log 2>&1
python tools/launch_summary.py gpurun_out/r02_c2_launches17.csv 12
