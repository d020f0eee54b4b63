set -x
timeout 900 python -m pytest tests/test_gpu_nmf.py tests/test_gpu_batch.py tests/test_gpu_nmtf.py tests/test_gpu_traces.py tests/test_gpu_summary.py tests/test_gpu_configs.py -m gpu -q -x -k "not c4 and not c3" > gpurun_out/r02_pytest22.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest22.log
tail -4 gpurun_out/r02_pytest22.log
timeout 400 python tools/bench_configs.py --only "C1 C2 C4gibbs" > gpurun_out/r02_configs22.log 2>&1; grep "^C" gpurun_out/r02_configs22.log | cut -c1-160
