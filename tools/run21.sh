set -x
timeout 900 python -m pytest tests/test_gpu_nmf.py tests/test_gpu_batch.py tests/test_gpu_traces.py tests/test_gpu_summary.py tests/test_gpu_cv.py tests/test_gpu_configs.py tests/test_gpu_chains.py tests/test_gpu_reference_suite.py -m gpu -q -x -k "not c4 and not c3" > gpurun_out/r02_pytest21.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest21.log
tail -5 gpurun_out/r02_pytest21.log
timeout 400 python tools/bench_configs.py --only "C1 C2 C5" > gpurun_out/r02_configs21.log 2>&1; grep "^C" gpurun_out/r02_configs21.log | cut -c1-330
BNMTF_NO_FOLD=1 timeout 100 python tools/bench_configs.py --only "C1 C2" > gpurun_out/r02_configs21_nofold.log 2>&1; grep "^C" gpurun_out/r02_configs21_nofold.log | cut -c1-200
