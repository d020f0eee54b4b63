set -x
timeout 900 python -m pytest tests/test_gpu_nmtf.py tests/test_gpu_configs.py tests/test_gpu_traces.py -m gpu -q -x -k "nmtf or c4 or tri or trace" > gpurun_out/r02_pytest24.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest24.log
tail -4 gpurun_out/r02_pytest24.log
timeout 400 python tools/bench_configs.py --only "C4gibbs C4vb" > gpurun_out/r02_configs24.log 2>&1; grep "^C" gpurun_out/r02_configs24.log | cut -c1-130
BNMTF_S_SWEEP_BLOCK=1 timeout 400 python tools/bench_configs.py --only "C4gibbs C4vb" > gpurun_out/r02_configs24_block.log 2>&1; grep "^C" gpurun_out/r02_configs24_block.log | cut -c1-130
