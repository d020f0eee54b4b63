set -x
python -m pytest tests/test_gpu_nmf.py tests/test_gpu_configs.py tests/test_gpu_traces.py -m gpu -q -x --durations=8 > gpurun_out/r02_pytest2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest2.log
python -m pytest tests/test_gpu_reference_suite.py -m gpu -q > gpurun_out/r02_refsuite2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_refsuite2.log
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_v4.json 2> gpurun_out/r02_bench_v4.err
BNMTF_V4=0 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_v2split.json 2> gpurun_out/r02_bench_v2split.err
cp build/libbnmtf_nosplit.so bnmtf_ard_b200/libbnmtf_b200.so
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_v2nosplit.json 2> gpurun_out/r02_bench_v2nosplit.err
tail -4 gpurun_out/r02_pytest2.log; tail -3 gpurun_out/r02_refsuite2.log
for f in v4 v2split v2nosplit; do python -c "import json;d=json.load(open('gpurun_out/r02_bench_$f.json'));print('$f',d['value'],d['roofline']['kernel_ms'],d['config']['train_mse_last'])"; done
