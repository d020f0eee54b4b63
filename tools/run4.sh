python -m pytest tests/test_gpu_nmf.py tests/test_gpu_configs.py tests/test_gpu_batch.py tests/test_gpu_chains.py -m gpu -q -x --durations=5 > gpurun_out/r02_pytest4.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest4.log
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_v4c.json 2> gpurun_out/r02_bench_v4c.err
BNMTF_LIB=$PWD/build/libbnmtf_prof.so BNMTF_V2_DBG=2 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_prof_v4c.json 2> gpurun_out/r02_prof_v4c.err
tail -4 gpurun_out/r02_pytest4.log
for f in v4c; do python -c "import json;d=json.load(open('gpurun_out/r02_bench_$f.json'));print('$f',d['value'],d['roofline']['kernel_ms'],d['config']['train_mse_last'])"; done
grep "prof side" gpurun_out/r02_prof_v4c.err | tail -2
