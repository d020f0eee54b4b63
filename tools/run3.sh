python -m pytest tests -m gpu -q -x --durations=8 --deselect tests/test_gpu_reference_suite.py > gpurun_out/r02_pytest3.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest3.log
python -m pytest tests/test_gpu_reference_suite.py -m gpu -q > gpurun_out/r02_refsuite3.log 2>&1; echo "rc=$?" >> gpurun_out/r02_refsuite3.log
python tools/refsuite_diffs.py > gpurun_out/r02_refsuite_diffs.log 2>&1
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_v4b.json 2> gpurun_out/r02_bench_v4b.err
BNMTF_V4=0 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_v2b.json 2> gpurun_out/r02_bench_v2b.err
BNMTF_LIB=$PWD/build/libbnmtf_prof.so BNMTF_V2_DBG=2 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_prof_v4.json 2> gpurun_out/r02_prof_v4.err
tail -4 gpurun_out/r02_pytest3.log; tail -3 gpurun_out/r02_refsuite3.log; cat gpurun_out/r02_refsuite_diffs.log
for f in v4b v2b; do python -c "import json;d=json.load(open('gpurun_out/r02_bench_$f.json'));print('$f',d['value'],d['roofline']['kernel_ms'],d['config']['train_mse_last'])"; done
grep "prof side" gpurun_out/r02_prof_v4.err | tail -2
