set -x
python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_gpu_reference_suite.py > gpurun_out/r02_pytest5.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest5.log
python -m pytest tests/test_gpu_reference_suite.py -m gpu -q > gpurun_out/r02_refsuite5.log 2>&1; echo "rc=$?" >> gpurun_out/r02_refsuite5.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench5_v4.json 2> gpurun_out/r02_bench5_v4.err
BNMTF_V4=0 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench5_v2.json 2> gpurun_out/r02_bench5_v2.err
tail -5 gpurun_out/r02_pytest5.log; tail -3 gpurun_out/r02_refsuite5.log
for f in v4 v2; do python -c "import json;d=json.load(open('gpurun_out/r02_bench5_$f.json'));print('$f',d['value'],d['roofline']['kernel_ms'],d['config']['train_mse_last'], d.get('e2e'))"; done
