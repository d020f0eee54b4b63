set -x
timeout 600 python -m pytest tests/test_gpu_nmtf.py tests/test_gpu_configs.py -m gpu -q -x --durations=5 -k "nmtf or c4 or gibbs or icm" > gpurun_out/r02_pytest14.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest14.log
tail -12 gpurun_out/r02_pytest14.log
timeout 300 python tools/bench_configs.py --out gpurun_out/r02_configs14.json > gpurun_out/r02_configs14.log 2>&1; tail -8 gpurun_out/r02_configs14.log
BNMTF_V5=0 timeout 200 python tools/bench_configs.py --only C4 > gpurun_out/r02_configs14_v2.log 2>&1; tail -3 gpurun_out/r02_configs14_v2.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench14.json 2> gpurun_out/r02_bench14.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench14.json'));print(d['value'],d['roofline']['kernel_ms'],d['e2e']['value'],d.get('parity_check'))"
