set -x
BNMTF_V2_DBG=4 timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench9.json'));print('rs11',d['value'],d['roofline']['kernel'],d['roofline']['kernel_ms'],d['config']['train_mse_last'])"
BNMTF_LIB=$PWD/build/libbnmtf_prof.so BNMTF_V2_DBG=2 timeout 300 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_prof9.json 2> gpurun_out/r02_prof9.err
grep "v5 prof" gpurun_out/r02_prof9.err | tail -4
BNMTF_V5_WIDE=1 BNMTF_LIB=$PWD/build/libbnmtf_prof.so BNMTF_V2_DBG=2 timeout 300 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_prof9w.json 2> gpurun_out/r02_prof9w.err
grep "v5 prof" gpurun_out/r02_prof9w.err | tail -2
