set -x
timeout 150 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "c3" > gpurun_out/r02_v5_c3.log 2>&1; echo "rc=$?" >> gpurun_out/r02_v5_c3.log
tail -3 gpurun_out/r02_v5_c3.log
timeout 120 python -m pytest tests/test_gpu_nmf.py tests/test_gpu_batch.py tests/test_gpu_summary.py -m gpu -q -x > gpurun_out/r02_v5_nmf.log 2>&1; echo "rc=$?" >> gpurun_out/r02_v5_nmf.log
tail -3 gpurun_out/r02_v5_nmf.log
BNMTF_V2_DBG=4 timeout 90 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench10.json'));print('rs11',d['value'],d['roofline']['kernel'],d['roofline']['kernel_ms'],d['config']['train_mse_last'])"
BNMTF_V5_WIDE=1 timeout 90 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench10w.json 2> gpurun_out/r02_bench10w.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench10w.json'));print('wide',d['value'],d['roofline']['kernel'],d['roofline']['kernel_ms'],d['config']['train_mse_last'])"
BNMTF_LIB=$PWD/build/libbnmtf_prof.so BNMTF_V2_DBG=2 timeout 90 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_prof10.json 2> gpurun_out/r02_prof10.err
grep "v5 prof" gpurun_out/r02_prof10.err | tail -2
