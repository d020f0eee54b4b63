set -x
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "c3" --durations=5 > gpurun_out/r02_v5_c3.log 2>&1; echo "rc=$?" >> gpurun_out/r02_v5_c3.log
tail -5 gpurun_out/r02_v5_c3.log
timeout 600 python -m pytest tests/test_gpu_nmf.py tests/test_gpu_batch.py tests/test_gpu_summary.py -m gpu -q -x > gpurun_out/r02_v5_nmf.log 2>&1; echo "rc=$?" >> gpurun_out/r02_v5_nmf.log
tail -5 gpurun_out/r02_v5_nmf.log
for rs in 11 9; do
BNMTF_V2_DBG=4 BNMTF_V5_RS=$rs timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench7_rs$rs.json 2> gpurun_out/r02_bench7_rs$rs.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench7_rs$rs.json'));print('rs$rs',d['value'],d['roofline']['kernel'],d['roofline']['kernel_ms'],d['config']['train_mse_last'])"
grep sweep5 gpurun_out/r02_bench7_rs$rs.err | head -2
BNMTF_V5_RS=$rs timeout 600 ncu --set full --import-source on --clock-control none -k regex:sweep5 --launch-skip 4 -c 1 -o gpurun_out/r02_v5_rs$rs -f python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu7_rs$rs.log 2>&1
done
