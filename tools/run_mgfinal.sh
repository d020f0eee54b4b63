set -x
N=${1:-2}
if [ "$N" = "2" ]; then
  timeout 400 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x > gpurun_out/r02_final_mg_test_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_mg_test_n$N.log
  tail -3 gpurun_out/r02_final_mg_test_n$N.log
  grep -c "BITWISE EQUAL" gpurun_out/multigpu_check_n$N.log; grep "MISMATCH\|Error\|error" gpurun_out/multigpu_check_n$N.log | head -5
fi
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_bench_n$N.json 2> gpurun_out/r02_final_bench_n$N.err
python -c "import json;d=json.load(open('gpurun_out/r02_final_bench_n$N.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['device_summary']['value'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/bench_configs.py --only C4 --out gpurun_out/r02_final_c4_n$N.json > gpurun_out/r02_final_c4_n$N.log 2>&1
grep "^C4" gpurun_out/r02_final_c4_n$N.log | cut -c1-140
