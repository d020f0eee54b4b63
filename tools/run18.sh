set -x
N=2
BNMTF_RUN_TIMING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench18_n$N.json 2> gpurun_out/r02_bench18_n$N.err
grep "bnmtf_nmf_run\|e2e-probe" gpurun_out/r02_bench18_n$N.err | tail -12
python -c "import json;d=json.load(open('gpurun_out/r02_bench18_n$N.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['device_summary']['value'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/bench_configs.py --only C4 --out gpurun_out/r02_c4_n$N.json > gpurun_out/r02_c4_n$N.log 2>&1
grep "^C4" gpurun_out/r02_c4_n$N.log | cut -c1-160
timeout 300 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x > gpurun_out/r02_mg_test18_n$N.log 2>&1; tail -3 gpurun_out/r02_mg_test18_n$N.log
