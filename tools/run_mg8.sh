set -x
N=${1:-8}
timeout 300 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x > gpurun_out/r02_mg_test_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02_mg_test_n$N.log
tail -3 gpurun_out/r02_mg_test_n$N.log
for n in 2 4 8; do grep -c "BITWISE EQUAL" gpurun_out/multigpu_check_n$n.log; done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_n$N.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['device_summary']['value'])"
BNMTF_PEER=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_nccl_n$N.json 2> gpurun_out/r02_bench_nccl_n$N.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench_nccl_n$N.json'));print('nccl',d['value'],d['ms_per_step'])"
