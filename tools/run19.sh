set -x
N=2
BNMTF_RUN_TIMING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench19_n$N.json 2> gpurun_out/r02_bench19_n$N.err
grep "bnmtf_nmf_run\|e2e-probe" gpurun_out/r02_bench19_n$N.err | tail -8 | cut -c1-250
python -c "import json;d=json.load(open('gpurun_out/r02_bench19_n$N.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['device_summary']['value'])"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench19.json 2> gpurun_out/r02_bench19.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench19.json'));print(d['value'],d['roofline']['kernel_ms'],d['e2e']['value'],d['e2e']['device_summary']['value'])"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches19.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu19.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches19.csv 10
