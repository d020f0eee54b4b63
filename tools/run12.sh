set -x
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r02_pytest12.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest12.log
tail -4 gpurun_out/r02_pytest12.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench12.json 2> gpurun_out/r02_bench12.err
python -c "import json;d=json.load(open('gpurun_out/r02_bench12.json'));print(d['value'],d['roofline'],d['e2e'],d.get('parity_check'))"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:sweep5 --launch-skip 4 -c 1 -o gpurun_out/r02_v5_final -f python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu12.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches12.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu12b.log 2>&1
