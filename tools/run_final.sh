# the commands behind profiles/r02_final_*: full GPU test-suite, smoke, bench line, launch list (one B200)
set -x
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r02_final_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_final_pytest.log
tail -4 gpurun_out/r02_final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
python -c "import json;d=json.load(open('gpurun_out/r02_final_bench.json'));print(d['value'],d['roofline']['kernel_ms'],d['roofline']['onchip']['frac'],d['e2e']['value'],d['e2e']['device_summary']['value'],d.get('parity_check',{}).get('max_rel'),d['gpu_launches'])"
if [ "$1" = "full" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_final_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r02_final_launches.csv 12
timeout 300 ncu --set full --import-source on --clock-control none -k regex:sweep5 --launch-skip 4 -c 1 -o gpurun_out/r02_v5_final2 -f python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r02_final_ncu_full.log 2>&1
timeout 500 python tools/bench_configs.py --out gpurun_out/r02_configs_final.json > gpurun_out/r02_configs_final.log 2>&1; grep "^C" gpurun_out/r02_configs_final.log | cut -c1-200
fi
