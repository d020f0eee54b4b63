BNMTF_LIB=$PWD/build/libbnmtf_prof.so BNMTF_V2_DBG=2 timeout 300 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_prof11.json 2> gpurun_out/r02_prof11.err
grep "v5 prof" gpurun_out/r02_prof11.err | tail -2
python -c "import json;d=json.load(open('gpurun_out/r02_prof11.json'));print(d['roofline']['kernel_ms'])"
