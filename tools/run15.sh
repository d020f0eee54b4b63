set -x
timeout 600 python -m pytest tests/test_gpu_nmtf.py tests/test_gpu_configs.py -m gpu -q -x --durations=5 -k "vb" > gpurun_out/r02_pytest15.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest15.log
tail -5 gpurun_out/r02_pytest15.log
timeout 300 python tools/bench_configs.py --only C4 > gpurun_out/r02_configs15.log 2>&1; tail -3 gpurun_out/r02_configs15.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c4_launches15.csv python tools/bench_configs.py --only C4 --its 3 > gpurun_out/r02_c4_ncu15.log 2>&1
python - <<'P'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02_c4_launches15.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%-72s n=%4d total %9.1f us  avg %8.1f us" % (k, n, t, t / n))
P
