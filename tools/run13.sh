timeout 200 python tools/v5_shapes.py > gpurun_out/r02_shapes_v5.log 2>&1
BNMTF_V5=0 timeout 200 python tools/v5_shapes.py > gpurun_out/r02_shapes_v2.log 2>&1
grep "K=" gpurun_out/r02_shapes_v5.log; echo ---; grep "K=" gpurun_out/r02_shapes_v2.log
