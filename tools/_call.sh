set -x
timeout 120 python tools/perf_probe.py --case=20000,8,0.028,2 > gpurun_out/s3_ws_first.log 2>&1; echo "rc=$?" >> gpurun_out/s3_ws_first.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_ws_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s3_ws_pytest.log
tail -15 gpurun_out/s3_ws_pytest.log
timeout 300 python tools/perf_probe.py --case=500000,8,0.028,5 --case=500000,8,0.0,5 --case=500000,8,0.2,5 --case=700000,6,0.028,5 --case=1000000,4,0.028,5 --case=1300000,3,0.028,5 --ms=300000,5,3,0.028,5 --ms=400000,4,2,0.028,5 > gpurun_out/s3_ws_probe.jsonl 2>&1
cat gpurun_out/s3_ws_probe.jsonl
ncu --set full --clock-control none --import-source on -k regex:celt_synth -s 3 -c 1 -o gpurun_out/s3_C8ws -f python tools/perf_probe.py --case=500000,8,0.0,1 > gpurun_out/s3_ncu_C8ws.log 2>&1
