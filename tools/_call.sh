set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 300 python tools/perf_probe.py --case=500000,8,0.028,5 --case=700000,6,0.028,5 --case=1000000,4,0.028,5 --ms=300000,5,3,0.028,5 --ms=400000,4,2,0.028,5 >> gpurun_out/s3_ws_probe2.jsonl 2>&1
done
cat gpurun_out/s3_ws_probe2.jsonl
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:celt_synth -s 3 -c 1 python tools/perf_probe.py --ms=300000,5,3,0.028,1 2>&1 | grep -E "gpu__time|inst_exec"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:celt_synth -s 3 -c 1 python tools/perf_probe.py --case=700000,6,0.028,1 2>&1 | grep -E "gpu__time|inst_exec"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:celt_synth -s 3 -c 1 python tools/perf_probe.py --case=500000,8,0.028,1 2>&1 | grep -E "gpu__time|inst_exec"
