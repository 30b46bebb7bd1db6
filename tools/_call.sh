set -x
for g in 7 6 5; do
echo "G=$g" >> gpurun_out/s3_ws_probe4.jsonl
NQ_GROUPS_PER_CTA=$g timeout 300 python tools/perf_probe.py --case=1000000,4,0.028,5 --case=1300000,3,0.028,5 --case=4000000,4,0.028,5 >> gpurun_out/s3_ws_probe4.jsonl 2>&1
done
for g in 3 2; do
echo "G=$g" >> gpurun_out/s3_ws_probe4.jsonl
NQ_GROUPS_PER_CTA=$g timeout 300 python tools/perf_probe.py --case=500000,8,0.028,5 --case=2000000,8,0.028,5 >> gpurun_out/s3_ws_probe4.jsonl 2>&1
done
cat gpurun_out/s3_ws_probe4.jsonl
