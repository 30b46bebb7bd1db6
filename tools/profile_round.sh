#!/bin/bash
# One gpurun call that produces the evidence under gpurun_out/ for profiles/ (round tag = $1):
#   bench line (not under a profiler), ncu launch list of the same command, and one ncu --set full
#   capture per kernel family (stereo 10 M frames, group 8ch, mono, post stage).
tag=${1:-r1}
out=gpurun_out
set -x
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:celt_synth -s 3 -c 1 -o $out/${tag}_full_stereo10M -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $out/${tag}_ncu_full_stereo.log 2>&1
ncu -i $out/${tag}_full_stereo10M.ncu-rep --page raw --csv > $out/${tag}_full_stereo10M.raw.csv
for c in 500000,8,0.028,1 4000000,1,0.028,1; do
  C=$(echo $c | cut -d, -f2)
  ncu --set full --clock-control none --import-source on -k regex:celt_synth -s 3 -c 1 -o $out/${tag}_full_C$C -f \
      python tools/perf_probe.py --case=$c > $out/${tag}_ncu_full_C$C.log 2>&1
  ncu -i $out/${tag}_full_C$C.ncu-rep --page raw --csv > $out/${tag}_full_C$C.raw.csv
done
ncu --set full --clock-control none --import-source on -k regex:celt_synth -s 3 -c 1 -o $out/${tag}_full_ms71 -f \
    python tools/perf_probe.py --ms=500000,5,3,0.028,1 > $out/${tag}_ncu_full_ms71.log 2>&1
ncu -i $out/${tag}_full_ms71.ncu-rep --page raw --csv > $out/${tag}_full_ms71.raw.csv
ncu --set full --clock-control none --import-source on -k regex:celt_post -s 2 -c 1 -o $out/${tag}_full_post -f \
    python tools/post_probe.py > $out/${tag}_ncu_full_post.log 2>&1
ncu -i $out/${tag}_full_post.ncu-rep --page raw --csv > $out/${tag}_full_post.raw.csv
python tools/perf_probe.py > $out/${tag}_perf_probe.jsonl 2>&1
python tools/post_probe.py > $out/${tag}_post_probe.jsonl 2>&1
rm -f $out/${tag}_full_*.ncu-rep   # the raw csv pages are what profiles/ keeps; the reports are 7+ MB each
ls -la $out | tail -30
