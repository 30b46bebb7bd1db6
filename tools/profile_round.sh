#!/bin/bash
# One gpurun call that produces the evidence under gpurun_out/ for profiles/ (round tag = $1):
#   bench line (not under a profiler), ncu launch list of the same command, and one ncu --set full
#   capture per kernel family (stereo 10 M frames, stereo all-transient, group 8ch, mono, 7.1 multistream,
#   5 channels with a lone mono stream, post stage single stream / many streams), each with its raw page
#   and a per-source-line summary (tools/ncu_src_summary.py).
tag=${1:-r2}
out=gpurun_out
set -x
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err || exit 1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $out/${tag}_ncu_launches.log 2>&1
full() {   # name kernel-regex skip frames-for-per-frame-numbers command...
  name=$1; rx=$2; skip=$3; frames=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $out/${tag}_full_$name -f "$@" > $out/${tag}_ncu_full_$name.log 2>&1
  ncu -i $out/${tag}_full_$name.ncu-rep --page raw --csv > $out/${tag}_full_$name.raw.csv
  python tools/ncu_src_summary.py $out/${tag}_full_$name.ncu-rep $frames 30 > $out/${tag}_src_$name.txt 2>&1
  rm -f $out/${tag}_full_$name.ncu-rep   # the raw csv page and the source summary are what profiles/ keeps; a report is 7+ MB
}
full stereo10M celt_synth 3 10000000 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras
full stereo_alltransient celt_synth 3 1000000 python tools/perf_probe.py --case=1000000,2,1.0,1
full C8 celt_synth 3 500000 python tools/perf_probe.py --case=500000,8,0.028,1
full C1 celt_synth 3 4000000 python tools/perf_probe.py --case=4000000,1,0.028,1
full ms71 celt_synth 3 500000 python tools/perf_probe.py --ms=500000,5,3,0.028,1
full C5 celt_synth 3 800000 python tools/perf_probe.py --case=800000,5,0.028,1
full post celt_post 2 4000 python tools/post_probe.py
full postmany celt_post 2 1048576 python tools/post_many_probe.py
timeout 300 python tools/perf_probe.py > $out/${tag}_perf_probe.jsonl 2>&1
timeout 300 python tools/post_probe.py > $out/${tag}_post_probe.jsonl 2>&1
ls -la $out | tail -40
