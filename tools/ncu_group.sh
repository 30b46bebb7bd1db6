# usage: bash tools/ncu_group.sh <tag> <case>...   (one ncu --set full capture per case)
tag=$1; shift
for c in "$@"; do
  C=$(echo $c | cut -d, -f2)
  ncu --set full --clock-control none --import-source on -k regex:celt_synth -s 3 -c 1 -o gpurun_out/prof_${tag}_C$C -f python tools/perf_probe.py --case=$c > gpurun_out/ncu_${tag}_C$C.log 2>&1
  ncu -i gpurun_out/prof_${tag}_C$C.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_C$C.raw.csv 2>/dev/null
done
