timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash tools/profile_round.sh r1e > gpurun_out/r1e_profile_round.log 2>&1
tail -3 gpurun_out/r1e_profile_round.log
cat gpurun_out/r1e_bench_n1.json | head -c 3000
