set -x
CMD="python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 python bench.py > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v2.json 2> gpurun_out/bench_ref_v2.err; echo "ref rc=$?"
$CMD > gpurun_out/plain_v2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1v2.csv $CMD > gpurun_out/ncu1_v2.log 2>&1
$CMD > gpurun_out/plain_v2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_traverse -s 5 -c 3 -o gpurun_out/prof_traverse_r1v2 $CMD > gpurun_out/ncu2_v2.log 2>&1
ls -la gpurun_out
