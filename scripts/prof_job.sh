# ncu passes only (launch list + one --set full capture) on a short bench run.  usage: bash scripts/prof_job.sh <tag> [kernel-regex] [skip] [count]
set -x
TAG=${1:-r1}; KREGEX=${2:-k_traverse}; SKIP=${3:-5}; COUNT=${4:-3}
CMD="python bench.py --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
$CMD > gpurun_out/plain_${TAG}b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $COUNT -o gpurun_out/prof_${KREGEX}_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log | cut -c1-600
