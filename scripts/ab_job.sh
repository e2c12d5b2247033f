# A/B of the runtime variants on one build.  usage: bash scripts/ab_job.sh
for f in "nobin" "" "fuse,nobin" "fuse"; do for p in f64 f32shade; do python scripts/stage_times.py 32 $p dragon "$f" | tail -1; done; done
for f in "split" ""; do for p in f64 f32shade; do python scripts/stage_times.py 32 $p dof "$f" | tail -1; done; done
