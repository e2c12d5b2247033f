# run scripts/stage_times.py for the main library and every build_variants/*.so
for lib in raymond_b200/libraymond_cuda.so build_variants/*.so; do
  printf "%-36s " $(basename $lib); RAYMOND_CUDA_LIB=$PWD/$lib timeout 120 python scripts/stage_times.py ${1:-16} 2>&1 | tail -1
done
