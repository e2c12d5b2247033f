for lib in build_variants/p_*.so; do printf "%-28s " $(basename $lib); RAYMOND_CUDA_LIB=$PWD/$lib timeout 120 python scripts/trav_profile.py ${1:-16} 2>&1 | tail -1; done
for lib in raymond_b200/libraymond_cuda.so build_variants/[a-oq-zA-Z]*.so; do printf "%-28s " $(basename $lib); RAYMOND_CUDA_LIB=$PWD/$lib timeout 120 python scripts/stage_times.py ${1:-16} 2>&1 | tail -1; done
