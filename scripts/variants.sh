#!/bin/bash
# A/B harness for the level kernel: several builds of the 9x9 kernels with different -D flags, timed in ONE gpurun
# call (a call costs ~40 s of box time whatever it runs).
#   scripts/variants.sh build <tag> "<extra nvcc flags>"    here (CPU box): scratch/variants/<tag>/libofb200.so
#   scripts/variants.sh run [pairs] [reps]                   on the GPU box: parity smoke + level times per variant
# Flags understood by csrc/lk_level.cuh: -DLK_DBG_SKIP=<bits> (timing only, results wrong), -DLK_MIN_BLOCKS=n,
# -DLK_SPLIT_H=1, -DLK_MAGIC_CVT=<bits>; anything else a working copy adds.
set -e
cd "$(dirname "$0")/.."
CS=cuda_optical_flow_2_b200/csrc
case "$1" in
build)
    tag=$2; flags=$3; out=scratch/variants/$tag
    mkdir -p "$out"
    make -s -j8 -C $CS all
    nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -DLK_WIN=9 $flags \
        -c $CS/lk_win.cu -o "$out/lk_win_9.o" 2>&1 | grep -A2 "ILi9ELi2ELb0" | grep -E "registers|spill" || true
    objs=$(ls $CS/*.o | grep -v lk_win_9.o)
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libofb200.so" $objs "$out/lk_win_9.o" -ldl
    echo "$flags" > "$out/flags.txt"
    echo "built $out/libofb200.so"
    ;;
run)
    pairs=${2:-256}; reps=${3:-10}
    echo "== baseline (in-tree library)"
    python scripts/prof_pairs.py "$pairs" "$reps" | head -4
    for d in scratch/variants/*/; do
        [ -f "$d/libofb200.so" ] || continue
        echo "== $(basename "$d"): $(cat "$d/flags.txt")"
        OFB200_LIB="$PWD/$d/libofb200.so" python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
        OFB200_LIB="$PWD/$d/libofb200.so" python scripts/prof_pairs.py "$pairs" "$reps" | head -4
    done
    ;;
*) echo "usage: $0 build <tag> \"<flags>\" | run [pairs] [reps]"; exit 2;;
esac
