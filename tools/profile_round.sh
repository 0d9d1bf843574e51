#!/bin/bash
# ncu evidence for one round: launch list of one bench step + full captures of the top kernels.
# usage (under gpurun): bash tools/profile_round.sh <tag>      (then, here: python tools/make_profiles.py <tag>)
# The profiled command is the bench's own workload (1e6 lines of sight, grid 100x60x24x16), one step, no warm-up.
set -u
TAG=${1:-r01}
OUT=gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
tail -c 600 $OUT/plain_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
# name -> kernel regex (the voxel-ray and the line-of-sight instantiations of the traversal are captured separately)
# ncu matches -k against the function name without template arguments unless --kernel-name-base demangled is given
declare -A RX=( [brightness_kernel]="brightness_kernel" [march_kernel]="march_kernel" [traverse_kernel]="traverse_fast_kernel<double, .bool.1"
                [traverse_los]="traverse_fast_kernel<double, .bool.0" [gemm128_kernel]="gemm128_kernel<.int.1, " [kry_loop]="kry_loop" )
# launches of the matched kernel to skip: the first <double, false> traversal of a step is the sun-ward rays of the single
# scattering (n_vox rays), the second the lines of sight
declare -A SKIP=( [traverse_los]=1 )
for K in ${KERNELS:-brightness_kernel march_kernel traverse_kernel traverse_los kry_loop}; do
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${RX[$K]}" -s ${SKIP[$K]:-0} -c 1 -f -o $OUT/prof_${K}_$TAG $CMD > $OUT/ncu_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
ls -la $OUT | head -30
