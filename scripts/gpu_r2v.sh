# round 2, final 1-GPU evidence pass: the driver's bench command, launch list, ncu --set full of the
# step kernels at 256^3, of the seam variant of the persistent SOR kernel (257^3 periodic) and of
# the LES RHS kernel (512^3)
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python scripts/show_bench.py gpurun_out/${TAG}_bench_default.json
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --legs none --no-parity"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_256.csv $B > gpurun_out/${TAG}_ncu_launch.log 2>&1
cap() { # tag, kernel regex, skip, count, bench args...
  t=$1; k=$2; sk=$3; c=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip $sk -c $c -o gpurun_out/${TAG}_$t python bench.py "$@" --no-e2e --no-cpu --legs none --no-parity > gpurun_out/${TAG}_ncu_$t.log 2>&1
  ncu -i gpurun_out/${TAG}_$t.ncu-rep --page raw --csv > gpurun_out/${TAG}_${t}_raw.csv 2>/dev/null
  python profiles/ncu_summary.py gpurun_out/${TAG}_${t}_raw.csv > gpurun_out/${TAG}_${t}_summary.txt 2>&1
  rm -f gpurun_out/${TAG}_$t.ncu-rep gpurun_out/${TAG}_${t}_raw.csv
}
cap step256 'march_kernel|sor_persist' 12 4 --steps 1 --warmup 3
cap sor257seam 'sor_persist' 3 1 --n 257 --bc periodic --steps 1 --warmup 3
cap les512 'march_kernel' 9 1 --n 512 --les --steps 1 --warmup 3
cat gpurun_out/${TAG}_*_summary.txt | grep -v "^   l1tex\|^   launch__block\|maximum_warps"
