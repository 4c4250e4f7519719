# round 2, pass b: why is the persistent SOR pass slower per plane than the launch-per-pass kernel?
# ncu --set full of both at 512^3 (K = 1) and 257^3 periodic, plus the new bench.py line.
TAG=${1:-r2b}
mkdir -p gpurun_out
B="python bench.py --n 512 --steps 1 --warmup 3 --no-e2e --no-cpu --legs none --no-parity"
for P in 1 0; do
  O3D_SOR_PERSIST=$P timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sor_' --launch-skip 2 -c 2 -o gpurun_out/${TAG}_sor512_persist$P $B > gpurun_out/${TAG}_ncu_sor512_persist$P.log 2>&1
  ncu -i gpurun_out/${TAG}_sor512_persist$P.ncu-rep --page raw --csv > gpurun_out/${TAG}_sor512_persist${P}_raw.csv 2>/dev/null
  python profiles/ncu_summary.py gpurun_out/${TAG}_sor512_persist${P}_raw.csv > gpurun_out/${TAG}_sor512_persist${P}_summary.txt 2>&1
  rm -f gpurun_out/${TAG}_sor512_persist$P.ncu-rep
done
cat gpurun_out/${TAG}_sor512_persist*_summary.txt
for C in 3 2; do for NCH in 0 13 26; do
  O3D_PERSIST_CTAS=$C O3D_NCH_P=$NCH timeout 300 $B 2> gpurun_out/${TAG}_v.err | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l['roofline']['stages']['sor']; print('ctas/sm $C nch $NCH: sor %.4f ms/iter frac %.3f' % (s['ms_per_launch'], s['frac']))"
done; done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
tail -c 3000 gpurun_out/${TAG}_bench_default.err
python scripts/show_bench.py gpurun_out/${TAG}_bench_default.json
