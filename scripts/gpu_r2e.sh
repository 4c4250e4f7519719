# round 2, pass e: DRAM bytes + time of the SOR pass at 512^3 under different CTA scheduling
TAG=${1:-r2e}
mkdir -p gpurun_out
B="python bench.py --n 512 --steps 1 --warmup 3 --no-e2e --no-cpu --legs none --no-parity"
M="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 ncu $M --clock-control none -k regex:'sor_persist|sor_tma' --launch-skip $SKIP -c 1 --csv $B 2>/dev/null | grep -v "^==" | python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>5]
h=rows[0]; i=h.index('Metric Name'); v=h.index('Metric Value')
d={r[i]:r[v] for r in rows[1:]}
print('%-34s' % '$name', ' | '.join('%s=%s' % (k.split('__')[1][:18], d[k]) for k in d))"
}
SKIP=2
run "persist static nch8" O3D_NCH_P=8
run "persist static nch6" O3D_NCH_P=6
run "persist dynamic nch8" O3D_NCH_P=8 O3D_PERSIST_DYN=1
run "persist dynamic nch26" O3D_NCH_P=26 O3D_PERSIST_DYN=1
run "persist dynamic nch8 2cta" O3D_NCH_P=8 O3D_PERSIST_DYN=1 O3D_PERSIST_CTAS=2
run "oneshot nch8" O3D_NCH_P=8 O3D_PERSIST_ONESHOT=1
run "persist static nch8 promo0" O3D_NCH_P=8 O3D_TMA_L2PROMO=0
run "persist static nch8 promo256" O3D_NCH_P=8 O3D_TMA_L2PROMO=256
SKIP=8
run "launch-per-pass" O3D_SOR_PERSIST=0
run "launch-per-pass promo0" O3D_SOR_PERSIST=0 O3D_TMA_L2PROMO=0
run "launch-per-pass promo256" O3D_SOR_PERSIST=0 O3D_TMA_L2PROMO=256
show='import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=l["roofline"]["stages"]; print("  ".join("%s %.4f (%.2f)" % (k, s["ms_per_launch"], s.get("frac",0)) for k,s in st.items()))'
B2="python bench.py --n 512 --steps 6 --warmup 3 --no-e2e --no-cpu --legs none --no-parity"
for V in "A=1" "O3D_PERSIST_DYN=1" "O3D_PERSIST_DYN=1 O3D_NCH_P=13" "O3D_PERSIST_DYN=1 O3D_NCH_P=26" "O3D_TMA_L2PROMO=0" "O3D_TMA_L2PROMO=256" "O3D_SOR_PERSIST=0" "O3D_SOR_PERSIST=0 O3D_TMA_L2PROMO=0" "O3D_SOR_PERSIST=0 O3D_TMA_L2PROMO=256" "O3D_PERSIST_ONESHOT=1"; do
  echo "== $V"; env $V timeout 300 $B2 2>> gpurun_out/${TAG}.err | python -c "$show"
done
