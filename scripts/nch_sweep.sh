# z-chunk count sweep for the cost model of pick_zchunk_slots (kernels.h): per-stage ms at 256^3 / 512^3
for n in 256 512; do
for var in S 3 2; do
for nch in 1 2 3 4 5 6 8 10 12 16; do
  env O3D_NCH_$var=$nch python bench.py --n $n --steps 8 --warmup 3 --no-e2e --no-cpu > gpurun_out/nch.json 2>/dev/null
  python - <<PY
import json
l=json.loads(open("gpurun_out/nch.json").read().strip().splitlines()[-1])
st=l["roofline"]["stages"]
print("n $n NCH_$var=$nch: " + "  ".join("%s %.4f" % (k, s["ms_per_launch"]) for k,s in st.items()) + "  step %.4f" % l["ms_per_step"], flush=True)
PY
done; done; done
