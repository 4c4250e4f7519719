# A/B timing of kernel variants selected by environment variables (256^3 and 512^3 stage times)
run() { env "$@" python bench.py --n $N --steps 10 --warmup 4 --no-e2e --no-cpu > gpurun_out/v.json 2>/dev/null; python - "$*" <<PY
import json,sys
l=json.loads(open("gpurun_out/v.json").read().strip().splitlines()[-1])
st=l["roofline"]["stages"]
print("n $N %-24s " % sys.argv[1] + "  ".join("%s %.4f" % (k, s["ms_per_launch"]) for k,s in st.items()) + "  step %.4f" % l["ms_per_step"], flush=True)
PY
}
for N in 256 512; do
run X=0
run O3D_DIV_UNROLL=1
run O3D_CORR_VARIANT=1
run O3D_CORR_VARIANT=2
run O3D_CORR_VARIANT=3
run X=0
done
