# timing of the time-step stages under RHS kernel variants (O3D_RHS_VARIANT); DESIGN.md "Measured"
python -m pytest tests/test_gpu_operators.py tests/test_gpu_step.py -x -q 2>&1 | tail -2
for v in ${VARIANTS:-0}; do for n in 256 512; do for les in "" "--les"; do O3D_RHS_VARIANT=$v python bench.py --n $n $les --steps 10 --warmup 4 --no-e2e --no-cpu > gpurun_out/rhsv_${v}_${n}${les}.json 2>/dev/null; python - <<PY
import json
l=json.loads(open("gpurun_out/rhsv_${v}_${n}${les}.json").read().strip().splitlines()[-1])
st=l["roofline"]["stages"]
print("variant $v n $n $les: " + "  ".join("%s %.4f ms %.3f" % (k, s["ms_per_launch"], s["frac"]) for k,s in st.items()) + "  step ms %.4f" % l["ms_per_step"])
PY
done; done; done
