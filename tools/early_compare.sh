# measurement aid: the bulk pass's early / dynamic form (default) against the static kernel (PAA_BULK_EARLY_PCT=-1)
cd $GRAFT_REPO_ROOT
tag=${1:-early_compare}
out=gpurun_out/$tag.txt
: > $out
for pct in -1 60; do
  echo "== PAA_BULK_EARLY_PCT=$pct : step trace 16 images / 2 images" >> $out
  PAA_BULK_EARLY_PCT=$pct python tools/step_trace.py 2>/dev/null >> $out
  PAA_BULK_EARLY_PCT=$pct python tools/step_trace.py --images 2 2>/dev/null >> $out
  for cfg in C2 C5 C1; do
    PAA_BULK_EARLY_PCT=$pct python bench.py --config $cfg --no-cpu-baseline --no-side --no-post --steps 40 2>/dev/null > gpurun_out/${tag}_${cfg}_$pct.json
    python - <<P >> $out
import json
d=json.load(open("gpurun_out/${tag}_${cfg}_$pct.json"))
print("$cfg pct $pct: ms_per_step %.4f  min/med/max %s  e2e %.0f" % (d["ms_per_step"], d.get("step_ms_min_med_max"), d["e2e"]["value"]))
P
  done
done
cat $out
