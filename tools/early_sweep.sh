# measurement aid: step trace of the loss step over the bulk pass's switches (GPU box)
cd $GRAFT_REPO_ROOT
out=gpurun_out/${1:-early_sweep}.txt
: > $out
run() {
  echo "== $*" >> $out
  env "$@" python tools/step_trace.py 2>/dev/null | grep "^select_gmm\|^bulk_focal\|^positive_list\|^pos_at_wait\|^step" >> $out
}
run PAA_BULK_EARLY_PCT=-1
for pf in 100 60; do
for pct in 0 40 50 60 70 80; do
  run PAA_BULK_EARLY_PCT=$pct PAA_L2_PREFETCH_PCT=$pf
done
done
run PAA_BULK_EARLY_PCT=60 PAA_BULK_BPS=3
cat $out
