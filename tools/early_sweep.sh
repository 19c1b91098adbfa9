# measurement aid: step trace of the loss step over the bulk pass's switches (GPU box)
cd $GRAFT_REPO_ROOT
out=gpurun_out/${1:-early_sweep}.txt
: > $out
run() {
  echo "== $*" >> $out
  env "$@" python tools/step_trace.py 2>/dev/null | grep "^select_gmm\|^bulk_focal\|^positive_list\|^step" >> $out
}
run PAA_BULK_EARLY_PCT=-1
for pb in 1184 592 296; do
for pct in 50 60 70; do
  run PAA_BULK_EARLY_PCT=$pct PAA_POS_BLOCKS=$pb
done
done
run PAA_BULK_EARLY_PCT=60 PAA_BULK_BPS=5
echo "== 2 images default" >> $out
python tools/step_trace.py --images 2 2>/dev/null | grep "^select_gmm\|^bulk_focal\|^positive_list\|^step" >> $out
echo "== 8 images default / static" >> $out
python tools/step_trace.py --images 8 2>/dev/null | grep "^select_gmm\|^bulk_focal\|^positive_list\|^step" >> $out
PAA_BULK_EARLY_PCT=-1 python tools/step_trace.py --images 8 2>/dev/null | grep "^select_gmm\|^bulk_focal\|^positive_list\|^step" >> $out
cat $out
