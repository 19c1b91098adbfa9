set -x
cd $GRAFT_REPO_ROOT
python tools/post_once.py 64 > gpurun_out/r2z_plain.log 2>&1 && python tools/loss_once.py 16 nchw >> gpurun_out/r2z_plain.log 2>&1 && python tools/loss_once.py 16 nhwc >> gpurun_out/r2z_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:post_ --launch-skip 22 -c 11 -f -o gpurun_out/r2z_post python tools/post_once.py 64 > gpurun_out/r2z_ncu_post.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"prep_step|iou_match|match_score|select_gmm|bulk_focal|positive_list" --launch-skip 12 -c 6 -f -o gpurun_out/r2z_loss python tools/loss_once.py 16 nchw > gpurun_out/r2z_ncu_loss.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"match_score|positive_list" --launch-skip 4 -c 2 -f -o gpurun_out/r2z_loss_nhwc python tools/loss_once.py 16 nhwc > gpurun_out/r2z_ncu_loss_nhwc.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side --no-graph > gpurun_out/r2z_bench_nograph.json 2> gpurun_out/r2z_bench.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side --no-graph > gpurun_out/r2z_ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
