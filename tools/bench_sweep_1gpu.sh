# Round-end single-GPU bench lines (gpurun): one JSON line each under gpurun_out/r2g_*.json
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python bench.py --metric post > gpurun_out/r2g_bench_post.json 2>> gpurun_out/r2g_bench.err
python bench.py --layout nhwc --no-post > gpurun_out/r2g_bench_nhwc.json 2>> gpurun_out/r2g_bench.err
python bench.py --metric post --layout nhwc --no-cpu-baseline > gpurun_out/r2g_bench_post_nhwc.json 2>> gpurun_out/r2g_bench.err
for c in C1 C3 C5; do python bench.py --config $c --no-post --no-side > gpurun_out/r2g_bench_$c.json 2>> gpurun_out/r2g_bench.err; done
python bench.py --images-per-gpu 2 --no-post --no-side --no-cpu-baseline > gpurun_out/r2g_bench_2img.json 2>> gpurun_out/r2g_bench.err
python bench.py --metric post --images-per-gpu 8 --no-side --no-cpu-baseline > gpurun_out/r2g_bench_post_8img.json 2>> gpurun_out/r2g_bench.err
python bench.py --impl reference > gpurun_out/r2g_ref.json 2>> gpurun_out/r2g_bench.err
python bench.py --impl reference --metric post > gpurun_out/r2g_ref_post.json 2>> gpurun_out/r2g_bench.err
(python tools/step_trace.py --images 16; python tools/step_trace.py --images 2; python tools/step_trace.py --post --images 64; python tools/step_trace.py --post --images 8) > gpurun_out/r2g_step_trace.txt 2>&1
tail -2 gpurun_out/r2g_bench.err
