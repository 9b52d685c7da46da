# clean rebuild sanity: kernel tests of the tcgen05 family + one more sample of the default bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q --timeout 300 -k "tcgen05 or channel_split or bn_act" 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2w3_bench_train.json 2> gpurun_out/r2w3_bench_train.err; echo "rc=$?"
head -c 260 gpurun_out/r2w3_bench_train.json; echo
