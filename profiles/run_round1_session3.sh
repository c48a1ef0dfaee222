set -x
python __graft_entry__.py smoke > gpurun_out/smoke_s3.log 2>&1; tail -2 gpurun_out/smoke_s3.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s3c.log 2> gpurun_out/bench_s3c.err; tail -c 600 gpurun_out/bench_s3c.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_s3c.log 2>&1; tail -c 400 gpurun_out/bench_ref_s3c.log
python bench.py --steps 2 --warmup 1 --no-e2e --no-components --layers 8 > gpurun_out/b_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s3.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-components --layers 8 > gpurun_out/ncu_launch_s3.log 2>&1
python profiles/prof_kernels.py > gpurun_out/prof_plain_s3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -o gpurun_out/prof_s3 -f python profiles/prof_kernels.py > gpurun_out/ncu_full_s3.log 2>&1
tail -3 gpurun_out/ncu_full_s3.log
ls -la gpurun_out/prof_s3.ncu-rep
