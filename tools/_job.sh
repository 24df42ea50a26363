T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $T bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/r02k_n2_reference.json 2> gpurun_out/r02k_n2_reference.err; tail -c 600 gpurun_out/r02k_n2_reference.json; echo
timeout 900 $T bench.py --gpus 2 > gpurun_out/r02k_n2_bench_c2.json 2> gpurun_out/r02k_n2_bench_c2.err; tail -c 2500 gpurun_out/r02k_n2_bench_c2.json; echo; tail -3 gpurun_out/r02k_n2_bench_c2.err
