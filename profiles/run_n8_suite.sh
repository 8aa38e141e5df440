T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 8 --master-port 29731 bench.py --gpus 8 --workload config4 --gather fused --steps 4 --warmup 1 > gpurun_out/r2_config4_n8_fused.json 2>gpurun_out/r2_config4_n8_fused.err
$T --nproc-per-node 8 --master-port 29732 bench.py --gpus 8 --workload config4 --gather-tiles 5 --steps 4 --warmup 1 > gpurun_out/r2_config4_n8_tiled.json 2>gpurun_out/r2_config4_n8_tiled.err
$T --nproc-per-node 8 --master-port 29733 bench.py --gpus 8 --workload graph --steps 4 --warmup 1 > gpurun_out/r2_graph_n8.json 2>gpurun_out/r2_graph_n8.err
$T --nproc-per-node 8 --master-port 29734 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8.json 2>gpurun_out/r2_bench_n8.err
$T --nproc-per-node 4 --master-port 29735 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_n4.json 2>gpurun_out/r2_bench_n4.err
$T --nproc-per-node 2 --master-port 29736 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2.json 2>gpurun_out/r2_bench_n2.err
for f in r2_config4_n8_fused r2_config4_n8_tiled r2_graph_n8 r2_bench_n8 r2_bench_n4 r2_bench_n2; do echo == $f; tail -c 600 gpurun_out/$f.json; tail -2 gpurun_out/$f.err | cut -c1-300; done
