TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/h8_bench.json 2> gpurun_out/h8_bench.err; tail -c 300 gpurun_out/h8_bench.json
timeout 200 $TR scripts/config4_flow.py 1e9 5e7 C4 > gpurun_out/h8_c4.json 2> gpurun_out/h8_c4.err; cat gpurun_out/h8_c4.json
timeout 200 $TR scripts/config3_flow.py 384 500000 > gpurun_out/h8_c3.json 2> gpurun_out/h8_c3.err; cat gpurun_out/h8_c3.json
for v in "NCCL_PROTO=LL128" "NCCL_PROTO=LL" "NCCL_ALGO=NVLS" "NCCL_ALGO=Ring" "NCCL_ALGO=Tree"; do echo "== $v"; env $v timeout 120 $TR scripts/allreduce_probe.py 2>/dev/null | tail -1; done > gpurun_out/h8_ar_variants.txt 2>&1; cat gpurun_out/h8_ar_variants.txt
timeout 300 python -m pytest tests/test_multigpu.py -q -m gpu 2>&1 | tail -2
