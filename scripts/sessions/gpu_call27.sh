#!/bin/bash
# pair kernel as default: whole GPU suite, per-shape GEMM table, Flux / Qwen benches, headline bench
mkdir -p gpurun_out
L=gpurun_out/call27.log
: > $L
echo "== smoke" >> $L
timeout 200 python __graft_entry__.py smoke >> $L 2>&1; echo "rc=$?" >> $L
echo "== pytest gpu" >> $L
timeout 900 python -m pytest tests -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
echo "== gemm shapes" >> $L
timeout 200 python scripts/gemm_shapes.py > gpurun_out/gemm_shapes_pair.json 2>> $L; echo "rc=$?" >> $L
echo "== flux" >> $L
timeout 200 python scripts/bench_flux.py --steps 10 --warmup 3 --graph > gpurun_out/bench_flux_pair.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_flux_pair.json >> $L
echo "== qwen" >> $L
timeout 200 python scripts/bench_qwen.py > gpurun_out/bench_qwen_pair.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_qwen_pair.json >> $L
echo "== bench.py" >> $L
timeout 600 python bench.py > gpurun_out/bench_n1_pair.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_n1_pair.json >> $L
tail -c 9000 $L
