set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt 2>&1
nproc > gpurun_out/r2a/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/r2a/host.txt; free -g >> gpurun_out/r2a/host.txt
tools/pipe_mix > gpurun_out/r2a/pipe_mix.jsonl 2>&1
tools/alu_peak > gpurun_out/r2a/alu_peak.jsonl 2>&1
which indexlr > gpurun_out/r2a/indexlr_probe.txt 2>&1; python -c "import btllib" >> gpurun_out/r2a/indexlr_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest_gpu.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a/bench_n1.json 2> gpurun_out/r2a/bench_n1.err
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --target-processes all --log-file gpurun_out/r2a/sanitizer_$tool.txt python -m pytest tests/test_gpu_sketch.py tests/test_gpu_dist.py -m gpu -x -q -k "messy or edge_cases or lockstep_nothing" > gpurun_out/r2a/sanitizer_${tool}_pytest.log 2>&1
done
echo done
