#!/bin/bash
# round-2 GPU call A: full GPU test suite, default bench, sort variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
for v in default g1 g32 lb4 lb16 k8m6; do
  if [ "$v" = default ]; then unset PCH_LIB_PATH; else export PCH_LIB_PATH=$PWD/pointcloudhookup_b200/libpch_b200_$v.so; fi
  timeout 300 python tools/variant_bench.py >> gpurun_out/r2a_variants.log 2>&1
done
unset PCH_LIB_PATH
cat gpurun_out/r2a_variants.log
