#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_path.py -m gpu -x -q -k "decode or encode or sample_config1 or smoke" > gpurun_out/ab11_pytest.log 2>&1
MASKBIT_B200_LIB=tools/lib_base.so timeout 600 python bench.py --workload tokenizer --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ab11_tok_base.json 2> gpurun_out/ab11_tok_base.err
timeout 600 python bench.py --workload tokenizer --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ab11_tok_new.json 2> gpurun_out/ab11_tok_new.err
