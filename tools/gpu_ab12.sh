#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_path.py -m gpu -x -q -k "decode or encode or sample_config1" > gpurun_out/ab12_pytest.log 2>&1
timeout 600 python bench.py --workload tokenizer --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ab12_tok_new.json 2> gpurun_out/ab12_tok_new.err
