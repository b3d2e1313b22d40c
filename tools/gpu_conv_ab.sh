#!/bin/bash
# conv-path A/B on one box: parity tests of the decoder / encoder, then the tokenizer workload (BASELINE configs[3]) with the
# upsample fold and the two-stream chunk overlap switched off / on.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "decode or encode or postprocess or tokenizer or smoke" > gpurun_out/conv_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/conv_pytest.log
for v in "0 0" "1 0" "0 1" "1 1"; do
  set -- $v
  echo "== FOLD_UP=$1 DEC_OVERLAP=$2"
  MASKBIT_B200_FOLD_UP=$1 MASKBIT_B200_DEC_OVERLAP=$2 timeout 600 python bench.py --workload tokenizer --steps 3 --warmup 3 > gpurun_out/conv_tok_f$1_o$2.json 2> gpurun_out/conv_tok_f$1_o$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/conv_tok_f$1_o$2.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "img/s  e2e", round(d["e2e"]["value"],1), d.get("kernel_time_share"), d["clocks"])
PY
done
