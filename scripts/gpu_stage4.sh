#!/bin/bash
# parity of the new softmax math + timing of EMU variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu_tests rc=$?" | tee -a gpurun_out/summary.txt
tail -8 gpurun_out/gpu_tests.log
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[4,16,2048,64,false],[4,32,4096,64,true],[8,32,4096,128,true]]'
for v in "" _emu0 _emu2 _emu4 _emu5; do
  echo "== variant ${v:-default(emu3)}" | tee -a gpurun_out/variants.log
  TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200$v.so timeout 300 python scripts/quick_time.py "$CFG" 2>&1 | grep QT | tee -a gpurun_out/variants.log
done
