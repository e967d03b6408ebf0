#!/bin/bash
# timing experiment: classify (128-byte lanes) and flatten side by side (variant build 'corun'); the default build beside it
KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
for ctas in 5 4 3 2; do
  echo "classify CTAs/SM = $ctas"
  SJB200_CLASSIFY_CTAS=$ctas SJB200_LIB_VARIANT=corun KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
