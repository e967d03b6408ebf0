#!/bin/bash
# timing experiment: classify and flatten side by side (variant build 'corun'); the default build beside it
KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
for ctas in 0 4 3 2 1; do
  echo "classify CTAs/SM = $ctas (0 = sequential pipeline of the same build)"
  SJB200_CLASSIFY_CTAS=$ctas SJB200_LIB_VARIANT=corun KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
