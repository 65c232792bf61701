#!/bin/bash
# sweep of the CTA shape of the generated small-model kernel (CA_SMALL_SHAPE=threads,stages,ctas)
#   tools/small_shape_sweep.sh J,K,L nstates[,...] shape [shape ...]
jkl=$1; ns=$2; shift 2
for sh in "$@"; do
  CA_SMALL_SHAPE=$sh python tools/small_case.py $jkl $ns 2>/dev/null
done
