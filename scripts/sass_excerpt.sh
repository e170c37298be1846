#!/bin/bash
# SASS evidence of the hot kernels (no GPU needed): mnemonic histogram + the shared / global memory and atomic
# instructions, from the objects of the last build.  Usage: scripts/sass_excerpt.sh > profiles/r02_sass_hot_kernels.txt
cd "$(dirname "$0")/../pycuda-euler_b200/csrc/build"
for spec in "bucket_part.o:bkt_partition_kernelILi20ELb0" "bucket_part.o:bkt_partition_kernelILi20ELb1" "bucket_build.o:bkt_build_kernelILi32E"; do
  obj=${spec%%:*}; pat=${spec##*:}
  echo "==== $pat ($obj, sm_100a)"
  cuobjdump -sass $obj | awk -v pat="$pat" '/Function : /{f=($0 ~ pat)} f' > /tmp/_sass.txt
  echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/_sass.txt)"
  echo "-- mnemonic histogram (top 30)"
  grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/_sass.txt | awk '{print $2}' | sed 's/;//' | sort | uniq -c | sort -rn | head -30 | awk '{printf "%s:%s  ", $2, $1} END{print ""}'
  echo "-- memory / atomic / warp-collective instructions (distinct forms)"
  grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/_sass.txt | awk '{print $2}' | sed 's/;//' | grep -E '^(LDG|STG|LDS|STS|LDL|STL|ATOM|RED|SHFL|VOTE|MATCH|BAR|WARPSYNC|NANOSLEEP|MEMBAR|CCTL|LDGSTS|UBLKCP|UTMA)' | sort | uniq -c | sort -rn | awk '{printf "%s:%s  ", $2, $1} END{print ""}'
done
