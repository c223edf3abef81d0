#!/bin/bash
# Static evidence of the final build (no GPU needed): ptxas resource usage of every kernel and the SASS mnemonics that prove
# tcgen05 / TMEM / TMA (B200_PROFILING.md "What proves a Blackwell-native kernel").   bash tools/sass_report.sh r2
tag=${1:-r2}
cd "$(dirname "$0")/.."
out=profiles
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
{
  echo "# ptxas -v of the product sources ($(nvcc --version | tail -2 | head -1))"
  for f in conv_tc kernels engine; do
    extra=""; [ $f = kernels ] && extra="--fmad=false"
    nvcc $FLAGS $extra -c mft_b200/csrc/$f.cu -o /tmp/${f}_v.o 2>&1 | c++filt | grep -E "Compiling entry|Used|spill" | sed 's/^ptxas info    : //' |
      awk '/Compiling entry/ {name=$0; sub(/Compiling entry function ./,"",name); sub(/. for .sm_100a.$/,"",name)} /spill/ {sp=$0} /Used/ {print name " | " $0 " |" sp}'
  done
} > $out/${tag}_ptxas.txt
cuobjdump -sass mft_b200/libmft_b200.so > /tmp/${tag}_sass.txt
{
  echo "# SASS mnemonic counts per kernel of mft_b200/libmft_b200.so (cuobjdump -sass)"
  echo "# UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier"
  awk '/Function : / {name=$3} /^[[:space:]]+\/\*[0-9a-f]+\*\// {op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); split(op,a,"."); if (a[1] ~ /^(UTCHMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UTCATOMSWS|UBLKCP|HMMA|SYNCS|STL|LDL|ELECT)$/) c[name" "a[1]]++}
       END {for (k in c) print k, c[k]}' /tmp/${tag}_sass.txt | c++filt | sort
} > $out/${tag}_sass_summary.txt
wc -l $out/${tag}_ptxas.txt $out/${tag}_sass_summary.txt
