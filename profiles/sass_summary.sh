#!/bin/bash
# SASS evidence of the Blackwell-native paths in recon_b200/libspkbgat.so (run anywhere nvcc's cuobjdump exists):
#   bash profiles/sass_summary.sh > profiles/r2_sass_summary.txt
so=recon_b200/libspkbgat.so
sass=$(mktemp)
cuobjdump -sass $so > $sass
echo "# cuobjdump -sass $so  ($(date -u +%F), $(nvcc --version | grep release | sed 's/.*release //'))"
echo "# mnemonic                      count   meaning"
for m in "UTCHMMA:tcgen05.mma (kind::tf32, 5th-gen tensor cores)" "UTCBAR:tcgen05.commit" "LDTM:tcgen05.ld (TMEM -> registers)" \
         "UTMALDG:cp.async.bulk.tensor load (TMA)" "UTMASTG:cp.async.bulk.tensor store (TMA)" "UTMAREDG:cp.reduce.async.bulk.tensor (TMA reduce-add)" \
         "UBLKCP:cp.async.bulk (1-D bulk copies of the edge streams)" "SYNCS:mbarrier ops" "UCGABAR:barrier.cluster" \
         "HMMA:legacy mma.sync (must be 0)" "HGMMA:wgmma (must be 0)" "LDGSTS:cp.async"; do
  k=${m%%:*}; d=${m#*:}
  printf "%-30s %6d   %s\n" "$k" "$(grep -c -E "(^|[^A-Z])$k" $sass)" "$d"
done
echo
echo "# kernels (entry points) per source file"
cuobjdump -elf $so 2>/dev/null | grep -o "\.text\.[A-Za-z0-9_]*" | sed 's/\.text\.//' | sort -u | wc -l | xargs echo "entry points:"
echo
echo "# kernels containing UTCHMMA"
awk '/Function :/{f=$3} /UTCHMMA/{print f}' $sass | sort | uniq -c | sed 's/_ZN3spk[0-9]*_GLOBAL__N__[0-9a-f_]*spk_gemm_tc_cu_[0-9a-f]*//' | cut -c1-120
rm -f $sass
