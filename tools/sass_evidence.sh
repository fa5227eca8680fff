#!/bin/bash
# SASS mnemonic counts of the built library (runs without a GPU): proof that the hot kernels use tcgen05 (UTCHMMA),
# TMA (UTMALDG), TMEM loads (LDTM) and no legacy mma.sync / wgmma paths.  usage: tools/sass_evidence.sh > profiles/rNN_sass_evidence.txt
SO=gddim_b200/libgddim_b200.so
TMP=$(mktemp)
cuobjdump -sass $SO > $TMP
echo "# SASS evidence (cuobjdump -sass $SO, sm_100a), instruction counts"
for m in UTCHMMA UTMALDG UTCBAR LDTM "HMMA\." HGMMA; do
  echo "$m: $(grep -cE "\b$m" $TMP)"
done
echo
echo "# kernels containing UTCHMMA (count per kernel)"
awk '/Function :/{fn=$3} /UTCHMMA/{c[fn]++} END{for (f in c) print c[f], f}' $TMP | sort -rn
echo
echo "# kernels containing UTMALDG (TMA loads)"
awk '/Function :/{fn=$3} /UTMALDG/{c[fn]++} END{for (f in c) print c[f], f}' $TMP | sort -rn
rm -f $TMP
