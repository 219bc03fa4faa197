#!/usr/bin/env bash
# Evidence that the GEMM path is tcgen05 / TMEM / TMA and the exchange kernel uses NVSwitch multicast: SASS opcode counts
# per kernel of the built libdinomc.so (cuobjdump -sass; no GPU needed).  Usage: tools/sass_opcodes.sh > profiles/sass_opcodes.txt
SO="self-supervised-learning-for-aerial-image-segmentation_b200/libdinomc.so"
echo "# cuobjdump -sass $SO  ($(date -u +%Y-%m-%dT%H:%MZ), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
echo "# per kernel: UTCHMMA (tcgen05.mma; .2CTA = cta_group::2) / LDTM (tcgen05.ld) / UTMALDG (TMA load) / UTMASTG (TMA store) / UTCBAR (tcgen05.commit) / LDGMC (multimem.ld_reduce: the NVSwitch adds in flight; multimem.st compiles to STG.E.128.STRONG.SYS on the multicast address) / HMMA (legacy, must be 0)"
cuobjdump -sass "$SO" | awk '
/Function :/ { if (name != "") printf "%-90s UTCHMMA=%d (2CTA=%d) UTCQMMA=%d LDTM=%d UTMALDG=%d UTMASTG=%d UTCBAR=%d LDGMC=%d HMMA=%d MUFU.EX2=%d\n", name, mma, mma2, qmma, ldtm, tmald, tmast, bar, mm, hmma, ex2;
               name=$3; mma=0; mma2=0; qmma=0; ldtm=0; tmald=0; tmast=0; bar=0; mm=0; hmma=0; ex2=0 }
/UTCHMMA/ { mma++; if ($0 ~ /2CTA/) mma2++ }
/UTCQMMA/ { qmma++ }
/LDTM/ { ldtm++ }
/UTMALDG/ { tmald++ }
/UTMASTG/ { tmast++ }
/UTCBAR/ { bar++ }
/LDGMC/ { mm++ }
/ HMMA/ { hmma++ }
/MUFU.EX2/ { ex2++ }
END { printf "%-90s UTCHMMA=%d (2CTA=%d) UTCQMMA=%d LDTM=%d UTMALDG=%d UTMASTG=%d UTCBAR=%d LDGMC=%d HMMA=%d MUFU.EX2=%d\n", name, mma, mma2, qmma, ldtm, tmald, tmast, bar, mm, hmma, ex2 }' | sed 's/_ZN3dmc[0-9]*_GLOBAL__N__[0-9a-f_]*cu_[0-9a-f]*//' | sort
