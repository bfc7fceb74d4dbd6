#!/bin/bash
# SASS evidence of the Blackwell-native paths: cuobjdump -sass of the sm_100a objects the Makefile builds (no GPU needed)
cd "$(dirname "$0")/../lancedb_b200/_lib/obj" || exit 1
OUT=../../../profiles/${1:-r02}_sass_tcgen05.txt
sass() { cuobjdump -sass "$1" | sed 's/ *\/\* 0x[0-9a-f]* \*\/ *$//; s/^ *//'; }
{
echo "# SASS evidence of the Blackwell-native paths (cuobjdump -sass of the sm_100a objects built by csrc/Makefile;"
echo "# regenerate with scripts/sass_excerpt.sh).  PTX -> SASS: tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM,"
echo "# cp.async.bulk.tensor -> UTMALDG, tcgen05.commit -> UTCBAR, tcgen05.alloc -> UTCATOMSWS, mbarrier -> SYNCS."
echo
echo "## gemm.o :: gemm_dist_kernel (coarse step, flat search)"
for m in UTCHMMA UTMALDG LDTM UTCBAR UTCATOMSWS; do echo "count $m = $(sass gemm.o | grep -c $m)"; done
echo "-- first occurrences:"
for m in UTCHMMA UTMALDG LDTM UTCBAR UTCATOMSWS; do sass gemm.o | grep -m 3 $m; done
echo
echo "## scan3.o :: scan3_kernel (filter scan): role-split registers, named barriers, 128-bit shared gathers,"
echo "## the 8x8 u16 transpose (PRMT), packed 16-bit accumulation as 3-input integer adds (IADD3), warp reductions (REDUX)"
for m in USETMAXREG "BAR.SYNC" "BAR.ARV" "LDS.128" "STS.128" PRMT IADD3 "LDG.E.128" REDUX; do echo "count $m = $(sass scan3.o | grep -c "$m")"; done
sass scan3.o | grep USETMAXREG
sass scan3.o | grep -m 2 "LDS.128"; sass scan3.o | grep -m 2 "IADD3 R"; sass scan3.o | grep -m 2 "STS.128"
echo
echo "## scan3.o :: mbarrier FULL hand-over (SYNCS) and programmatic dependent launch (griddepcontrol.launch_dependents -> PREEXIT, griddepcontrol.wait -> ACQBULK)"
for m in "SYNCS" "ACQBULK" "PREEXIT"; do echo "count $m = $(sass scan3.o | grep -c "$m")"; done
sass scan3.o | grep -m 2 "SYNCS"
echo
echo "## dist.o / gemm.o :: cp.async staging of long score rows (LDGSTS), list epilogue kept in TMEM for two passes (LDTM count above)"
echo "count LDGSTS (dist.o) = $(sass dist.o | grep -c LDGSTS)"; echo "count LDGSTS (gemm.o) = $(sass gemm.o | grep -c LDGSTS)"
echo
echo "## tables.o :: qtable_* (per-query tables): packed f32x2 FMA"
echo "count FFMA2 = $(sass tables.o | grep -c FFMA2)"; sass tables.o | grep -m 2 FFMA2
echo
echo "## scan2.o :: scan2_kernel (exact kernel): packed f32x2 arithmetic, cp.async staging"
for m in FFMA2 FADD2 LDGSTS; do echo "count $m = $(sass scan2.o | grep -c $m)"; done
} > "$OUT"
echo "wrote $OUT"
