# SASS evidence that the shipped library is sm_100a tcgen05 / TMA code (no GPU needed): opcode counts per kernel family.
#   bash scripts/sass_opcodes.sh > profiles/sass_opcodes.txt
LIB=satlas_super_resolution_b200/lib/libssr_b200.so
python -m satlas_super_resolution_b200.build > /dev/null
echo "# cuobjdump -sass $LIB (built from csrc/*.cu, digest $(cat satlas_super_resolution_b200/lib/libssr_b200.stamp | cut -c1-16))"
echo "# arch: $(cuobjdump -lelf $LIB | head -3 | tr '\n' ' ')"
cuobjdump -sass $LIB > /tmp/ssr_sass.txt
echo
echo "## whole library: opcode -> count"
for op in UTCHMMA UTCBAR UTMALDG UTMASTG UTMAPF LDTM STTM UTCATOMSWS SYNCS.ARRIVE SYNCS.PHASECHK UCGABAR_ARV UCGABAR_WAIT REDG ACQBULK; do
  printf "%-16s %s\n" $op $(grep -c "$op" /tmp/ssr_sass.txt)
done
echo
echo "## per kernel: UTCHMMA (tcgen05.mma) / UTMALDG (TMA load) / LDTM (tcgen05.ld) counts"
awk '/Function :/ {name=$3} /UTCHMMA/ {m[name]++} /UTMALDG/ {t[name]++} /LDTM/ {l[name]++} END {for (n in m) printf "%-100s UTCHMMA %4d  UTMALDG %3d  LDTM %3d\n", n, m[n], t[n], l[n]}' /tmp/ssr_sass.txt | sort
