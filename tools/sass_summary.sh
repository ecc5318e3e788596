# SASS opcode summary of the in-tree library: evidence that the contraction kernels are Blackwell-native
# (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor loads, UBLKPF = bulk L2 prefetch, ...).
# Usage: bash tools/sass_summary.sh > profiles/r02_sass_summary.txt
LIB=latent-diffusion-segmentation_b200/lib/libldmseg_b200.so
T=$(mktemp)
cuobjdump -sass $LIB > $T 2>/dev/null
echo "library: $LIB  ($(stat -c %s $LIB) bytes)  sources @ $(git rev-parse --short HEAD)"
echo "SASS lines: $(wc -l < $T)"
echo "--- tensor-core / TMEM / TMA / mbarrier opcodes (count of SASS instructions)"
grep -o "UTCHMMA[A-Z.0-9]*\|LDTM[A-Z.0-9x]*\|STTM[A-Z.0-9x]*\|UTMALDG[A-Z.0-9]*\|UTMASTG[A-Z.0-9]*\|UTMAPF[A-Z.0-9]*\|UBLKPF[A-Z.0-9]*\|UTCBAR[A-Z.0-9]*\|UTMACCTL[A-Z.0-9]*\|UTCATOMSWS[A-Z.0-9]*\|SYNCS[A-Z.0-9]*\|ACQBULK\|FFMA2\|FADD2\|FMUL2\|MUFU\.[A-Z0-9]*\|HMMA[A-Z.0-9]*\|STG\.E\.ENL2\.256\|LDG\.E\.ENL2\.256[A-Z.]*\|RED\.[A-Z.0-9]*\|ATOMG[A-Z.0-9]*" $T | sort | uniq -c | sort -rn
echo "--- per kernel: tcgen05.mma (UTCHMMA) / TMA load (UTMALDG) instruction counts"
awk '/Function :/ {name=$3} /UTCHMMA/ {m[name]++} /UTMALDG/ {t[name]++} END {for (k in m) printf "%4d UTCHMMA %4d UTMALDG  %s\n", m[k], t[k], k}' $T | sort -k5 | c++filt | cut -c1-150
echo "UTMASTG (TMA tensor store): $(grep -c UTMASTG $T)  -- outputs leave through 256-bit per-thread stores (thread = accumulator row), see DESIGN.md section 4a"
rm -f $T
