#!/usr/bin/env bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (tcgen05 MMA, TMA tensor/bulk copies,
# TMEM loads, cluster barriers) in the in-tree library.  No GPU needed:  tools/sass_summary.sh > profiles/sass_summary.txt
SO=${1:-distributed-vector-database_b200/libvdb_b200.so}
echo "# cuobjdump -sass $SO  ($(date -u +%F), $(/usr/local/cuda/bin/nvcc --version | tail -1))"
echo "# columns: UTCHMMA (tcgen05.mma; .2CTA = cta_group::2) | UTMALDG (cp.async.bulk.tensor) | UBLKCP (cp.async.bulk) | LDTM (tcgen05.ld) | UTCBAR (tcgen05.commit) | SYNCS (mbarrier) | UCGABAR (cluster barrier) | ACQBULK/griddepcontrol"
/usr/local/cuda/bin/cuobjdump -sass "$SO" | awk '
  /Function :/ { if (name != "") emit(); name=$3; delete c; next }
  { for (m in pat) if ($0 ~ pat[m]) c[m]++ }
  function emit(  s) {
    s = sprintf("%-96s", substr(name,1,96));
    for (i = 1; i <= n; ++i) s = s sprintf(" %s=%d", ord[i], c[ord[i]]+0);
    print s
  }
  BEGIN { n=split("UTCHMMA UTCHMMA.2CTA UTMALDG UBLKCP LDTM UTCBAR SYNCS UCGABAR_ARV ACQBULK", ord, " ");
          pat["UTCHMMA"]="UTCHMMA"; pat["UTCHMMA.2CTA"]="UTCHMMA\\.2CTA"; pat["UTMALDG"]="UTMALDG"; pat["UBLKCP"]="UBLKCP"; pat["LDTM"]="LDTM";
          pat["UTCBAR"]="UTCBAR"; pat["SYNCS"]="SYNCS"; pat["UCGABAR_ARV"]="UCGABAR"; pat["ACQBULK"]="ACQBULK" }
  END { emit() }' | c++filt | sort
