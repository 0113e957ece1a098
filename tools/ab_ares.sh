source tools/brun.sh
echo "sanity:"; timeout 300 python tools/sanity_small.py 2>&1 | tail -3
for m in 3 0; do echo "VDB_ARES=$m headline"; VDB_ARES=$m timeout 200 bash -c 'source tools/brun.sh; run'; done
for m in 3 0; do echo "VDB_ARES=$m c3"; VDB_ARES=$m timeout 200 bash -c 'source tools/brun.sh; run --rows 1250000 --metric l2 --k 100 --batch 4096'; done
for m in 3; do echo "VDB_ARES=$m b8192x125k"; VDB_ARES=$m timeout 200 bash -c 'source tools/brun.sh; run --rows 125000 --batch 8192'; done
echo "tf32 (VDB_SHADOW=0)"; VDB_SHADOW=0 timeout 200 bash -c 'source tools/brun.sh; run'
