export PASIO_B200_LIB=$PWD/build/lib_exp.so
for cfg in "0 3" "0 2" "0 1" "1 3" "2 3" "4 3" "8 3" "15 3" "7 3"; do set -- $cfg; echo "== skip=$1 ctas=$2"; PASIO_WD_SKIP=$1 PASIO_WD_CTAS=$2 python tools/prune_stats.py 100000000 2>&1 | sed -n 2,3p; done
