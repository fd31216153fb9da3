mkdir -p gpurun_out
out=gpurun_out/gs_tune.log; : > $out
run() { g=$1; shift; echo "== g=$g $*" >> $out; env "$@" timeout 40 python tools/gs_tune.py $g lap7 >> $out 2>&1; echo "rc=$?" >> $out; }
run 24 SPB_GS_STATS=1 SPB_GS_BLOCK_ROWS=16384
run 24 SPB_GS_STATS=1 SPB_GS_BLOCK_ROWS=576
run 128 SPB_GS_STATS=1
run 128 SPB_GS_STATS=1 SPB_GS_BLOCK_ROWS=16384
run 128 SPB_X=1
cat $out
