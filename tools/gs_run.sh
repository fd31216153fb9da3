mkdir -p gpurun_out
out=gpurun_out/gs_tune.log; : > $out
run() { g=$1; shift; echo "== g=$g $*" >> $out; env "$@" timeout 40 python tools/gs_tune.py $g lap7 >> $out 2>&1; echo "rc=$?" >> $out; }
run 128 SPB_GS_STATS=1 SPB_GS_BLOCK_ROWS=16384
run 128 SPB_GS_BLOCK_ROWS=16384 SPB_GS_BACKOFF=50
run 128 SPB_GS_BLOCK_ROWS=16384 SPB_GS_BACKOFF=200
run 128 SPB_GS_BLOCK_ROWS=16384 SPB_GS_POLL1=1
run 128 SPB_GS_BLOCK_ROWS=16384 SPB_GS_POLL1=1 SPB_GS_BACKOFF=100
run 128 SPB_GS_BLOCK_ROWS=16384 SPB_GS_STAGE_ROWS=64 SPB_GS_STAGES=6 SPB_GS_STAGE_BYTES=4096
cat $out | grep -v "^{"
