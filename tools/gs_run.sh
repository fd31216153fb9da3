mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gauss or gs or minres or golden" > gpurun_out/gs_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gs_pytest.log
tail -5 gpurun_out/gs_pytest.log
out=gpurun_out/gs_tune2.log; : > $out
timeout 40 python tools/gs_tune.py 96 cd27 2>&1 | grep sgs_apply >> $out
SPB_GS_STAGE_BYTES=32768 SPB_GS_STAGE_OTHER=2048 timeout 40 python tools/gs_tune.py 96 cd27 2>&1 | grep sgs_apply >> $out
SPB_GS_STAGES=4 timeout 40 python tools/gs_tune.py 96 cd27 2>&1 | grep sgs_apply >> $out
timeout 40 python tools/gs_tune.py 128 lap7 2>&1 | grep sgs_apply  >> $out
SPB_GS_STAGES=4 timeout 40 python tools/gs_tune.py 128 lap7 2>&1 | grep sgs_apply  >> $out
SPB_GS_BLOCK_ROWS=16384 timeout 40 python tools/gs_tune.py 128 lap7 2>&1 | grep sgs_apply  >> $out
SPB_GS_STATS=1 timeout 40 python tools/gs_tune.py 96 cd27 2>&1 | grep "ticket\|stats" >> $out
cat $out
