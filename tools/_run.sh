cd /root/repo
ABCD_TIMING=3 ABCD_TIMING_CTA=0 timeout 300 python tools/diag.py 64 512 2>&1 | grep -A1 "dec_bwd CTA0" > gpurun_out/r2r_bwd_cta0.log
