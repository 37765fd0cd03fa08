mkdir -p gpurun_out
V="9=2;9=2,12=1184;9=2,12=2368;9=2,12=4736;9=2,12=9472;9=2,12=18944;;12=4736;12=9472;2=2;2=2,12=4736;2=2,12=9472;6=1,3=1;6=1,3=1,12=4736;6=1,3=1,12=9472;7=4,12=9472;9=1,7=4,12=9472;9=1,7=4,23=1,12=9472;9=1,2=2,7=4,12=9472"
python scripts/interp_lab.py --variants "$V" > gpurun_out/lab4_c2.jsonl 2> gpurun_out/lab4_c2.err; tail -3 gpurun_out/lab4_c2.err
V26="9=2;9=2,12=2368;9=2,12=4736;9=2,12=9472;;12=2368;12=4736;12=9472;6=1;6=1,12=4736;6=1,12=9472"
python scripts/interp_lab.py --k26 --snapshots 2000 --layouts pitched --variants "$V26" > gpurun_out/lab4_k26.jsonl 2> gpurun_out/lab4_k26.err; tail -3 gpurun_out/lab4_k26.err
