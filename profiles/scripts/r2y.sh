# bucket dedupe variants (experimental kernel template, not kept): plain load in front of every CAS 20.7 ms vs 15.5 ms;
# claim-and-verify rounds without atomics + counting scan 40.2 ms and WRONG (two threads holding the same key can see a
# racing slot differently and claim two slots: 2 113 745 912 vs 2 113 686 139 distinct); without the N(len) add the CAS
# loop is no faster (the add is hidden).  Results: gpurun_out/r2y_dedupe.txt
python profiles/scripts/prof_dedupe.py 0 1 2 3 4 5 > gpurun_out/r2y_dedupe.txt 2>&1; cat gpurun_out/r2y_dedupe.txt
