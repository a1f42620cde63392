for nb in 4 5 6; do for c in rotated25 xzzx21_biased xzzx21_alpha; do echo -n "minblocks $nb: "; QECMC_LIB=$PWD/mcmc-qec-toric-rl_b200/csrc/_variants/libqecmc_lad$nb.so timeout 120 python profiles/scripts/prof_ladder.py $c 200; done; done > gpurun_out/e17_ladder.log 2>&1
cat gpurun_out/e17_ladder.log
