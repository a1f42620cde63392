# weight tables by 32-bit offsets, 72-register instantiation for CTAs of <= 448 threads; oracle (MT19937) decodes of configs 3 / 4 on the box's host cores
timeout 900 python -m pytest tests/test_gpu_native.py -q -k "ladder or pteq or lane_split" > gpurun_out/r2m_native.log 2>&1; tail -3 gpurun_out/r2m_native.log
for c in xzzx21_biased xzzx21_alpha rotated25 toric15; do python profiles/scripts/prof_ladder.py $c 400 4736 0.5 8; done > gpurun_out/r2m_lt.txt 2>&1
cat gpurun_out/r2m_lt.txt
timeout 1200 python profiles/scripts/run_config34.py oracle rotated25 4000000 104 > gpurun_out/r2m_rot_oracle.json 2> gpurun_out/r2m_rot_oracle.err; cut -c1-700 gpurun_out/r2m_rot_oracle.json
timeout 600 python profiles/scripts/run_config34.py oracle xzzx21_biased 2000000 128 > gpurun_out/r2m_xb_oracle.json 2> gpurun_out/r2m_xb_oracle.err; cut -c1-700 gpurun_out/r2m_xb_oracle.json
