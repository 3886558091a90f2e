cd $GRAFT_REPO_ROOT
L=stereo_matching_cuda_b200/libstereo_b200.so
AB_GUIDE=gray python tools/ab_rgb.py head=$L bs2=gpurun_ab/lib_g_bs2.so head_2=$L bs2_2=gpurun_ab/lib_g_bs2.so 2>&1 | tee gpurun_out/ab_g6.txt
