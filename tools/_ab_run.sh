cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=stereo_matching_cuda_b200/libstereo_b200.so
AB_GUIDE=gray python tools/ab_rgb.py head=$L r1=gpurun_ab/lib_g_roll1.so r2=gpurun_ab/lib_g_roll2.so r3=gpurun_ab/lib_g_roll3.so head2=$L r1_2=gpurun_ab/lib_g_roll1.so r2_2=gpurun_ab/lib_g_roll2.so r3_2=gpurun_ab/lib_g_roll3.so 2>&1 | tee gpurun_out/ab_g9.txt
