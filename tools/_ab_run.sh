cd $GRAFT_REPO_ROOT
AB_GUIDE=gray python tools/ab_rgb.py ab=gpurun_ab/lib_g_ab.so c=gpurun_ab/lib_g_c.so m1=gpurun_ab/lib_g_m1.so m2=gpurun_ab/lib_g_m2.so all=gpurun_ab/lib_g_all.so ab_2=gpurun_ab/lib_g_ab.so c_2=gpurun_ab/lib_g_c.so m1_2=gpurun_ab/lib_g_m1.so m2_2=gpurun_ab/lib_g_m2.so all_2=gpurun_ab/lib_g_all.so 2>&1 | tee gpurun_out/ab_g4.txt
