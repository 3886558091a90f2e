cd $GRAFT_REPO_ROOT
L=stereo_matching_cuda_b200/libstereo_b200.so
python tools/ab_rgb.py head=$L am=gpurun_ab/lib_c_am.so head_2=$L am_2=gpurun_ab/lib_c_am.so 2>&1 | tee gpurun_out/ab_c2.txt
